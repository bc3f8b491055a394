"""Prints the headline fields of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), "frames/s device,", round(d["e2e"]["value"]), "end to end;", {k: round(v, 3) for k, v in d["stage_ms"].items()},
      {k: round(v, 3) for k, v in d["kernel_ms"].items()}, "parity", d.get("parity"))
