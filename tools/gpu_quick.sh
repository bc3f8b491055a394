#!/bin/bash
# Short GPU check: parity tests, two bench lines, a launch list, the stream-split experiment.
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --batch 32 --no-cpu > $O/${TAG}_bench_b32.json 2>> $O/${TAG}_bench.err
timeout 300 python tools/two_stream.py --batch 64 --steps 10 > $O/${TAG}_two_stream.txt 2>&1
cat $O/${TAG}_two_stream.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_batch32_tensor.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 2 > $O/${TAG}_launches.log 2>&1
python - <<PY
import json
for f in ("$O/${TAG}_bench.json", "$O/${TAG}_bench_b32.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 3) for k, v in d["stage_ms"].items()}, {k: round(v, 3) for k, v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 $O/${TAG}_bench.err
