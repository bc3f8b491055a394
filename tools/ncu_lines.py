"""Per-source-line instruction counts and stall samples of one kernel of an .ncu-rep (captured with --import-source on, built -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, items, tot, tots = None, None, [], 0, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        ti = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= ii or not r[0]:
        continue
    try:
        n, s, t = int(r[ii]), int(r[si]), int(r[ti])
    except ValueError:
        continue
    items.append((n, s, t, cur_file, r[0], r[1].strip()[:120]))
    tot += n
    tots += s
print("kernel %s: %d warp instructions, %d stall samples" % (rx, tot, tots))
for n, s, t, f, l, src in sorted(items, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% samp  lanes {t / max(n, 1):4.1f}  {f}:{l:>4s}  {src}")
