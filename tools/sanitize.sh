#!/bin/bash
# compute-sanitizer passes over (a) one small end-to-end detection (smoke(): 160x208 frame, all stages, compared with the oracle) and
# (b) tools/sanitize_paths.py (tensor kernels, concurrent DP streams, device NMS, pipelined API).
# Usage (on a GPU box): bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool : smoke() ==="
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  echo "exit=$?"
  echo "=== compute-sanitizer --tool $tool : tools/sanitize_paths.py ==="
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_paths.py 2>&1 | tail -6
  echo "exit=$?"
done
