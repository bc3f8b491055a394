#!/bin/bash
# compute-sanitizer passes over one small end-to-end detection (smoke(): 160x208 frame, all stages, compared with the oracle).
# Usage (on a GPU box): bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
  echo "exit=$?"
done
