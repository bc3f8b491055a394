#!/bin/bash
# A/B of library variants built by tools/build_variant.sh: per-stage device times of the same step (batch 64 VGA frames, tensor mode)
for v in "$@"; do
  echo "== $v"
  PBD_B200_LIB=build/variants/libpbd_b200_$v.so timeout 200 python tools/run_step.py --batch 64 --steps 6 --mode 3 2>&1 | tail -1
done
