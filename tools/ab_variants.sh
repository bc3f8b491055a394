#!/bin/bash
# A/B of library variants built by tools/build_variant.sh: per-stage device times of the same step (batch 64 VGA frames, exact mode)
# usage: tools/ab_variants.sh "run_step options" variant...
OPTS=$1; shift
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then L=""; else L=build/variants/libpbd_b200_$v.so; fi
  PBD_B200_LIB=$L timeout 200 python tools/run_step.py --batch 64 --steps 6 $OPTS 2>&1 | tail -2 | head -1
done
