#!/bin/bash
# Round-2 (second session) quick GPU check: DT parity subset + single-frame config per dt_segment + A/B of dt.cu variants (tools/build_variant.sh).
O=gpurun_out; mkdir -p $O
timeout 800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${PYK:-windowed or segmented or variants or dt2d or bitexact or golden or groups or vga}" > $O/r2c_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2c_pytest.log
tail -5 $O/r2c_pytest.log
[ -n "$DEBUGSEG" ] && timeout 300 python tools/debug_seg.py $DEBUGSEG 2>&1 | tail -40
for seg in ${SEGS:-0 -1 32 48 64}; do
  echo "== vga1 dt_segment $seg"
  timeout 200 python bench.py --config vga1 --no-cpu --opt dt_segment=$seg > $O/r2c_vga1_$seg.json 2> $O/r2c_vga1_$seg.err; python tools/print_bench.py $O/r2c_vga1_$seg.json 2>/dev/null || tail -3 $O/r2c_vga1_$seg.err
done
[ -n "$1" ] && tools/ab_variants.sh "--mode 3" "$@" 2>&1 | tee $O/r2c_ab.txt
exit 0
