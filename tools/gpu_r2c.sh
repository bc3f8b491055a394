#!/bin/bash
# Round-2 (second session) quick GPU check: DT parity subset + A/B of dt.cu variants (tools/build_variant.sh).
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "windowed or variants or dt2d or bitexact or golden or groups" > $O/r2c_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2c_pytest.log
tail -3 $O/r2c_pytest.log
tools/ab_variants.sh "--mode 3" "$@" 2>&1 | tee $O/r2c_ab.txt
