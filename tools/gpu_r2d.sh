#!/bin/bash
# final check of the second session: full GPU suite, the bench lines that changed last, a 64-frame step
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2d_pytest_full.log 2>&1; echo "pytest exit $?" >> $O/r2d_pytest_full.log; tail -3 $O/r2d_pytest_full.log
timeout 300 python bench.py --config vga1 > $O/r2c_bench_vga1.json 2> $O/r2c_bench_vga1.err; python tools/print_bench.py $O/r2c_bench_vga1.json | cut -c1-330
timeout 300 python bench.py --config vga1 --no-cpu --opt dt_segment=0 > $O/r2c_bench_vga1_noseg.json 2> $O/r2c_bench_vga1_noseg.err; python tools/print_bench.py $O/r2c_bench_vga1_noseg.json | cut -c1-330
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches_batch1_exact.csv python tools/run_step.py --batch 1 --steps 2 --mode 0 > $O/r2c_launches_b1.log 2>&1
python tools/run_step.py --batch 64 --steps 4 --mode 0 | tail -1
