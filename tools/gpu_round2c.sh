#!/bin/bash
# Third profiling pass of round 2 (ring-split tier 1, mid-round exit, segmented walk for small launches): the full GPU suite, one bench
# line per BASELINE config, launch lists, and a --set full capture of dt_pass_win.  Usage: tools/gpurun_retry.sh 3000 'bash tools/gpu_round2c.sh'
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest_full.log 2>&1; echo "pytest exit $?" >> $O/r2c_pytest_full.log; tail -3 $O/r2c_pytest_full.log
timeout 600 python bench.py > $O/r2c_bench_vga.json 2> $O/r2c_bench_vga.err; echo "bench vga exit $?"
timeout 300 python bench.py --config vga1 > $O/r2c_bench_vga1.json 2> $O/r2c_bench_vga1.err; echo "bench vga1 exit $?"
timeout 300 python bench.py --config vga1 --no-cpu --opt dt_segment=0 > $O/r2c_bench_vga1_noseg.json 2> $O/r2c_bench_vga1_noseg.err; echo "bench vga1 noseg exit $?"
timeout 600 python bench.py --config 1080p > $O/r2c_bench_1080p.json 2> $O/r2c_bench_1080p.err; echo "bench 1080p exit $?"
timeout 600 python bench.py --mode tensor16 > $O/r2c_bench_vga_tensor16.json 2> $O/r2c_bench_vga_tensor16.err; echo "bench tensor16 exit $?"
if [ -z "$NOPROF" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c_launches_batch32_exact.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 0 > $O/r2c_launches_exact.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c_launches_batch32_tensor16.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 3 > $O/r2c_launches_tensor16.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches_batch1_exact.csv \
    python tools/run_step.py --batch 1 --steps 2 --mode 0 > $O/r2c_launches_b1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dt_pass_win" -c 4 -f -o $O/r2c_full_dt_pass_win \
    python tools/run_step.py --batch 32 --steps 1 --mode 0 > $O/r2c_full_dt_pass_win.log 2>&1
fi
for f in vga vga1 vga1_noseg 1080p vga_tensor16; do echo "== $f"; python tools/print_bench.py $O/r2c_bench_$f.json 2>/dev/null || tail -2 $O/r2c_bench_$f.err; done
ls $O | wc -l
