#!/bin/bash
# Second profiling pass of round 2 (after dt_variant 3 became the default): bench lines of every BASELINE config that runs the detector,
# launch lists, and --set full captures of the kernels that changed.  Usage: tools/gpurun_retry.sh 3000 'bash tools/gpu_round2b.sh'
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
timeout 600 python bench.py > $O/r2_bench_vga.json 2> $O/r2_bench_vga.err; echo "bench vga exit $?"
timeout 300 python bench.py --config vga1 > $O/r2_bench_vga1.json 2> $O/r2_bench_vga1.err; echo "bench vga1 exit $?"
timeout 600 python bench.py --config 1080p > $O/r2_bench_1080p.json 2> $O/r2_bench_1080p.err; echo "bench 1080p exit $?"
timeout 600 python bench.py --mode tensor16 > $O/r2_bench_vga_tensor16.json 2> $O/r2_bench_vga_tensor16.err; echo "bench tensor16 exit $?"
timeout 300 python bench.py --no-cpu --parity-frames 0 --opt dt_variant=0 > $O/r2_bench_vga_dt_variant0.json 2> $O/r2_bench_vga_dt_variant0.err; echo "bench dt_variant 0 exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_batch32_exact.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 0 --nms 0.1 > $O/r2_launches_exact.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_batch32_tensor16.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 3 > $O/r2_launches_tensor16.log 2>&1
for K in dt_pass_win; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^${K}|::${K}" -c 4 -f -o $O/r2_full_${K} \
      python tools/run_step.py --batch 32 --steps 1 --mode 0 > $O/r2_full_${K}.log 2>&1
done
python tools/run_step.py --batch 64 --steps 4 | tail -2
ls $O | wc -l
