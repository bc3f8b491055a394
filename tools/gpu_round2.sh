#!/bin/bash
# One gpurun call of round 2: GPU parity tests, one bench line per BASELINE config, ncu launch lists and --set full captures of every
# kernel of the path.  Usage (from the repo root): tools/gpurun_retry.sh 3000 'bash tools/gpu_round2.sh'
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 600 python bench.py > $O/r2_bench_vga.json 2> $O/r2_bench_vga.err; echo "bench vga exit $?"
timeout 300 python bench.py --config vga1 > $O/r2_bench_vga1.json 2> $O/r2_bench_vga1.err; echo "bench vga1 exit $?"
timeout 600 python bench.py --config 1080p > $O/r2_bench_1080p.json 2> $O/r2_bench_1080p.err; echo "bench 1080p exit $?"
timeout 900 python bench.py --config dt > $O/r2_bench_dt.json 2> $O/r2_bench_dt.err; echo "bench dt exit $?"
timeout 300 python bench.py --mode tensor16 --no-cpu --parity-frames 0 > $O/r2_bench_vga_tensor16.json 2> $O/r2_bench_vga_tensor16.err; echo "bench tensor16 exit $?"
# launch lists (per-launch durations; cold-cache and serialised under ncu: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_batch32_exact.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 0 --nms 0.1 > $O/r2_launches_exact.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_batch32_tensor16.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 3 > $O/r2_launches_tensor16.log 2>&1
# one --set full capture per kernel of the path (first launch of each, batch 32)
for K in pyr_resize_u8 pyr_down_u8 hog_hist hog_feat part_response feat_split_f16 part_response_tc dt_pass mix_max root_select hits_select backtrack root_nms nms_; do
  MODE=0; [ "$K" = feat_split_f16 ] && MODE=3; [ "$K" = part_response_tc ] && MODE=3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^${K}|::${K}" -c 2 -f -o $O/r2_full_${K} \
      python tools/run_step.py --batch 32 --steps 1 --mode $MODE --nms 0.1 --root-nms 2 > $O/r2_full_${K}.log 2>&1
done
timeout 200 ncu --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_op_hmma.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg \
    --clock-control none -k regex:part_response_tc -c 1 --csv --log-file $O/r2_tensor_pipe.csv python tools/run_step.py --batch 32 --steps 1 --mode 3 > $O/r2_tensor_pipe.log 2>&1
# two detectors on two streams in exact mode: does the DP of one batch hide under the response kernel of the other?
timeout 300 python tools/two_stream.py --batch 128 --steps 5 --mode 0 > $O/r2_two_stream_exact.log 2>&1; cat $O/r2_two_stream_exact.log
ls $O | wc -l
