#!/bin/bash
# Profiling pass of the round-2 second session: launch lists (batch 32 exact / tensor16, single frame) and --set full captures of the
# kernels that changed (dt_pass_win with source, mix_max, hog_hist).  Usage: tools/gpurun_retry.sh 2000 'bash tools/gpu_prof_r2c.sh'
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c_smi.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c_launches_batch32_exact.csv python tools/run_step.py --batch 32 --steps 2 --mode 0 > $O/r2c_launches_exact.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c_launches_batch32_tensor16.csv python tools/run_step.py --batch 32 --steps 2 --mode 3 > $O/r2c_launches_tensor16.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches_batch1_exact.csv python tools/run_step.py --batch 1 --steps 2 --mode 0 > $O/r2c_launches_b1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dt_pass_win" -c 4 -f -o $O/r2c_full_dt_pass_win python tools/run_step.py --batch 32 --steps 1 --mode 0 > $O/r2c_full_dt_pass_win.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"mix_max" -c 2 -f -o $O/r2c_full_mix_max python tools/run_step.py --batch 32 --steps 1 --mode 0 > $O/r2c_full_mix_max.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hog_hist" -c 1 -f -o $O/r2c_full_hog_hist python tools/run_step.py --batch 32 --steps 1 --mode 3 > $O/r2c_full_hog_hist.log 2>&1
ls -la $O | tail -8
