O=gpurun_out; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c_launches_batch32_exact.csv python tools/run_step.py --batch 32 --steps 2 --mode 0 > $O/r2c_launches_exact.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c_launches_batch32_tensor16.csv python tools/run_step.py --batch 32 --steps 2 --mode 3 > $O/r2c_launches_tensor16.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches_batch1_exact.csv python tools/run_step.py --batch 1 --steps 2 --mode 0 > $O/r2c_launches_b1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dt_pass_win" -c 4 -f -o $O/r2c_full_dt_pass_win python tools/run_step.py --batch 64 --steps 1 --mode 0 --opt dp_streams=1 > $O/r2c_full_dt_pass_win.log 2>&1
ls -la $O | tail -5
