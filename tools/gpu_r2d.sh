O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2d_pytest_full.log 2>&1; echo "pytest exit $?" >> $O/r2d_pytest_full.log; tail -3 $O/r2d_pytest_full.log
python tools/run_step.py --batch 64 --steps 4 --mode 0 | tail -1
python tools/run_step.py --batch 8 --steps 4 --mode 0 --h 1080 --w 1920 --max-levels 10 | tail -1
