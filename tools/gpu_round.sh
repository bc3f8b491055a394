#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one --set full capture of the top kernels.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --batch 32 --no-cpu > $O/${TAG}_bench_b32.json 2>> $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_batch32_tensor16.csv \
    python tools/run_step.py --batch 32 --steps 2 --mode 3 > $O/${TAG}_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"dt_pass|part_response_tc|mix_max|feat_split" -c 7 -f -o $O/${TAG}_top_full \
    python tools/run_step.py --batch 32 --steps 1 --mode 3 > $O/${TAG}_full.log 2>&1
tail -2 $O/${TAG}_full.log
cat $O/${TAG}_bench.json | cut -c1-1500
