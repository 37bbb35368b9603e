#!/bin/bash
# launch list of one bf16 training step (2 pairs, one stream) after the weight-gradient change
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_train_r02_wg.csv python bench.py --stage train --batch 2 --streams 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1; echo "ncu list rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_train_r02_wg.csv | head -60
