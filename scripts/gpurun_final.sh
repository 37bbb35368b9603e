cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --stage register --precision bf16 2>&1 | tail -1 > gpurun_out/bench_register_bf16.json
python -c "
import json; d=json.load(open('gpurun_out/bench_register_bf16.json')); print('REGISTER bf16 value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'igemm TF/s', round(d['roofline']['achieved'],1), 'frac', round(d['roofline']['frac'],3))"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16 2>&1 | tail -1 > gpurun_out/bench_full_bf16.json
python -c "
import json; d=json.load(open('gpurun_out/bench_full_bf16.json')); print('FULL bf16 value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_full.csv > gpurun_out/launch_summary_full.txt; head -14 gpurun_out/launch_summary_full.txt
