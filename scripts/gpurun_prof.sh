cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_full_v2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_full_v2.json')); print('FULL value', d['value'], 'e2e', d['e2e']['value'], d['config']['extract']['stage_ms'], d['roofline']['achieved'])"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 > gpurun_out/bench_register_v2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_register_v2.json')); print('REGISTER value', d['value'], 'e2e', d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'])"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01_v2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_v2.log 2>&1
wc -l gpurun_out/launches_r01_v2.csv
