cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py tests/test_forward_gpu.py -x 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 > gpurun_out/bench_register.json
python -c "
import json; d=json.load(open('gpurun_out/bench_register.json')); print('REGISTER value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), round(d['roofline']['achieved'],1), d['roofline']['kernel_share_of_step'])"
