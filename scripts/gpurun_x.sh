cd /root/repo
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py tests/test_forward_gpu.py -x 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('REGISTER value', round(d['value'],2), 'ms', round(d['ms_per_step'],3))"; done
