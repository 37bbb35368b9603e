cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests -x 2>&1 | tail -2
for ST in register full; do timeout 600 python bench.py --steps 9 --warmup 3 --no-cpu-baseline --stage $ST 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d.get('roofline_register', d['roofline']); print('$ST value', round(d['value'],3), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],3), 'igemm ms', round(r['kernel_ms_per_step'],3), 'TF/s', round(r['achieved'],1), d['config'].get('extract'), 'launches', d['gpu_launches'])"; done
