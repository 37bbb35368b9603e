cd /root/repo
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -m gpu -p no:cacheprovider"
$T tests/test_extract_gpu.py 2>&1 | tail -2
for CS in 1 2 4; do echo cluster $CS; DRB_IGEMM_CLUSTER=$CS $T tests/test_kernels_gpu.py -k "igemm or im2col" 2>&1 | tail -1; done
$T tests/test_forward_gpu.py 2>&1 | tail -1
for TH in 256 512 1024; do DRB_SURFACE_THREADS=$TH timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('threads $TH FULL value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['config']['extract']['stage_ms'], d['config']['masked_voxels'])"; done
for CS in 1 2 4; do DRB_IGEMM_CLUSTER=$CS timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cluster $CS REGISTER value', round(d['value'],3), 'igemm ms', round(d['roofline']['kernel_ms_per_step'],3), 'TF/s', round(d['roofline']['achieved'],1))"; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('auto REGISTER value', round(d['value'],3), 'igemm ms', round(d['roofline']['kernel_ms_per_step'],3), 'TF/s', round(d['roofline']['achieved'],1))"
