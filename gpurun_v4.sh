cd /root/repo
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -m gpu -p no:cacheprovider"
$T tests/test_kernels_gpu.py tests/test_forward_gpu.py -s 2>&1 | grep -E "passed|failed|relative errors|Error" | tail -8
DRB_IGEMM_CLUSTER=2 $T tests/test_kernels_gpu.py -k "igemm" 2>&1 | tail -1
for NS in 1 0; do DRB_IGEMM_NO_SPLITK=$NS timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nosplit=$NS REGISTER value', round(d['value'],3), 'ms', round(d['ms_per_step'],3), 'igemm ms', round(d['roofline']['kernel_ms_per_step'],3), 'TF/s', round(d['roofline']['achieved'],1))"; done
