cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,driver_version --format=csv
T="timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 180"
$T tests/test_kernels_gpu.py -k "not igemm and not im2col" 2>&1 | tail -30
echo ==== igemm linear
$T tests/test_kernels_gpu.py -k "igemm_linear" -x 2>&1 | tail -30
echo ==== igemm conv
$T tests/test_kernels_gpu.py -k "igemm_conv3d or epilogue or im2col" 2>&1 | tail -40
echo ==== forward
$T tests/test_forward_gpu.py -x -s 2>&1 | tail -40
