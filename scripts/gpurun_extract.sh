cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_extract_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v "^E   *+" | tail -40
timeout 300 python scripts/extract_once.py 5
