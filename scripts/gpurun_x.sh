cd /root/repo
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_extract_gpu.py -x -s -k visibility 2>&1 | grep -v "^ \|^$\|^E " | tail -12
