cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests -x -s 2>&1 | grep -v "^$" | tail -12
