cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -s 2>&1 | grep -vE "^\s*$" | tail -40
