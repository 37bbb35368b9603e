cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_extract_gpu.py -x -s 2>&1 | tail -7
python scripts/seeds.py 500 501 502 505
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/extract_launches.csv python scripts/extract_once.py 3 500 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/extract_launches.csv | head -8
