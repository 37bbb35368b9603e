cd /root/repo
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_extract_gpu.py -x 2>&1 | tail -3
echo default; python scripts/seeds.py 500 501 502 505
for v in $VARIANTS; do echo $v; DRB_LIB_PATH=variants/libdregb200_$v.so python scripts/seeds.py 500 501 502 505; done
