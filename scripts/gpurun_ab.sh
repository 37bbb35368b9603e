cd /root/repo
echo default; python scripts/seeds.py 500 501 502 503 504 505
for k in $SKIPS; do echo skips $k; DRB_MARCH_SKIPS=$k python scripts/seeds.py 500 501 502 503 504 505; done
for v in $VARIANTS; do echo $v; DRB_LIB_PATH=variants/libdregb200_$v.so python scripts/seeds.py 500 501 502 503 504 505; done
