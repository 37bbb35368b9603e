cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 169 -c 5 -f -o gpurun_out/igemm_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --stage register > gpurun_out/ncu_igemm.log 2>&1
tail -2 gpurun_out/ncu_igemm.log
