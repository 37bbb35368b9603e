cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'surface_mask|ngp_rgb|ngp_density' -s 6 -c 3 -f -o gpurun_out/extract_kernels python scripts/extract_once.py 4 501 > gpurun_out/ncu_extract.log 2>&1
tail -2 gpurun_out/ncu_extract.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mha_core' -s 30 -c 2 -f -o gpurun_out/mha_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --stage register > gpurun_out/ncu_mha.log 2>&1
tail -2 gpurun_out/ncu_mha.log
