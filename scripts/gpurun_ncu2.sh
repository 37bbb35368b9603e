cd /root/repo
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:surface_mask -s 4 -c 1 -o gpurun_out/prof_surface_r01_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_surface2.log 2>&1
tail -1 gpurun_out/ncu_surface2.log | cut -c1-100
