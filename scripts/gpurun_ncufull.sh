cd /root/repo
mkdir -p gpurun_out
# igemm: the dominant conv (3^3 256->256 at 64^3) is launch ~ index; capture a few igemm launches after warm-up
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 150 -c 6 -o gpurun_out/prof_igemm_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --stage register > gpurun_out/ncu_igemm.log 2>&1
tail -1 gpurun_out/ncu_igemm.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:surface_mask -s 4 -c 1 -o gpurun_out/prof_surface_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_surface.log 2>&1
tail -1 gpurun_out/ncu_surface.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
