cd /root/repo
mkdir -p gpurun_out
SEED=${SEED:-501}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:surface_mask -s 2 -c 1 -f -o gpurun_out/surface_full python scripts/extract_once.py 3 $SEED > gpurun_out/ncu_surface.log 2>&1
tail -1 gpurun_out/ncu_surface.log
