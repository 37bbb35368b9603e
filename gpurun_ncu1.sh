cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01_v1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_v1.log 2>&1
tail -2 gpurun_out/ncu_bench_v1.log
wc -l gpurun_out/launches_r01_v1.csv
