cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_full.json
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); r=d['roofline']; print('FULL value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['config']['extract']['stage_ms'], 'surf ms/launch', round(r['kernel_ms_per_launch'],2), 'share', round(r['kernel_share_of_step'],3), 'samples/s %.3g' % r['density_samples_per_s'], 'issued GB/s', round(r['issued_gather_gbs'],1), r['per_launch'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --stage register 2>&1 | tail -1 > gpurun_out/bench_register.json
python -c "
import json; d=json.load(open('gpurun_out/bench_register.json')); print('REGISTER value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['roofline']['achieved'],1), d['roofline']['kernel_share_of_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_full.csv > gpurun_out/launch_summary_full.txt; head -30 gpurun_out/launch_summary_full.txt
