cd /root/repo
mkdir -p gpurun_out
DRB_LIB_PATH=variants/libdregb200_t384u2.so timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=surface_mask --print-limit 30 python scripts/extract_once.py 1 505 > gpurun_out/racecheck_t384u2.log 2>&1
grep -m12 "hazard\|Race\|RACECHECK SUMMARY\|rep 0" gpurun_out/racecheck_t384u2.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=surface_mask --print-limit 30 python scripts/extract_once.py 1 505 > gpurun_out/racecheck_default_128.log 2>&1
grep -m6 "hazard\|Race\|RACECHECK SUMMARY\|rep 0" gpurun_out/racecheck_default_128.log
