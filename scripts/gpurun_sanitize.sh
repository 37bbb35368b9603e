cd /root/repo
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=surface_mask --print-limit 20 python -m pytest -q -m gpu -p no:cacheprovider tests/test_extract_gpu.py -k "block_small or visibility" -x > gpurun_out/racecheck.log 2>&1
grep -c "Race reported\|hazard" gpurun_out/racecheck.log; grep -m8 "hazard\|Race\|ERROR SUMMARY\|passed\|failed" gpurun_out/racecheck.log
timeout 300 compute-sanitizer --tool memcheck --kernel-regex kns=surface_mask --print-limit 10 python -m pytest -q -m gpu -p no:cacheprovider tests/test_extract_gpu.py -k "block_small or visibility" -x > gpurun_out/memcheck.log 2>&1
grep -m6 "Invalid\|ERROR SUMMARY\|passed\|failed" gpurun_out/memcheck.log
