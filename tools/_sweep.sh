timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/san_mem.log 2>&1; echo "memcheck rc=$?" ; tail -4 gpurun_out/san_mem.log
RIFE_HK_CHAIN=64 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/san_mem_chain.log 2>&1; echo "memcheck chain rc=$?"; tail -3 gpurun_out/san_mem_chain.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/san_race.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_race.log
