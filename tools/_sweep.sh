python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
for p in 8 12; do python bench.py --pairs $p --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/nw_p$p.json 2> gpurun_out/nw_p$p.err; done
python tools/e2e_probe.py 8 > gpurun_out/e2e_probe.log 2>&1; python tools/e2e_probe.py 12 >> gpurun_out/e2e_probe.log 2>&1
RIFE_HK_DEBUG=1 python tools/hk_debug_run.py 2>&1 | tail -42 | grep "kind=2" > gpurun_out/hk_dbg_narrow.log
