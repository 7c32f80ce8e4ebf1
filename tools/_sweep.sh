python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
run() { tag=$1; shift; env "$@" python bench.py --pairs 8 --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; }
run new A=1
run old RIFE_BI_OLD2=1
run new2 A=1
run old2 RIFE_BI_OLD2=1
