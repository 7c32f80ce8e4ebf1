timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
run() { tag=$1; shift; timeout 300 env "$@" python bench.py --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/ch_$tag.json 2> gpurun_out/ch_$tag.err; }
run c64 A=1
run layer RIFE_HK_CHAIN=0
run c96 RIFE_HK_CHAIN=96
run c32 RIFE_HK_CHAIN=32
run c64b A=1
run layerb RIFE_HK_CHAIN=0
