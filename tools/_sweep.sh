python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
run() { tag=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ck_$tag.json 2> gpurun_out/ck_$tag.err; }
run p12 --pairs 12
run p24c12 --pairs 24 --chunk 12
run p24c8 --pairs 24 --chunk 8
run p24c6 --pairs 24 --chunk 6
run p24c4 --pairs 24 --chunk 4
run p32c8 --pairs 32 --chunk 8
run p12c6 --pairs 12 --chunk 6
run p12c4 --pairs 12 --chunk 4
