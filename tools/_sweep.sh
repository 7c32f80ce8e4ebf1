python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v19.log 2>&1; tail -4 gpurun_out/smoke_v19.log
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
python bench.py > gpurun_out/bench_v19.json 2> gpurun_out/bench_v19.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_pre_ncu.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v19.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v19.json 2> gpurun_out/bench_ref_v19.err
tail -c 300 gpurun_out/bench_ref_v19.json
