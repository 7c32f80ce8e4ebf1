python bench.py > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_pre_ncu.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v18.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_target.py hdv3 8 > gpurun_out/ncu_target.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/full_v18 -f python tools/ncu_target.py hdv3 8 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/full_v18.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/full_v18_raw.csv.gz
gzip -c /tmp/full_v18.ncu-rep > gpurun_out/full_v18.ncu-rep.gz
du -sh gpurun_out; ls -la gpurun_out | tail -8
