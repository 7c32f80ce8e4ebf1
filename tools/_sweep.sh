run() { tag=$1; shift; timeout 300 env "$@" python bench.py --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/ue_$tag.json 2> gpurun_out/ue_$tag.err; }
run on A=1
run off RIFE_UMMA_EARLY=0
run on2 A=1
run off2 RIFE_UMMA_EARLY=0
python tools/latency_probe.py 2>&1 | tail -4
RIFE_UMMA_EARLY=0 python tools/latency_probe.py 2>&1 | tail -4
