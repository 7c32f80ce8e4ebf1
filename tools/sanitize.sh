#!/bin/bash
# compute-sanitizer over the op-level tcgen05 tests and one small Model.inference; logs under gpurun_out/
# (copy the summaries to profiles/).  Run on the GPU box:  bash tools/sanitize.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/san_small.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import AMP, seeded_model
pkg = importlib.import_module("eccv2022-rife_b200")
for prec in ("fp16", "bf16", "fp32"):
    m = seeded_model(pkg, AMP, prec)
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(1, 3, 64, 128, generator=g).cuda(), torch.rand(1, 3, 64, 128, generator=g).cuda()
    if prec != "fp32":
        a, b = a.half(), b.half()
    for _ in range(3):            # plain, CUDA-graph capture + replay, replay
        out = m.inference(a, b, 0.4)
    torch.cuda.synchronize()
    print(prec, float(out.float().mean()))
print("done")
PY
for tool in memcheck racecheck; do
  echo "== $tool: small Model.inference (fp16 / bf16 / fp32, graph replay included)" > gpurun_out/r02_sanitizer_$tool.txt
  timeout 400 $SAN --tool $tool --print-limit 20 python /tmp/san_small.py >> gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/r02_sanitizer_$tool.txt
done
echo "== memcheck: tests/test_gpu_ops.py (tcgen05 conv / conv-transpose / warp / resize operators)" > gpurun_out/r02_sanitizer_ops.txt
timeout 500 $SAN --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "tcgen05 and (16-16 or 32-32 or 64-64 or 96-96 or 32-52 or 64-52)" >> gpurun_out/r02_sanitizer_ops.txt 2>&1
echo "exit code $?" >> gpurun_out/r02_sanitizer_ops.txt
tail -5 gpurun_out/r02_sanitizer_*.txt
