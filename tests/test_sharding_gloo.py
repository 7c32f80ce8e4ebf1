"""N > 1 path on CPU: two gloo ranks shard frame pairs contiguously, no data-path collective, ordered gather."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import importlib, os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    pkg = importlib.import_module("eccv2022-rife_b200")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    frames = [torch.full((3, 4, 4), float(i)) for i in range(8)]           # 7 pairs
    calls = []
    def fake_infer(a, b, t):                                               # stands in for Model.inference
        calls.append(float(a.mean()))
        return a * (1 - t) + b * t
    out = pkg.run_sharded(fake_infer, frames, timesteps=(0.25, 0.75))
    lo, hi = pkg.shard_range(7, dist.get_rank(), dist.get_world_size())
    assert calls == [float(i) for i in range(lo, hi) for _ in (0, 1)], calls   # only this rank's pairs were computed
    assert len(out) == 14
    want = [i + t for i in range(7) for t in (0.25, 0.75)]
    got = [float(o.mean()) for o in out]
    assert got == want, got                                                # global order = (pair, timestep)
    dist.barrier()
    dist.destroy_process_group()
    print("ok", dist_rank := int(os.environ["RANK"]))
""") % ROOT


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err
        assert out.startswith("ok")
