"""pairs sharded over several ranks give the same frames, in the same order, as one rank.
Two processes (one per GPU when the box has two, else both on cuda:0), each computing its contiguous range of pairs
through Model.inference on the GPU; the ordered gather is run_sharded's single all_gather."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import importlib, os, sys
    sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
    import torch, torch.distributed as dist
    from conftest import AMP, seeded_model
    pkg = importlib.import_module("eccv2022-rife_b200")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"
    torch.cuda.set_device(rank %% ngpu)
    dist.init_process_group(backend, init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=rank, world_size=world)
    m = seeded_model(pkg, AMP, "fp16")
    g = torch.Generator().manual_seed(5)
    frames = [torch.rand(1, 3, 128, 192, generator=g).cuda().half() for _ in range(6)]       # 5 pairs over 2 ranks: 3 + 2
    out = pkg.run_sharded(lambda a, b, t: m.inference(a, b, t), frames, timesteps=(0.25, 0.5))
    assert len(out) == 10
    if rank == 0:
        k = 0
        for i in range(5):
            for t in (0.25, 0.5):
                want = m.inference(frames[i], frames[i + 1], t)
                assert torch.equal(out[k].to(want.device), want), (i, t)                     # bit-identical, (pair, timestep) order
                k += 1
    dist.barrier()
    dist.destroy_process_group()
    print("ok", rank, backend)
""") % (ROOT, ROOT)


def test_two_ranks_reproduce_one_rank_bit_for_bit(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-3000:]
        assert out.startswith("ok")
