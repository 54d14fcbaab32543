"""world_size-2 gloo test (CPU) of the sample-sharded risk-neutral evaluation: shard arithmetic + one all-reduce.
The per-rank evaluator is the CPU oracle here (no GPU in this container); on the GPU box the same function is
driven by Working_Arrays.evaluate in tests/test_gpu_parity.py::test_risk_neutral_weighted_sum and bench.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from juqbox_b200.distributed import shard_range
    for n in (0, 1, 7, 9, 1001):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1


def _oracle_eval(params):
    from oracle import oracle_traceobjgrad

    def ev(pcof, shifts, weights):
        o = oracle_traceobjgrad(params, pcof, shifts)
        w = np.asarray(weights)
        return {"infid": (o["infid"] * w).sum(1), "leak": (o["leak"] * w).sum(1),
                "grad": (o["grad"] * w[None, :, None]).sum(1), "infidgrad": (o["infidgrad"] * w[None, :, None]).sum(1),
                "leakgrad": (o["leakgrad"] * w[None, :, None]).sum(1)}
    return ev


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from juqbox_b200 import configs
    from juqbox_b200.distributed import risk_neutral_eval
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps = 30.0, 400
    pc = configs.synthetic_pcof(cfg, 1)[0] * 10
    res = risk_neutral_eval(pc, cfg.params, cfg.nodes, cfg.weights, _oracle_eval(cfg.params))
    q.put((rank, res[0], res[1], res[2]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_risk_neutral_matches_single_process():
    import torch.multiprocessing as mp
    from juqbox_b200 import configs
    from juqbox_b200.distributed import risk_neutral_eval
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps = 30.0, 400
    pc = configs.synthetic_pcof(cfg, 1)[0] * 10
    want = risk_neutral_eval(pc, cfg.params, cfg.nodes, cfg.weights, _oracle_eval(cfg.params))   # world = 1
    for rank, infid, leak, g in got:
        assert abs(infid - want[0]) < 1e-13 and abs(leak - want[1]) < 1e-15
        assert np.linalg.norm(g - want[2]) <= 1e-13 * np.linalg.norm(want[2])
    assert got[0][1] == got[1][1] and np.array_equal(got[0][3], got[1][3])      # identical on every rank
