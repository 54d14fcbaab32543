"""2-GPU test of the sample-sharded risk-neutral evaluation with the library's own NCCL all-reduce
(jq_comm_init): every rank must obtain the same weighted sums as one GPU evaluating all samples.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from juqbox_b200.distributed import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps = 30.0, 600
    pc = configs.synthetic_pcof(cfg, 3) * 10
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes)
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff, device=rank)
    wa.comm_init(rank, world)
    lo, hi = shard_range(len(cfg.nodes), rank, world)
    r = wa.evaluate(pc, shifts[lo:hi], cfg.weights[lo:hi])            # ends with the NCCL all-reduce
    per = wa.evaluate(pc, shifts[lo:hi])                             # no weights -> no communication
    q.put((rank, r["infid"], r["leak"], r["grad"], per["infid"].shape))
    dist.barrier()
    wa.comm_destroy()
    wa.close()
    dist.destroy_process_group()


def test_two_gpu_sample_shards_allreduce():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps = 30.0, 600
    pc = configs.synthetic_pcof(cfg, 3) * 10
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff, device=0)
    want = wa.evaluate(pc, configs.noise_shift(cfg.params.Ntot, cfg.nodes), cfg.weights)
    wa.close()
    for rank, infid, leak, grad, shp in got:
        assert np.allclose(infid, want["infid"], rtol=1e-13, atol=1e-15)
        assert np.allclose(leak, want["leak"], rtol=1e-12, atol=1e-18)
        assert np.linalg.norm(grad - want["grad"]) <= 1e-12 * np.linalg.norm(want["grad"])
    assert np.array_equal(got[0][3], got[1][3])                      # identical on both ranks


def _coop_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import time
    import torch
    import torch.distributed as dist
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    os.environ["JQ_SEG_COOP_FORCE"] = "1"          # share out the small problem too (by default only launches of >= 2 waves)
    out = []
    for name, kw in (("three_qudits", dict(T=40.0)), ("cnot2", {})):
        cfg = configs.qudit_system([2, 2, 1], [2, 2, 3], **kw) if name == "three_qudits" else configs.example(name)
        pc = np.random.default_rng(4).uniform(-1, 1, (2, cfg.nCoeff)) * cfg.maxpar[0] * 0.3
        wa = jq.Working_Arrays(cfg.params, cfg.nCoeff, device=rank)
        wa.set_kernel(7)
        wa.comm_init(rank, world)
        wa.comm_set_cooperative(True)
        r = wa.evaluate(pc)                                           # both ranks: the same arguments
        dist.barrier()
        t0 = time.perf_counter()
        r = wa.evaluate(pc)
        ms = (time.perf_counter() - t0) * 1e3
        f = wa.evaluate(pc, evaladjoint=False)
        out.append((name, r["infid"], r["leak"], r["grad"], f["infid"], int(wa.query(7)), ms))
        dist.barrier()
        wa.comm_destroy()
        wa.close()
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_gpus_share_one_evaluation():
    """jq_comm_set_cooperative: the time segments of ONE time-parallel evaluation are shared out over the ranks, the propagators
    all-gathered; both ranks must return exactly the bits a single GPU returns with the same number of segments."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_coop_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for idx, (name, kw) in enumerate((("three_qudits", dict(T=40.0)), ("cnot2", {}))):
        cfg = configs.qudit_system([2, 2, 1], [2, 2, 3], **kw) if name == "three_qudits" else configs.example(name)
        pc = np.random.default_rng(4).uniform(-1, 1, (2, cfg.nCoeff)) * cfg.maxpar[0] * 0.3
        a, b = got[0][idx], got[1][idx]
        assert a[5] == b[5] and a[5] % 2 == 0
        wa = jq.Working_Arrays(cfg.params, cfg.nCoeff, device=0)
        wa.set_kernel(7)
        wa.set_time_segments(a[5])
        want = wa.evaluate(pc)
        wa.set_kernel(3)
        plain = wa.evaluate(pc)
        wa.close()
        for r in (a, b):
            assert np.array_equal(r[1], want["infid"]) and np.array_equal(r[2], want["leak"]) and np.array_equal(r[3], want["grad"]), name
            assert np.array_equal(r[4], want["infid"])
        assert np.linalg.norm(a[3] - plain["grad"]) <= 1e-12 * np.linalg.norm(plain["grad"])
        print(name, "segments", a[5], "host call ms on two GPUs", a[6], b[6])
