"""Multi-GPU: shard the naturally independent axes of the path across ranks (one process per GPU).

The reference has no parallelism at all; its two independent axes are the noise samples of the risk-neutral loop
(src/ipopt_interface.jl:38-65, examples/Risk_Neutral/run_all.jl:9-28) and candidate pcof vectors.  Each rank owns
a contiguous block; candidates need no exchange, the risk-neutral objective needs exactly one all-reduce (sum) of
[infid, leak, infidgrad(Npar), leakgrad(Npar)] per evaluation, with the quadrature weights applied on the device
before the reduction.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import numpy as np

from .configs import noise_shift


def shard_range(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of n items for `rank`; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_partial(r, npar: int, obj_func_type: int) -> np.ndarray:
    """[infid, leak, infidgrad(npar), leakgrad(npar)] of one candidate's weighted partial sums."""
    out = np.zeros(2 + 2 * npar)
    out[0], out[1] = r["infid"][0], r["leak"][0]
    out[2:2 + npar] = r["infidgrad"][0] if "infidgrad" in r else r["grad"][0]
    if obj_func_type != 1:
        out[2 + npar:] = r["leakgrad"][0]
    return out


def risk_neutral_eval(pcof, params, nodes, weights, evaluate, group=None):
    """eval_f_g_grad! over ranks: `evaluate(pcof[None], shifts, weights)` is the per-rank batched evaluation
    (Working_Arrays.evaluate on a GPU).  Returns (infid, leak, infidgrad, leakgrad) identical on every rank.

    The two reduction mechanisms are mutually exclusive: a handle with an attached NCCL communicator (Working_Arrays.comm_init)
    already all-reduces its weighted sums inside the C ABI, so handing its `evaluate` to this function would reduce twice."""
    import torch
    import torch.distributed as dist
    owner = getattr(evaluate, "__self__", None)
    if getattr(owner, "comm_size", 1) > 1:
        raise ValueError("risk_neutral_eval: this Working_Arrays has its own NCCL communicator (comm_init) and reduces inside "
                         "evaluate(); call evaluate() with the rank's shard directly, or comm_destroy() first")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nodes, weights = np.atleast_1d(np.asarray(nodes, float)), np.atleast_1d(np.asarray(weights, float))
    lo, hi = shard_range(len(nodes), rank, world)
    npar = len(pcof)
    if hi > lo:
        r = evaluate(np.asarray(pcof, float)[None, :], noise_shift(params.Ntot, nodes[lo:hi]), weights[lo:hi])
        part = pack_partial(r, npar, params.objFuncType)
    else:
        part = np.zeros(2 + 2 * npar)
    t = torch.from_numpy(part)
    if world > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
    v = t.numpy()
    return float(v[0]), float(v[1]), v[2:2 + npar].copy(), v[2 + npar:].copy()
