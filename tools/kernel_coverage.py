"""Which kernel does the library pick for problem shapes beyond the five named configs, is it right, how fast is it?

    python tools/kernel_coverage.py [--batch 2048]          (needs a GPU; the oracle is the checker)

Prints one markdown row per shape: auto-selected kernel (1 generic, 2 slot, 3 fibre, 4 tile, 5 latency layout), why the register-resident
kernels declined (if they did), relative gradient error vs the CPU oracle on two candidates, evals/s on `batch`.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq                                    # noqa: E402
from juqbox_b200 import configs                             # noqa: E402
from oracle import oracle_traceobjgrad                      # noqa: E402

SHAPES = [
    ([2], [0], {}), ([2], [1], {}), ([3], [1], {}), ([4], [2], {}), ([3], [2], {}), ([4], [4], {}), ([5], [3], {}),
    ([2, 2], [0, 0], {}), ([2, 2], [1, 1], {}), ([2, 2], [1, 2], {}), ([2, 2], [2, 2], {}), ([3, 3], [1, 1], {}),
    ([3, 3], [2, 2], {}), ([2, 3], [2, 1], {}), ([2, 2, 2], [0, 0, 0], {}), ([2, 2, 2], [1, 1, 1], {}),
    ([3, 2], [2, 1], {}), ([3, 3], [3, 1], {}), ([3], [2], {"use_sparse": True}), ([3, 2, 2], [2, 1, 1], {}),
    ([2, 2, 1], [2, 2, 3], {}), ([2, 2], [2, 2], {"exchange": 0.02}), ([3, 3], [1, 1], {"exchange": 0.02}), ([2, 2, 2], [1, 1, 1], {"exchange": 0.02}),
    ([2, 2], [2, 2], {"Nfreq": 3}), ([4], [2], {"Nfreq": 1}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    a = ap.parse_args()
    print("| Ne / Ng | extra | n x m | nsteps | J | kernel | declined because | rel. grad err | evals/s |")
    print("|---|---|---|---|---|---|---|---|---|")
    for Ne, Ng, kw in SHAPES:
        cfg = configs.qudit_system(Ne, Ng, **kw)
        p = cfg.params
        wa = jq.Working_Arrays(p, cfg.nCoeff)
        rng = np.random.default_rng(3)
        pc = rng.uniform(-1, 1, (2, cfg.nCoeff)) * cfg.maxpar[0] * 0.5
        r = wa.evaluate(pc)
        o = oracle_traceobjgrad(p, pc)
        err = max(np.linalg.norm(r["grad"][b, 0] - o["grad"][b, 0]) / np.linalg.norm(o["grad"][b, 0]) for b in range(2))
        errf = np.abs(r["infid"] - o["infid"]).max()
        k = wa.last_kernel
        why = ""
        if k == 1:
            for kid in (4, 3, 2):
                try:
                    wa.set_kernel(kid)
                except Exception as e:                       # the library's reason string
                    why += f"[{kid}] {str(e).split(':')[-1].strip()} "
            wa.set_kernel(0)
        big = rng.uniform(-1, 1, (a.batch if k != 1 else max(a.batch // 16, 8), cfg.nCoeff)) * cfg.maxpar[0] * 0.5
        wa.evaluate(big)
        wa.evaluate(big)
        rate = big.shape[0] / (wa.last_kernel_ms * 1e-3)
        k = f"{wa.last_kernel} ({k} for 2 candidates)" if wa.last_kernel != k else k
        print(f"| {Ne} / {Ng} | {kw or ''} | {p.Ntot} x {p.N} | {p.nsteps} | {p.linear_solver.max_iter} | {k} | {why} | "
              f"{err:.1e} (infid {errf:.0e}) | {rate:.3g} |", flush=True)
        wa.close()


if __name__ == "__main__":
    main()
