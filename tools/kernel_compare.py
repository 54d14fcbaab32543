#!/usr/bin/env python3
"""Development: kernel time of several kernel ids on one BASELINE config, results compared with the first kernel listed.

    python tools/kernel_compare.py cnot2 16384 3,4        (needs a GPU)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import juqbox_b200 as jq
    from juqbox_b200 import _lib, configs
    from bench import alg_flops_per_eval
    name, B = sys.argv[1], int(sys.argv[2])
    kernels = [int(k) for k in sys.argv[3].split(",")]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    dev = torch.device("cuda", 0)
    peak = _lib.fp64_peak_tflops(0)
    cfg = configs.example(name)
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
    nsamp = 1 if shifts is None else len(shifts)
    flops = alg_flops_per_eval(cfg.params, cfg.nCoeff)
    sh = torch.from_numpy(shifts).to(dev) if shifts is not None else None
    pc = torch.from_numpy(configs.synthetic_pcof(cfg, B)).to(dev)
    ref = None
    for k in kernels:
        wa = jq.Working_Arrays(cfg.params, cfg.nCoeff, device=0)
        try:
            wa.set_kernel(k)
        except Exception as e:
            print(f"{name} kernel {k}: unavailable ({e})")
            wa.close()
            continue
        ms = []
        out = None
        for it in range(reps + 1):
            out = wa.evaluate_device(pc, sh, None, True, out=out)
            torch.cuda.synchronize()
            if it >= 1:
                ms.append(wa.last_kernel_ms)
        t = float(np.mean(ms)) * 1e-3
        ev = B * nsamp / t
        g = out["grad"].cpu().numpy()
        f = (out["infid"] + out["leak"]).cpu().numpy()
        if ref is None:
            ref = (f, g)
        dg = np.linalg.norm(g - ref[1]) / np.linalg.norm(ref[1])
        df = np.max(np.abs(f - ref[0]) / np.abs(ref[0]))
        print(f"{name} B={B} kernel {k}: {t * 1e3:.2f} ms  {ev:.5g} evals/s  {ev * flops / 1e12:.2f} TFLOP/s alg = {100 * ev * flops / 1e12 / peak:.1f}% of {peak:.1f}"
              f"  regs {int(wa.query(5))} ctas {int(wa.query(4))} traj/cta {int(wa.query(3))} smem {int(wa.query(6))}  relerr vs first: obj {df:.2e} grad {dg:.2e}", flush=True)
        wa.close()


if __name__ == "__main__":
    main()
