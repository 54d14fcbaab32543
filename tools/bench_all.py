#!/usr/bin/env python3
"""Throughput table for every BASELINE config and batch size: evals/s, state-steps/s, fraction of
the measured FP64 peak, through the device-resident C-ABI entry point.  Writes profiles/<tag>_throughput.{json,md}.

    python tools/bench_all.py [tag]        (needs a GPU; run under gpurun and copy gpurun_out/ back)
"""
import json
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
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    dev = torch.device("cuda", 0)
    peak = _lib.fp64_peak_tflops(0)
    rows = []
    for name in configs.EXAMPLES:
        cfg = configs.example(name)
        shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
        nsamp = 1 if shifts is None else len(shifts)
        flops = alg_flops_per_eval(cfg.params, cfg.nCoeff)
        wa = jq.Working_Arrays(cfg.params, cfg.nCoeff, device=0)
        sh = torch.from_numpy(shifts).to(dev) if shifts is not None else None
        batches = [1, 64, 1024, 16384] if name != "cnot3" else [1, 148, 1184, 2368, 4736]
        if name == "rabi":
            batches.append(262144)
        for B in batches:
            pc = torch.from_numpy(configs.synthetic_pcof(cfg, B)).to(dev)
            out = None
            ms = []
            for it in range(5):
                out = wa.evaluate_device(pc, sh, None, True, out=out)
                torch.cuda.synchronize()
                if it >= 2:
                    ms.append(wa.last_kernel_ms)
            t = float(np.mean(ms)) * 1e-3
            ev = B * nsamp / t
            rows.append({"config": name, "kernel": wa.last_kernel, "candidates": B, "noise_samples": nsamp, "kernel_ms": t * 1e3,
                         "evals_per_sec": ev, "state_steps_per_sec": ev * 3 * cfg.params.nsteps,
                         "alg_tflops": ev * flops / 1e12, "frac_fp64_peak": ev * flops / 1e12 / peak})
            print(rows[-1], flush=True)
        wa.close()
    outdir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(outdir, exist_ok=True)
    json.dump({"fp64_peak_tflops_measured": peak, "rows": rows}, open(os.path.join(outdir, f"{tag}_throughput.json"), "w"), indent=1)
    with open(os.path.join(outdir, f"{tag}_throughput.md"), "w") as f:
        f.write(f"Measured FP64 FMA peak (jq_fp64_peak): {peak:.2f} TFLOP/s. Kernel time = CUDA events around the trajectory kernel.\n\n")
        f.write("| config | kernel | candidates x samples | kernel ms | evals/s | state-steps/s | alg. TFLOP/s | of FP64 peak |\n|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['config']} | {r['kernel']} | {r['candidates']} x {r['noise_samples']} | {r['kernel_ms']:.2f} | {r['evals_per_sec']:.4g} | "
                    f"{r['state_steps_per_sec']:.4g} | {r['alg_tflops']:.3f} | {100 * r['frac_fp64_peak']:.1f}% |\n")


if __name__ == "__main__":
    main()
