#!/usr/bin/env python3
"""Development: single-trajectory kernel time, objective-only (forward sweep) vs objective + gradient, per kernel id."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import juqbox_b200 as jq
from juqbox_b200 import configs
name = sys.argv[1]
kernels = [int(k) for k in sys.argv[2].split(",")]
cfg = configs.example(name)
dev = torch.device("cuda", 0)
shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
sh = torch.from_numpy(shifts).to(dev) if shifts is not None else None
pc = torch.from_numpy(configs.synthetic_pcof(cfg, 1)).to(dev)
for k in kernels:
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    try:
        wa.set_kernel(k)
    except Exception as e:
        print(name, k, "unavailable"); continue
    res = []
    for adj in (False, True):
        for _ in range(3):
            wa.evaluate_device(pc, sh, None, adj)
            torch.cuda.synchronize()
        res.append(wa.last_kernel_ms)
    print(f"{name} kernel {k}: forward only {res[0]:.3f} ms, forward + backward {res[1]:.3f} ms (backward {res[1] - res[0]:.3f})", flush=True)
    wa.close()
