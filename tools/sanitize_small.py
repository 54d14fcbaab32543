"""Small runs of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import juqbox_b200 as jq
from juqbox_b200 import configs

for name, T, nsteps in (("risk_neutral", 3.0, 40), ("cnot2", 2.0, 40), ("cnot3", 2.0, 24), ("cnot1", 2.0, 40), ("rabi", 10.0, 12)):
    cfg = configs.example(name)
    cfg.params.T, cfg.params.nsteps = T, nsteps
    pc = configs.synthetic_pcof(cfg, 5) * 20
    sh = configs.noise_shift(cfg.params.Ntot, [-0.03, 0.05]) if name == "risk_neutral" else None
    res = {}
    for obj in (1, 3):
        cfg.params.objFuncType = obj
        wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
        for k in (1, 2, 3):
            try:
                wa.set_kernel(k)
            except Exception:
                continue
            r = wa.evaluate(pc, sh)
            res[(obj, k)] = r["grad"]
        wa.close()
    ref = res[(1, 1)]
    for key, g in res.items():
        err = np.linalg.norm(g - ref) / np.linalg.norm(ref)
        print(name, key, "rel diff vs generic", err)
        assert err < 1e-10
# the other kernels: state history from every trajectory kernel, weighted sums over many samples, control read-out
cfg = configs.example("cnot2")
cfg.params.T, cfg.params.nsteps = 2.0, 40
pc = configs.synthetic_pcof(cfg, 3) * 20
wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
hists = {}
for k in (1, 2, 3):
    wa.set_kernel(k)
    hists[k] = wa.forward_history(pc, save_every=4)[0]
assert np.abs(hists[3] - hists[1]).max() < 1e-12 and np.abs(hists[2] - hists[1]).max() < 1e-12
wa.set_kernel(0)
p, q = wa.controls(pc[0], np.linspace(0, cfg.params.T, 77))
wa.close()
cfg = configs.example("risk_neutral")
cfg.params.T, cfg.params.nsteps = 3.0, 40
sh = configs.noise_shift(cfg.params.Ntot, np.linspace(-0.05, 0.05, 130))
w = np.full(130, 1.0 / 130)
wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
pc = configs.synthetic_pcof(cfg, 2) * 20
tot = wa.evaluate(pc, sh, w)
per = wa.evaluate(pc, sh)
assert np.allclose(tot["grad"], (per["grad"] * w[None, :, None]).sum(1), rtol=1e-12, atol=1e-16)
wa.close()
print("sanitize_small OK")
