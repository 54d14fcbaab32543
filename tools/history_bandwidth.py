"""HBM write rate of the forward propagator with state history (jq_eval_forward) next to the same
forward sweep without history.      python tools/history_bandwidth.py      (needs a GPU)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq                                    # noqa: E402
from juqbox_b200 import configs                             # noqa: E402

cfg = configs.example("cnot2")
p = cfg.params
print("| kernel | trajectories | saveEvery | kernel ms | forward sweep without history, ms | history GB | GB/s written |")
print("|---|---|---|---|---|---|---|")
for nb, se in ((64, 1), (256, 1), (1024, 4), (2048, 2), (4096, 4)):
    pc = configs.synthetic_pcof(cfg, nb)
    wa = jq.Working_Arrays(p, cfg.nCoeff)
    for k in (0, 1):
        if k == 1 and nb > 64:
            continue
        wa.set_kernel(k)
        wa.evaluate(pc, evaladjoint=False)
        wa.evaluate(pc, evaladjoint=False)
        ms0 = wa.last_kernel_ms
        wa.forward_history(pc, save_every=se)
        hist, _, _ = wa.forward_history(pc, save_every=se)
        nbytes = hist.size * 16
        print(f"| {wa.last_kernel} | {nb} | {se} | {wa.last_kernel_ms:.1f} | {ms0:.1f} | {nbytes / 1e9:.2f} | "
              f"{nbytes / 1e9 / (wa.last_kernel_ms * 1e-3):.0f} |", flush=True)
        del hist
    wa.close()
