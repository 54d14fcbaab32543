"""One launch of the default trajectory kernel for `ncu --set full` (run under ncu; never a bench number)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq
from juqbox_b200 import configs
name = sys.argv[1] if len(sys.argv) > 1 else "cnot2"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
cfg = configs.example(name)
pc = configs.synthetic_pcof(cfg, nb)
wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
r = wa.evaluate(pc, shifts)
print(name, nb, "kernel", wa.last_kernel, "ms", wa.last_kernel_ms)
