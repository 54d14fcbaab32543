"""Shared test helpers: the reference's acceptance criterion (test/evalGrad.jl:43-69) and config loading."""
import json
import os

import numpy as np

from juqbox_b200 import configs, tikhonov_grad, tikhonov_pen

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-10, 1e-14   # test/evalGrad.jl:4-5


def load_golden(case):
    with open(os.path.join(GOLDEN_DIR, f"{case}.json")) as f:
        return json.load(f)


def golden_config(case):
    g = load_golden(case)
    return configs.test_case(case, g.get("pcof0")), g


def ref_pass(objv, grad, objRef, gradRef, rtol=RTOL, atol=ATOL):
    """Exactly the pass/fail rule of evalObjGrad (test/evalGrad.jl:43-69). Returns (ok, objDiff, relGradErr)."""
    objv, objRef = np.atleast_1d(objv).astype(float), np.atleast_1d(objRef).astype(float)
    grad, gradRef = np.asarray(grad, float), np.asarray(gradRef, float)
    objDiff = abs(objv[0] - objRef[0]) if len(objv) == 1 else np.linalg.norm(objv - objRef)
    refNorm, aNorm = np.linalg.norm(gradRef), np.linalg.norm(grad - gradRef)
    nrm = abs(objRef[0]) if len(objRef) == 1 else np.linalg.norm(objRef)  # reference uses abs(objvRef) on the vector
    pass1 = objDiff < atol or (nrm >= atol and objDiff / nrm < rtol)
    pass2 = aNorm < atol or (refNorm >= atol and aNorm / refNorm < rtol)
    return bool(pass1 and pass2), float(objDiff), float(aNorm / refNorm)


def with_tikhonov(cfg, res, b=0, s=0):
    """Turn raw traceobjgrad outputs into what eval_f_par / eval_grad_f_par (+ eval_jac_g_par) return
    (src/ipopt_interface.jl:77-148; test/evalGrad.jl:14-26)."""
    p, pc = cfg.params, cfg.pcof0
    tp, tg = tikhonov_pen(pc, p), tikhonov_grad(pc, p)
    if p.objFuncType == 1:
        return res["infid"][b, s] + res["leak"][b, s] + tp, res["grad"][b, s] + tg
    objv = np.array([res["infid"][b, s] + tp, res["leak"][b, s]])
    grad = np.concatenate([res["infidgrad"][b, s] + tg, res["leakgrad"][b, s]])
    return objv, grad
