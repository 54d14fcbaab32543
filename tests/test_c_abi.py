"""The drop-in boundary pinned without Julia: tests/c_abi_harness.c (plain gcc) checks the struct layout the Julia shim
mirrors by hand at compile time, and drives the C ABI the way `ccall` would (column-major arrays, 1-based -> 0-based CSC)."""
import ctypes as C
import os
import subprocess
import struct

import numpy as np
import pytest

from helpers import golden_config, ref_pass

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c_abi_harness.c")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_harness")
    subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror", "-o", exe, SRC, "-ldl"])   # _Static_asserts fire here
    return exe


def test_struct_layout_matches_ctypes_and_library(tmp_path):
    from juqbox_b200 import _lib
    exe = _build(tmp_path)
    vals = [int(x) for x in subprocess.check_output([exe, "--layout"], text=True).split()]
    P, O = _lib.jq_problem, _lib.jq_operator
    want = [C.sizeof(P), C.sizeof(O), P.nsteps.offset, P.T.offset, P.uinit.offset, P.h0.offset, P.hsym.offset, P.solver_tol.offset,
            O.nnz.offset, O.nzval.offset]
    assert vals == want
    assert vals[:2] == [208, 40]          # what julia/JuqboxB200.jl's JqProblem / JqOperator occupy
    lib = _lib.load(build_if_missing=False)
    assert [lib.jq_abi_info(k) for k in range(1, 11)] == vals
    assert lib.jq_abi_info(0) == _lib.ABI_VERSION and lib.jq_abi_info(99) == -1


def _write_problem(path, cfg, sparse, tik0):
    p = cfg.params
    pc = np.asarray(cfg.pcof0, dtype=np.float64)
    with open(path, "wb") as f:
        f.write(struct.pack("<9q", p.Ntot, p.N, p.Ncoupled, p.Nfreq, p.linear_solver.max_iter, p.objFuncType, int(sparse), len(pc), p.nsteps))
        f.write(struct.pack("<2d", p.T, tik0))
        for a in (p.Uinit, p.Utarget_r, p.Utarget_i):
            f.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
        f.write(np.ascontiguousarray(p.wmat_real, dtype=np.float64).tobytes())
        f.write(np.asfortranarray(p.Cfreq[:p.Ncoupled], dtype=np.float64).tobytes(order="F"))
        for a in [p.Hconst] + list(p.Hsym_ops) + list(p.Hanti_ops):
            f.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
        f.write(pc.tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("case,sparse", [("swap02", False), ("swap02", True), ("cnot2", True)])
def test_harness_reproduces_reference_golden(tmp_path, case, sparse):
    """swap02 / cnot2 through gcc + dlopen + the C ABI only, inputs laid out as Julia lays them out: the reference's own
    acceptance rule (test/evalGrad.jl:43-69) on eval_f_par / eval_grad_f_par, served by the fused jq_eval_f_grad."""
    from juqbox_b200 import _lib, tikhonov_grad, tikhonov_pen
    exe = _build(tmp_path)
    cfg, g = golden_config(case)
    prob = str(tmp_path / "problem.bin")
    _write_problem(prob, cfg, sparse, cfg.params.tik0)
    out = subprocess.check_output([exe, _lib.library_path(), prob], text=True).strip().splitlines()
    assert out[0].startswith("traceobjgrad")
    infid, leak, tinf = [float(x) for x in out[0].split()[1:]]
    grad = np.array([float(x) for x in out[1].split()])
    objv = infid + leak + tikhonov_pen(cfg.pcof0, cfg.params)
    ok, dobj, dgrad = ref_pass(objv, grad + tikhonov_grad(cfg.pcof0, cfg.params), g["obj0"], g["grad0"])
    assert ok, (dobj, dgrad)
    # fused callback entry: first call evaluates, second is the cache hit; both equal the golden
    for k, want_ev in ((2, 1), (4, 0)):
        tag, ev, fval = out[k].split()
        assert tag == "eval_f_grad" and int(ev) == want_ev
        gf = np.array([float(x) for x in out[k + 1].split()])
        ok, dobj, dgrad = ref_pass(float(fval), gf, g["obj0"], g["grad0"])
        assert ok, (k, dobj, dgrad)
    assert out[6].startswith("badlen -2 ")


@pytest.mark.parametrize("T,nsteps,nseg", [(50.0, 4472, 74), (550.0, 31325, 129), (300.0, 7937, 148), (10.0, 12, 3), (1.0, 5, 5), (3.0, 7, 1)])
def test_time_segments_follow_the_reference_time_recurrences(T, nsteps, nseg):
    """Host side of the time-parallel evaluation (no device): segment p covers steps [p nsteps / nseg, (p + 1) nsteps / nseg) and starts
    from the value the reference's own recurrences reach -- t = t + dt from 0 (src/evalobjgrad.jl:745), t = t - dt from T (:810, :919)
    -- bit for bit, which is not k dt: the backward recurrence is shifted against the forward one by the rounding of nsteps additions."""
    from juqbox_b200 import _lib
    lib = _lib.load()
    first, last = np.zeros(nseg), np.zeros(nseg)
    dptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    longest = lib.jq_time_segments(T, nsteps, nseg, dptr(first), dptr(last))
    dt = T / nsteps
    tf, t = [0.0], 0.0
    for _ in range(nsteps):
        t = t + dt
        tf.append(t)
    tb, t = [T], T
    for _ in range(nsteps):
        t = t + (-dt)
        tb.append(t)                                  # tb[j]: time after j backward steps = step index nsteps - j
    k0 = [p * nsteps // nseg for p in range(nseg)]
    k1 = [(p + 1) * nsteps // nseg for p in range(nseg)]
    assert k0[0] == 0 and k1[-1] == nsteps and all(a == b for a, b in zip(k1[:-1], k0[1:]))
    assert longest == max(b - a for a, b in zip(k0, k1))
    assert all(first[p] == tf[k0[p]] for p in range(nseg))
    assert all(last[p] == tb[nsteps - k1[p]] for p in range(nseg))
    assert last[-1] == T and first[0] == 0.0
    assert lib.jq_time_segments(T, nsteps, nsteps + 1, dptr(first), dptr(last)) == -1      # more segments than steps
