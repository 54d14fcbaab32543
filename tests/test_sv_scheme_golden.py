"""The reference's integrator known-answer test (test/test-stormer-verlet.jl:137-182): final-time errors of the
Stormer-Verlet scheme on four analytic 2x2 problems x three CFL numbers vs err-mat-ref.jld2, max abs diff <= 1e-13."""
import numpy as np

from helpers import load_golden
from oracle.sv_scheme import timesteptest


def test_stormer_verlet_convergence_errors_match_reference():
    g = load_golden("err-mat")
    assert g["hdf5_dims"] == [4, 2, 3]                    # Julia err_mat is 3 x 2 x 4 (CFL, {cg, ce}, testcase)
    ref = np.array(g["data"]).reshape(4, 2, 3)            # [testcase][cg/ce][cfl] = Julia column-major order
    cfls = 10.0 ** np.arange(-1.0, -2.01, -0.5)
    got = np.zeros_like(ref)
    for j in range(4):
        for i, cfl in enumerate(cfls):
            got[j, 0, i], got[j, 1, i] = timesteptest(cfl, j)
    assert np.max(np.abs(got - ref)) <= 1e-13, np.max(np.abs(got - ref))
    # second-order convergence: error drops ~10x per half decade of CFL... i.e. ~100x per decade
    assert np.all(got[:, :, 2] < got[:, :, 0] * 0.02 + 1e-12)
