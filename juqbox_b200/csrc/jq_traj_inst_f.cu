// Part of the instantiation table of the register-resident trajectory kernels (see jq_traj_kernels.cuh): segment sweeps of the
// time-parallel evaluation (jq_seg.cu), fibre layouts.
#include "jq_traj_kernels.cuh"

const Inst kInstF[] = {
    FIBERS(4, 1, 1, 2, 5, 3), FIBERS(4, 1, 1, 2, 0, 0), FIBERS(4, 1, 1, 1, 0, 0),      // single qudits: risk-neutral SWAP 0-2 (n = 4, J = 5)
    FIBERS(6, 1, 1, 2, 3, 4), FIBERS(6, 1, 1, 2, 0, 0),                                // cnot1 (n = 6)
    FIBERS(3, 1, 1, 1, 0, 0), FIBERS(3, 1, 1, 2, 0, 0), FIBERS(5, 1, 1, 2, 0, 0), FIBERS(2, 1, 1, 1, 0, 0),
    FIBERS(4, 2, 1, 1, 0, 0), FIBERS(3, 2, 1, 1, 0, 0), FIBERS(4, 3, 1, 1, 0, 0), FIBERS(3, 3, 1, 1, 0, 0),      // coupled qudits without a tile layout
    // objFuncType 2/3: the gradient sweep (mode 5) with the second adjoint set; the other modes run on the instantiations above
    FIBERSO(3, 2, 1, 1), FIBERSO(4, 2, 1, 1), FIBERSO(4, 1, 1, 1), FIBERSO(4, 1, 1, 2), FIBERSO(6, 1, 1, 2), FIBERSO(4, 3, 1, 1), FIBERSO(2, 1, 1, 1),
};
const int kInstFCount = (int)(sizeof(kInstF) / sizeof(kInstF[0]));
