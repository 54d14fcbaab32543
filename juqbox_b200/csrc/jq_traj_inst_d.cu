// Part of the instantiation table of the register-resident trajectory kernels (see jq_traj_kernels.cuh): tile layout.
#include "jq_traj_kernels.cuh"

const Inst kInstD[] = {
    TILEJ(2, 2, 1, 4, 16),                 // cnot2 example shape: 4 x 4 levels, 2 x 2 tiles, J = 4, 16-lane groups (two trajectories per warp)
    TILEJ(2, 2, 1, 0, 0),                  // 4 x 4 levels, run-time J and group size
    TILEJW(3, 3, 1, 3, 32, 8, 1, 0),       // cnot3 example shape: 4 x 4 x 4 levels, 2 x 2 x 2 tiles (one warp per trajectory), J = 3, 8 warps per CTA
    TILEJW(3, 3, 1, 0, 0, 8, 1, 0),        // 4 x 4 x 4 levels, run-time J and group size
    TILEJ(2, 0, 1, 4, 32), TILEJ(2, 0, 1, 0, 0),   // latency layout for small batches: one element per lane, 64 lanes per 4 x 4 x 4-column trajectory
    TILEJ(3, 1, 1, 3, 32), TILEJ(3, 1, 1, 0, 0),   // 4 x 4 x 4 levels: first subsystem in halves, 2 elements per lane, one warp per column
    TILEP(2, 0, 1, 4, 32, 2), TILEP(2, 0, 1, 0, 0, 2),       // the same with pipelined state / adjoint roles (kernel id 5)
    TILEP2(3, 1, 1, 3, 32, 4), TILEP2(3, 1, 1, 0, 0, 4),     // three subsystems: state + adjoint share a role (register budget)
    FIBERP(4, 1, 1, 2, 5, 3, 1), FIBERP(4, 1, 1, 2, 0, 0, 1), FIBERP(4, 1, 1, 1, 0, 0, 1),       // single qudits: risk-neutral SWAP 0-2 (J = 5), n = 4
    FIBERP(6, 1, 1, 2, 3, 4, 1), FIBERP(6, 1, 1, 2, 0, 0, 1), FIBERP(2, 1, 1, 1, 0, 0, 1),       // cnot1 (n = 6), rabi (n = 2)
    TILEJ(3, 2, 1, 0, 0),                  // 2 x 2 tiles x remote third subsystem (JQ_TILE_NT=2; 4 elements per lane, no spills, slower)
};
const int kInstDCount = (int)(sizeof(kInstD) / sizeof(kInstD[0]));
