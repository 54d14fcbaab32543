// placeholder until the warp-slot kernel lands
#include "jq_common.h"
#include <cstdio>
SlotPlan *jq_slot_plan_create(const DevProblem &, const HostOps &, char *err, size_t errlen) {
    snprintf(err, errlen, "warp-slot kernel not built yet");
    return nullptr;
}
void jq_slot_plan_destroy(SlotPlan *) {}
cudaError_t jq_slot_launch(SlotPlan *, const DevProblem &, const LaunchArgs &, cudaStream_t, int *, int *, size_t *, int *) {
    return cudaErrorNotSupported;
}
