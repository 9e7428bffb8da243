// Per-instance tracking (InstsFeatManager::InstsTrack / Output) — see instances section of DESIGN.md.
#include "kernels.cuh"
#include "state.cuh"
#include "tracker.h"

void dvfe_tracker::free_instances() {}

extern "C" int dvfe_insts_track(dvfe_tracker* t, int stream, const dvfe_inst_in* insts, int n_insts, double time0) {
    (void)t; (void)stream; (void)insts; (void)n_insts; (void)time0;
    dvfe_set_error("dvfe_insts_track: not built yet");
    return DVFE_ERR_INVALID;
}
extern "C" int dvfe_insts_output(dvfe_tracker* t, int stream, dvfe_inst_obs* out, int cap, int* n_out) {
    (void)t; (void)stream; (void)out; (void)cap;
    if (n_out) *n_out = 0;
    dvfe_set_error("dvfe_insts_output: not built yet");
    return DVFE_ERR_INVALID;
}
