// Host-side tracker object behind the opaque `dvfe_tracker` handle of include/dvfe.h.
#pragma once
#include <vector>

#include "kernels.cuh"
#include "state.cuh"

struct GfttScratch {
    int n_jobs, w, h, mask_pitch, cand_cap, n_cells;
    uint8_t* mask;
    unsigned long long *cand, *cand2, *cand3;
    uint8_t* state;
    int* cell_count;
    int* counters;
};

int alloc_point_sets(PointSetArrays* S, int n_sets, int cap);
void free_point_sets(PointSetArrays* S);
int alloc_gftt_scratch(GfttScratch* sc, int n_jobs, int w, int h, float min_dist);
void free_gftt_scratch(GfttScratch* sc);
void gftt_job_bind_scratch(GfttJob* J, const GfttScratch& sc, int j);
int gftt_cells(int w, int h, float min_dist);

struct InstanceState;   // instances.cu

struct dvfe_tracker {
    dvfe_config cfg{};
    int B = 0, W = 0, H = 0, cap = 0;
    cudaStream_t st = nullptr;
    PyrDesc desc{};
    CamParams cam0{}, cam1{};
    uint8_t* pyr[3] = {nullptr, nullptr, nullptr};   // left (even frames), left (odd frames), right
    int cur = 0;                                     // pyr[cur] receives the current left image
    long frames = 0;
    bool last_has_right = false;                     // the last uploaded frame had a right image
    PointSetArrays bg{};                             // background point sets, one per stream
    uint32_t* d_next_id = nullptr;                   // [B] InstFeat::global_id_count per stream
    double* d_dt = nullptr;
    double* h_dt = nullptr;
    std::vector<double> prev_time;
    dvfe_obs* d_obs = nullptr;
    dvfe_obs* h_obs = nullptr;
    int* d_nobs = nullptr;
    int* h_nobs = nullptr;
    uint8_t *d_region = nullptr, *d_region_tmp = nullptr, *d_inv_in = nullptr;
    int* d_exist = nullptr;
    int* h_exist = nullptr;
    GfttScratch gsc{};
    LkGroup* d_groups[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    GfttJob* d_jobs[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    InstanceState* inst = nullptr;
    bool own_stream = true;

    // per-stage device timers
    enum { ST_PYRAMID, ST_LK_TEMPORAL, ST_COMPACT, ST_GFTT, ST_LEFT_POST, ST_LK_STEREO, ST_PACK, ST_D2H, ST_COUNT };
    bool prof = false;
    cudaEvent_t ev[ST_COUNT + 1] = {};
    double prof_ms[ST_COUNT] = {};
    long prof_steps = 0;
    void mark(int i) { if (prof) cudaEventRecord(ev[i], st); }

    int init();
    int upload_in_place(const uint8_t* left, const uint8_t* right, size_t stream_stride, int pitch);
    int step_device(const uint8_t* d_left, const uint8_t* d_right, size_t stream_stride, int pitch, const double* time0,
                    bool semantic, bool level0_in_place = false, bool has_right = false);
    int init_instances();
    void free_instances();
};
