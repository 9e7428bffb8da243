// Host-side tracker object behind the opaque `dvfe_tracker` handle of include/dvfe.h.
#pragma once
#include <map>
#include <tuple>
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
    // A tracker is either a leaf (owns device state for its B streams) or a container of G leaf trackers
    // ("stream groups", dvfe_config::n_groups) that splits every call by stream range.
    std::vector<dvfe_tracker*> groups;
    std::vector<int> group_first;                    // first stream of each group (size G + 1)
    dvfe_config cfg{};
    int B = 0, W = 0, H = 0, cap = 0;
    cudaStream_t st = nullptr;                       // compute stream (caller-replaceable)
    cudaStream_t cs = nullptr;                       // upload stream: H2D of step k+1 overlaps the kernels of step k
    cudaStream_t ds = nullptr;                       // download stream: D2H of step k overlaps the kernels of step k+1
    cudaStream_t rs = nullptr;                       // right-image pyramid stream: fills the SMs the selection kernel leaves idle
    bool own_stream = true;
    PyrDesc desc{};
    CamParams cam0{}, cam1{};
    // Pyramids: 3 left slots (current / previous / being uploaded) and 2 right slots.  Step k uses
    // left[k % 3] (previous = left[(k + 2) % 3]) and right[k % 2]; "phase" = k % 6 selects the descriptor set.
    uint8_t* pyrL[3] = {nullptr, nullptr, nullptr};
    uint8_t* pyrR[2] = {nullptr, nullptr};
    long frames = 0;                                 // steps submitted
    long completed = 0;                              // steps whose outputs have been waited for
    bool last_has_right = false;                     // the last submitted frame had a right image
    PointSetArrays bg{};                             // background point sets, one per stream
    uint32_t* d_next_id = nullptr;                   // [B] InstFeat::global_id_count per stream
    double* d_dt = nullptr;
    double* h_dt[2] = {nullptr, nullptr};
    std::vector<double> prev_time;
    dvfe_obs* d_obs[2] = {nullptr, nullptr};         // one per in-flight step
    dvfe_obs* h_obs[2] = {nullptr, nullptr};         // pinned, one per in-flight step
    int* d_nobs[2] = {nullptr, nullptr};
    int* h_nobs[2] = {nullptr, nullptr};             // [B + 1]: counts, then the device error word of that step
    int* d_err = nullptr;                            // [2]: one capacity-overflow flag per in-flight step
    int out_slot = 0;                                // which h_obs holds the newest completed step
    cudaEvent_t ev_up[2] = {}, ev_packed[2] = {}, ev_done[2] = {}, ev_resp[2] = {}, ev_rpyr[2] = {}, ev_r0[2] = {}, ev_begin[2] = {};
    uint8_t *d_region = nullptr, *d_region_tmp = nullptr;
    uint8_t* d_inv_in[2] = {nullptr, nullptr};       // uploaded inv_merge_mask, one per in-flight step
    int* d_exist = nullptr;
    int* h_exist[2] = {nullptr, nullptr};            // pinned, one per in-flight step
    cudaEvent_t ev_inst[2] = {};                     // instance records of the step are on the host
    bool inst_pending[2] = {false, false};           // the step carried a deferred InstsTrack (dvfe_track_dynamic_async)
    GfttScratch gsc{};
    LkGroup* d_groups[6][3] = {};                    // [phase][temporal raw | temporal semantic | stereo]
    GfttJob* d_jobs[6][3] = {};                      // [phase][raw | semantic | semantic, cv::cuda detector threshold]
    InstanceState* inst = nullptr;
    unsigned* d_tcache = nullptr;                    // LK template cache [B][cap][DVFE_MAX_PYR_LEVELS][LK_TCACHE_WORDS] (stereo only)
    bool tcache_valid = false;                       // the last step ran the stereo LK on the points `bg` holds now
    unsigned* d_tcache_bwd = nullptr;                // backward templates of the temporal call, same layout (stereo only)
    int* d_old_idx = nullptr;                        // [B][cap] index of each point in the temporal call of the step, -1 = new
    // backward LK maxLevel / forward-backward threshold per call site [DVFE_LK_*]: the CPU FeatureTrackByLK pair
    // (feature_utils.cpp:51,57: 1, 0.5) everywhere except TrackSemanticImage's right image, which the reference tracks with
    // the cv::cuda call pattern (TrackRightGPU, background_tracker.cpp:801: 3 levels, 1.0 px)
    int lk_back_level[4] = {1, 1, 1, 3};
    double lk_fb_thresh[4] = {0.5, 0.5, 0.5, 1.0};

    // The compute part of a frame step is a fixed launch sequence for a given buffer phase, input location and mode: it is
    // captured once into a CUDA graph and replayed (one launch per step and stream group instead of ~20).
    typedef std::tuple<int, unsigned> StepKey;       // buffer phase, mode flags (the caller's image pointers stay outside)
    struct StepGraph { cudaGraphExec_t exec = nullptr; unsigned n_kernels = 0; };
    std::map<StepKey, StepGraph> step_graphs;
    bool detect_cuda = false;                        // dvfe_set_detect_mode: semantic-path detection with the cv::cuda detector's threshold
    bool use_graphs = true;
    bool use_reuse = true;                           // DVFE_REUSE=0: the stereo call rebuilds every forward template (A/B, debugging)
    void drop_graphs();
    int capture_step(int ph, unsigned flags, long k);
    int enqueue_compute(const uint8_t* d_left, const uint8_t* d_right, size_t stream_stride, int pitch, bool semantic,
                        int level0_mode, bool stereo_now, long k, bool with_marks);

    // per-stage device timers (one event set per in-flight step)
    enum { ST_PYRAMID, ST_LK_TEMPORAL, ST_COMPACT, ST_GFTT_MASK, ST_GFTT_DISCS, ST_GFTT_RESPONSE, ST_GFTT_SELECT, ST_LEFT_POST,
           ST_LK_STEREO, ST_PACK, ST_D2H, ST_COUNT };
    bool prof = false;
    bool prof_step[2] = {false, false};
    cudaEvent_t ev[2][ST_COUNT + 1] = {};
    double prof_ms[ST_COUNT] = {};
    long prof_steps = 0;

    // TMA descriptors (CUtensorMap, 128 B each) over level 0 of the three left pyramid sets: the corner response loads its
    // tiles with cp.async.bulk.tensor; use_tma is false when the driver cannot encode them (then cp.async is used)
    alignas(64) unsigned char tmapL[3][128] = {};
    bool use_tma = false;
    uint8_t* left_slot(long k) const { return pyrL[k % 3]; }
    uint8_t* right_slot(long k) const { return pyrR[k % 2]; }

    int init();
    int init_instances();
    void free_instances();
    uint8_t* d_stage[2] = {nullptr, nullptr};       // dense upload staging [2 cameras][B][H*W], one per in-flight step
    bool staged_upload = false;                      // rows that are not a multiple of 64 B make pitched DMA slow
    // frame ingest (dvfe_set_input / dvfe_set_undistort_maps): BGR and/or remapped input goes through a device staging
    // buffer and the ingest kernel, which writes level 0 in place
    int in_ch = 1;
    short* d_map1[2] = {nullptr, nullptr};
    unsigned short* d_map2[2] = {nullptr, nullptr};
    uint8_t* d_raw[2] = {nullptr, nullptr};          // [2 cameras][B][H*W*in_ch] per in-flight step
    bool prep_active() const { return in_ch != 1 || d_map1[0] || d_map1[1]; }
    int ensure_raw();
    int ingest(const uint8_t* d_src, size_t stream_stride, int pitch, int cam, cudaStream_t s);
    int upload_prepared(const uint8_t* left, const uint8_t* right, size_t stream_stride, int pitch);
    int upload_in_place(const uint8_t* left, const uint8_t* right, size_t stream_stride, int pitch);
    int upload_staged(const uint8_t* left, const uint8_t* right, size_t stream_stride, int pitch);
    // enqueue one frame step (no host synchronisation); at most two steps are in flight
    int submit(const uint8_t* d_left, const uint8_t* d_right, size_t stream_stride, int pitch, const double* time0,
               bool semantic, bool level0_in_place, bool has_right);
    // TrackSemanticImage of one frame, enqueued only (uploads on the copy stream, mask erosion + the frame step)
    // flags: DVFE_DYN_DEVICE_INPUT (the three pointers are device memory, nothing is uploaded), DVFE_DYN_LABELS (the mask is a
    // label image: the region is everything no instance bit is set on)
    int semantic_submit(const uint8_t* left, const uint8_t* right, const uint8_t* mask, size_t stream_stride,
                        int pitch, const int* exist_inst, const double* time0, unsigned flags = 0);
    uint8_t* d_labels[2] = {nullptr, nullptr};       // uploaded label images, one per in-flight step (allocated on first use)
    const uint8_t* lab_ptr[2] = {nullptr, nullptr};  // the label images of the in-flight steps (device), for the ROI masks
    size_t lab_stride[2] = {0, 0};
    int lab_pitch[2] = {0, 0};
    int finish_instances(int par);                   // instances.cu: host side of a deferred InstsTrack
    int wait_one();                                  // oldest in-flight step -> outputs readable
    int wait_all();
};
