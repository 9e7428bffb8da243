// Seam-level operators of include/dvfe.h: each runs the same kernels as the frame step on caller-provided
// HOST buffers (upload -> kernels -> download).  They exist for unit parity against the oracle and for
// integrators who want a single stage; they allocate per call and are not the fast path.
#include <string.h>

#include <vector>

#include "kernels.cuh"
#include "state.cuh"
#include "tracker.h"

#define DVFE_CHECK(call)                 \
    do {                                 \
        int rc__ = (call);               \
        if (rc__ != DVFE_OK) return rc__; \
    } while (0)

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        DVFE_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        DVFE_CUDA(cudaMemset(p, 0, bytes ? bytes : 1));
        return DVFE_OK;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

int ensure_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        dvfe_set_error("no CUDA device available (%s): libdvfe has no CPU fallback", cudaGetErrorString(e));
        return DVFE_ERR_NO_DEVICE;
    }
    return DVFE_OK;
}

int upload_image(DevBuf& d, const uint8_t* img, int w, int h, int pitch) {
    DVFE_CHECK(d.alloc((size_t)w * h));
    DVFE_CUDA(cudaMemcpy2D(d.p, w, img, pitch, w, h, cudaMemcpyHostToDevice));
    return DVFE_OK;
}
}  // namespace

extern "C" int dvfe_op_build_pyramid(const uint8_t* img, int w, int h, int pitch, int max_level, uint8_t* const* out_levels,
                                     int* level_w, int* level_h, int* n_levels) {
    if (!img || w < 1 || h < 1 || pitch < w || max_level < 0 || max_level >= DVFE_MAX_PYR_LEVELS || !n_levels) {
        dvfe_set_error("op_build_pyramid: bad argument");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(ensure_device());
    const PyrDesc desc = make_pyr_desc(w, h, max_level);
    DevBuf d_img, d_pyr, d_out;
    DVFE_CHECK(upload_image(d_img, img, w, h, pitch));
    DVFE_CHECK(d_pyr.alloc(desc.bytes));
    DVFE_CHECK(d_out.alloc((size_t)w * h));
    PyrImgSet set{};
    set.src[0] = d_img.as<uint8_t>(); set.dst[0] = d_pyr.as<uint8_t>();
    set.per_set = 1;
    DVFE_CHECK(launch_build_pyramids(set, 1, desc, w, 0));
    *n_levels = desc.n_levels;
    for (int l = 0; l < desc.n_levels; l++) {
        if (level_w) level_w[l] = desc.lv[l].w;
        if (level_h) level_h[l] = desc.lv[l].h;
        if (out_levels && out_levels[l]) {
            DVFE_CHECK(launch_pyr_extract(d_pyr.as<uint8_t>(), desc.lv[l], d_out.as<uint8_t>(), 0));
            DVFE_CUDA(cudaMemcpy(out_levels[l], d_out.p, (size_t)desc.lv[l].w * desc.lv[l].h, cudaMemcpyDeviceToHost));
        }
    }
    DVFE_CUDA(cudaDeviceSynchronize());
    return DVFE_OK;
}

extern "C" int dvfe_op_build_pyramid_bordered(const uint8_t* img, int w, int h, int pitch, int max_level, int level, int border,
                                              uint8_t* out) {
    if (!img || !out || w < 1 || h < 1 || pitch < w || max_level < 0 || max_level >= DVFE_MAX_PYR_LEVELS || level < 0 || border < 0 ||
        border > DVFE_WIN) {
        dvfe_set_error("op_build_pyramid_bordered: bad argument (border <= 21 = winSize)");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(ensure_device());
    const PyrDesc desc = make_pyr_desc(w, h, max_level);
    if (level >= desc.n_levels) { dvfe_set_error("op_build_pyramid_bordered: level %d of %d", level, desc.n_levels); return DVFE_ERR_INVALID; }
    DevBuf d_img, d_pyr, d_out;
    DVFE_CHECK(upload_image(d_img, img, w, h, pitch));
    DVFE_CHECK(d_pyr.alloc(desc.bytes));
    const PyrLevel& L = desc.lv[level];
    const size_t nout = (size_t)(L.w + 2 * border) * (L.h + 2 * border);
    DVFE_CHECK(d_out.alloc(nout));
    PyrImgSet set{};
    set.src[0] = d_img.as<uint8_t>(); set.dst[0] = d_pyr.as<uint8_t>();
    set.per_set = 1;
    DVFE_CHECK(launch_build_pyramids(set, 1, desc, w, 0));
    DVFE_CHECK(launch_pyr_extract_bordered(d_pyr.as<uint8_t>(), L, border, d_out.as<uint8_t>(), 0));
    DVFE_CUDA(cudaDeviceSynchronize());
    DVFE_CUDA(cudaMemcpy(out, d_out.p, nout, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}

extern "C" int dvfe_op_lk(const uint8_t* img1, const uint8_t* img2, int w, int h, int pitch, const float* pts1, int n,
                          int flow_back, int max_level, const uint8_t* mask, int mask_pitch, float* pts2_out,
                          uint8_t* status_out, float* rev_out) {
    if (!img1 || !img2 || !pts1 || n <= 0 || w < 1 || h < 1 || pitch < w || max_level < 0 ||
        max_level >= DVFE_MAX_PYR_LEVELS) {
        // FeatureTrackByLK throws on empty input (front_end/feature_utils.cpp:39-41)
        dvfe_set_error("FeatureTrackByLK() input wrong, received at least one of parameter are empty");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(ensure_device());
    // the backward call always asks for maxLevel 1 (feature_utils.cpp:51), whatever the forward maxLevel is
    const PyrDesc desc = make_pyr_desc(w, h, max_level > 1 ? max_level : 1);
    DevBuf d_img, d_pyr, d_p1, d_p2, d_rev, d_st, d_n, d_mask, d_grp;
    DVFE_CHECK(d_img.alloc((size_t)2 * w * h));
    DVFE_CUDA(cudaMemcpy2D(d_img.p, w, img1, pitch, w, h, cudaMemcpyHostToDevice));
    DVFE_CUDA(cudaMemcpy2D(d_img.as<uint8_t>() + (size_t)w * h, w, img2, pitch, w, h, cudaMemcpyHostToDevice));
    DVFE_CHECK(d_pyr.alloc((size_t)2 * desc.bytes));
    DVFE_CHECK(d_p1.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_p2.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_rev.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_st.alloc(n));
    DVFE_CHECK(d_n.alloc(sizeof(int)));
    DVFE_CUDA(cudaMemcpy(d_p1.p, pts1, sizeof(float2) * n, cudaMemcpyHostToDevice));
    DVFE_CUDA(cudaMemcpy(d_n.p, &n, sizeof(int), cudaMemcpyHostToDevice));
    if (mask) {
        DVFE_CHECK(d_mask.alloc((size_t)w * h));
        DVFE_CUDA(cudaMemcpy2D(d_mask.p, w, mask, mask_pitch, w, h, cudaMemcpyHostToDevice));
    }
    PyrImgSet set{};
    set.src[0] = d_img.as<uint8_t>(); set.src[1] = d_img.as<uint8_t>() + (size_t)w * h;
    set.dst[0] = d_pyr.as<uint8_t>(); set.dst[1] = d_pyr.as<uint8_t>() + desc.bytes;
    set.per_set = 1;
    DVFE_CHECK(launch_build_pyramids(set, 2, desc, w, 0));
    LkGroup G;
    memset(&G, 0, sizeof(G));
    G.pyrA = set.dst[0]; G.pyrB = set.dst[1]; G.desc = desc;
    G.ptsA = d_p1.as<float2>(); G.ptsB = d_p2.as<float2>(); G.rev = d_rev.as<float2>();
    G.status = d_st.as<uint8_t>(); G.n = d_n.as<int>();
    G.mask = mask ? d_mask.as<uint8_t>() : nullptr; G.mask_pitch = w;
    DVFE_CHECK(d_grp.alloc(sizeof(LkGroup)));
    DVFE_CUDA(cudaMemcpy(d_grp.p, &G, sizeof(G), cudaMemcpyHostToDevice));
    DVFE_CHECK(launch_lk(d_grp.as<LkGroup>(), 1, n, max_level, flow_back, 0));
    DVFE_CUDA(cudaDeviceSynchronize());
    if (pts2_out) DVFE_CUDA(cudaMemcpy(pts2_out, d_p2.p, sizeof(float2) * n, cudaMemcpyDeviceToHost));
    if (status_out) DVFE_CUDA(cudaMemcpy(status_out, d_st.p, n, cudaMemcpyDeviceToHost));
    if (rev_out) DVFE_CUDA(cudaMemcpy(rev_out, d_rev.p, sizeof(float2) * n, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}

extern "C" int dvfe_op_min_eigen_val(const uint8_t* img, int w, int h, int pitch, float* eig_out) {
    if (!img || !eig_out || w < 3 || h < 3 || pitch < w) { dvfe_set_error("op_min_eigen_val: bad argument"); return DVFE_ERR_INVALID; }
    DVFE_CHECK(ensure_device());
    DevBuf d_img, d_eig;
    DVFE_CHECK(upload_image(d_img, img, w, h, pitch));
    DVFE_CHECK(d_eig.alloc(sizeof(float) * (size_t)w * h));
    DVFE_CHECK(launch_min_eigen_val(d_img.as<uint8_t>(), w, w, h, d_eig.as<float>(), 0));
    DVFE_CUDA(cudaDeviceSynchronize());
    DVFE_CUDA(cudaMemcpy(eig_out, d_eig.p, sizeof(float) * (size_t)w * h, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}

static int op_good_features(const uint8_t* img, int w, int h, int pitch, const float* eig, const uint8_t* mask,
                            int mask_pitch, int max_corners, double quality, double min_dist, float* corners_out,
                            int* n_out, int* n_candidates_out, int max_unmasked) {
    if ((!img && !eig) || w < 3 || h < 3 || max_corners < 1 || max_corners > 2048 || !corners_out || !n_out) {
        dvfe_set_error("op_good_features: bad argument (max_corners must be in 1..2048)");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(ensure_device());
    DevBuf d_img, d_eig_in, d_mask, d_pts, d_n, d_job;
    GfttScratch sc{};
    if (img) DVFE_CHECK(upload_image(d_img, img, w, h, pitch));
    if (eig) {
        DVFE_CHECK(d_eig_in.alloc(sizeof(float) * (size_t)w * h));
        DVFE_CUDA(cudaMemcpy(d_eig_in.p, eig, sizeof(float) * (size_t)w * h, cudaMemcpyHostToDevice));
    }
    if (mask) {
        DVFE_CHECK(d_mask.alloc((size_t)w * h));
        DVFE_CUDA(cudaMemcpy2D(d_mask.p, w, mask, mask_pitch, w, h, cudaMemcpyHostToDevice));
    }
    DVFE_CHECK(d_pts.alloc(sizeof(float2) * max_corners));
    DVFE_CHECK(d_n.alloc(sizeof(int)));
    int rc = alloc_gftt_scratch(&sc, 1, w, h, (float)min_dist);
    if (rc == DVFE_OK) {
        GfttJob J;
        memset(&J, 0, sizeof(J));
        J.img = d_img.as<uint8_t>(); J.img_pitch = w; J.w = w; J.h = h;
        J.region_mask = mask ? d_mask.as<uint8_t>() : nullptr; J.region_pitch = w;
        gftt_job_bind_scratch(&J, sc, 0);
        J.eig_in = eig ? d_eig_in.as<float>() : nullptr;
        J.pts = d_pts.as<float2>(); J.n = d_n.as<int>();
        J.max_cnt = max_corners; J.min_needed = 1; J.disc_radius = 0;
        J.min_dist = (float)min_dist; J.quality = quality;
        J.max_unmasked = max_unmasked;
        rc = d_job.alloc(sizeof(GfttJob));
        if (rc == DVFE_OK && cudaMemcpy(d_job.p, &J, sizeof(J), cudaMemcpyHostToDevice) != cudaSuccess) rc = DVFE_ERR_CUDA;
        if (rc == DVFE_OK) rc = launch_gftt(d_job.as<GfttJob>(), &J, 1, w, h, 0, 0);
        if (rc == DVFE_OK && cudaDeviceSynchronize() != cudaSuccess) {
            dvfe_set_error("op_good_features: %s", cudaGetErrorString(cudaGetLastError()));
            rc = DVFE_ERR_CUDA;
        }
        if (rc == DVFE_OK) {
            int counters[8];
            cudaMemcpy(counters, sc.counters, sizeof(counters), cudaMemcpyDeviceToHost);
            cudaMemcpy(n_out, d_n.p, sizeof(int), cudaMemcpyDeviceToHost);
            if (*n_out > 0) cudaMemcpy(corners_out, d_pts.p, sizeof(float2) * (*n_out), cudaMemcpyDeviceToHost);
            if (n_candidates_out) *n_candidates_out = counters[3];
            if (counters[7]) { dvfe_set_error("op_good_features: candidate buffer overflow"); rc = DVFE_ERR_CAPACITY; }
        }
    }
    free_gftt_scratch(&sc);
    return rc;
}

extern "C" int dvfe_op_good_features(const uint8_t* img, int w, int h, int pitch, const float* eig, const uint8_t* mask,
                                     int mask_pitch, int max_corners, double quality, double min_dist, float* corners_out,
                                     int* n_out, int* n_candidates_out) {
    return op_good_features(img, w, h, pitch, eig, mask, mask_pitch, max_corners, quality, min_dist, corners_out, n_out, n_candidates_out, 0);
}

extern "C" int dvfe_op_good_features_cuda(const uint8_t* img, int w, int h, int pitch, const float* eig, const uint8_t* mask,
                                          int mask_pitch, int max_corners, double quality, double min_dist, float* corners_out,
                                          int* n_out, int* n_candidates_out) {
    return op_good_features(img, w, h, pitch, eig, mask, mask_pitch, max_corners, quality, min_dist, corners_out, n_out, n_candidates_out, 1);
}

extern "C" int dvfe_op_disc_mask(uint8_t* mask, int w, int h, int pitch, const float* pts, int n, int radius) {
    if (!mask || w < 1 || h < 1 || pitch < w || n < 0 || radius < 0 || (n > 0 && !pts)) {
        dvfe_set_error("op_disc_mask: bad argument");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(ensure_device());
    if (n == 0) return DVFE_OK;
    DevBuf d_mask, d_pts, d_n;
    DVFE_CHECK(upload_image(d_mask, mask, w, h, pitch));
    DVFE_CHECK(d_pts.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_n.alloc(sizeof(int)));
    DVFE_CUDA(cudaMemcpy(d_pts.p, pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    DVFE_CUDA(cudaMemcpy(d_n.p, &n, sizeof(int), cudaMemcpyHostToDevice));
    DVFE_CHECK(launch_disc_mask(d_mask.as<uint8_t>(), w, w, h, d_pts.as<float2>(), d_n.as<int>(), n, radius, 0));
    DVFE_CUDA(cudaDeviceSynchronize());
    DVFE_CUDA(cudaMemcpy2D(mask, pitch, d_mask.p, w, w, h, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}

extern "C" int dvfe_op_erode_rect(const uint8_t* src, int w, int h, int pitch, int k, uint8_t* dst) {
    if (!src || !dst || w < 1 || h < 1 || pitch < w || k < 1) { dvfe_set_error("op_erode_rect: bad argument"); return DVFE_ERR_INVALID; }
    DVFE_CHECK(ensure_device());
    DevBuf d_src, d_tmp, d_dst;
    DVFE_CHECK(upload_image(d_src, src, w, h, pitch));
    DVFE_CHECK(d_tmp.alloc((size_t)w * h));
    DVFE_CHECK(d_dst.alloc((size_t)w * h));
    DVFE_CHECK(launch_erode_rect(d_src.as<uint8_t>(), w, d_dst.as<uint8_t>(), w, d_tmp.as<uint8_t>(), w, h, k, 1,
                                 (size_t)w * h, nullptr, 0));
    DVFE_CUDA(cudaDeviceSynchronize());
    DVFE_CUDA(cudaMemcpy(dst, d_dst.p, (size_t)w * h, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}

extern "C" int dvfe_op_lift_projective(const dvfe_camera* cam, const float* pts, int n, float off_x, float off_y,
                                       float* out) {
    if (!cam || n < 0 || (n > 0 && (!pts || !out))) { dvfe_set_error("op_lift_projective: bad argument"); return DVFE_ERR_INVALID; }
    DVFE_CHECK(ensure_device());
    if (n == 0) return DVFE_OK;
    DevBuf d_in, d_out;
    DVFE_CHECK(d_in.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_out.alloc(sizeof(float2) * n));
    DVFE_CUDA(cudaMemcpy(d_in.p, pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    DVFE_CHECK(launch_lift(make_cam(*cam), d_in.as<float2>(), n, off_x, off_y, d_out.as<float2>(), 0));
    DVFE_CUDA(cudaDeviceSynchronize());
    DVFE_CUDA(cudaMemcpy(out, d_out.p, sizeof(float2) * n, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}
