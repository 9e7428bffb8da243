// Shi-Tomasi corner detection with region mask + min-distance suppression — replaces
//   cv::circle(mask, pt, min_dist, 0, -1) per tracked point   (front_end/background_tracker.cpp:79-80,
//                                                               front_end/instance_feature.cpp:367-369,
//                                                               front_end/dynamic_tracker.cpp:430)
//   cv::goodFeaturesToTrack(gray, pts, K, 0.01, min_dist, mask) (front_end/background_tracker.cpp:85,
//                                                               front_end/instance_feature.cpp:381,
//                                                               front_end/dynamic_tracker.cpp:435)
// and the id assignment that follows (background_tracker.cpp:92-96).  Arithmetic restated from OpenCV
// 3.4.16 modules/imgproc/src/{corner,featureselect}.cpp (oracle/spec.c):
//   response   lambda = (a+c) - sqrt((a-c)^2 + b^2), a = Sxx/2, b = Sxy, c = Syy/2, with Sobel 3x3 scaled by
//              1/(4*3*255) (fp32, the FMA forms of OpenCV's SIMD filters) and an unnormalised 3x3 box sum
//              accumulated in fp64;
//   candidates lambda > (float)(max_masked * 0.01), 3x3 local maximum, mask != 0, 1 <= x < W-1, 1 <= y < H-1;
//   selection  greedy in (lambda desc, address desc) order, reject if an accepted corner is closer than
//              min_dist (dx^2+dy^2 < min_dist^2), stop at K.
// The greedy pass is the lexicographically-first maximal independent set of the conflict graph; it is
// computed in parallel by rounds (a candidate is accepted once every stronger neighbour is rejected,
// rejected once a stronger neighbour is accepted), which yields exactly the sequential result, and the
// first K accepted corners in rank order are the reference's output.
#include <float.h>
#include <string.h>
#include <math.h>

#include <cuda.h>

#include "kernels.cuh"

#define NMS_THREADS 1024
#define NMS_MAX_K 2048

__device__ __forceinline__ int f2ord(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : (o ^ 0x7fffffff)); }

__device__ __forceinline__ bool gftt_job_active(const GfttJob& J) {
    const int K = J.max_cnt - *J.n;
    return K > 0 && K >= J.min_needed;
}

// ---- detection mask: region (or 255) ... -------------------------------------------------------
// one thread = 16 bytes of a mask row (mask rows are 16-byte aligned)
__global__ void __launch_bounds__(256) k_gftt_mask_fill(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J)) return;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 8 && threadIdx.y == 0) {
        // reset the per-job counters: n_precand, max, overflow, ...
        J.counters[threadIdx.x] = (threadIdx.x == 1) ? INT_MIN : 0;
    }
    if (x >= J.w || y >= J.h) return;
    uint8_t* m = J.mask + (size_t)y * J.mask_pitch + x;
    const bool aligned_dst = (J.mask_pitch & 15) == 0 && ((uintptr_t)J.mask & 15) == 0;   // the row may be over-written
    if (J.region_mask == nullptr) {                                                       // up to the pitch
        if (aligned_dst) *reinterpret_cast<uint4*>(m) = make_uint4(~0u, ~0u, ~0u, ~0u);
        else for (int i = 0; i < 16 && x + i < J.w; i++) m[i] = 255;
    } else {
        const uint8_t* r = J.region_mask + (size_t)y * J.region_pitch + x;
        if (aligned_dst && x + 15 < J.w && (((uintptr_t)r) & 15) == 0) *reinterpret_cast<uint4*>(m) = *reinterpret_cast<const uint4*>(r);
        else for (int i = 0; i < 16 && x + i < J.w; i++) m[i] = r[i];
    }
}

// ---- ... minus a filled disc around every tracked point:  cleared <=> dx^2+dy^2 <= r^2 ------------------
// one warp per disc; a lane clears whole row spans [cx - hw, cx + hw], hw = floor(sqrt(r^2 - dy^2))
__device__ __forceinline__ void disc_clear(uint8_t* __restrict__ mask, int pitch, int w, int h, float2 p, int r, int lane) {
    const int cx = __float2int_rn(p.x), cy = __float2int_rn(p.y);
    const int r2 = r * r;
    for (int dy = -r + lane; dy <= r; dy += 32) {
        const int y = cy + dy;
        if (y < 0 || y >= h) continue;
        const int rem = r2 - dy * dy;
        int hw = (int)sqrtf((float)rem);
        while (hw * hw > rem) hw--;
        while ((hw + 1) * (hw + 1) <= rem) hw++;
        int xa = max(cx - hw, 0), xb = min(cx + hw, w - 1);      // inclusive
        uint8_t* row = mask + (size_t)y * pitch;
        while (xa <= xb && (((uintptr_t)(row + xa)) & 3)) row[xa++] = 0;
        for (; xa + 3 <= xb; xa += 4) *reinterpret_cast<unsigned*>(row + xa) = 0u;
        while (xa <= xb) row[xa++] = 0;
    }
}

__global__ void __launch_bounds__(128) k_gftt_discs(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.y];
    if (!gftt_job_active(J)) return;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= *J.n) return;
    disc_clear(J.mask, J.mask_pitch, J.w, J.h, J.pts[i], J.disc_radius, threadIdx.x & 31);
}

__global__ void __launch_bounds__(128) k_disc_mask_op(uint8_t* mask, int pitch, int w, int h, const float2* pts, const int* n,
                                                      int r) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= *n) return;
    disc_clear(mask, pitch, w, h, pts[i], r, threadIdx.x & 31);
}

// ---- response map (cv::cornerMinEigenVal, blockSize 3, ksize 3), masked max, local maxima ------------------
// Marching stencil: a warp owns a strip of 28 columns x RS_ROWS rows and walks down the image one row per step
// with everything in registers.  Lane l holds column xb + l - 2 (two halo columns on each side); horizontal
// neighbours come from warp shuffles.  Taking image row y yields, in the same step,
//     Sobel derivatives and their products (the CV_32F cov image) of row y-1,
//     the horizontal 3-sums H(y-1) in double        (cv::boxFilter RowSum<float,double>),
//     the vertical 3-sum / lambda of row y-2        (ColumnSum<double,float> + calcMinEigenVal),
//     the 3x3 local-maximum test of row y-3.
// Nothing but the pre-candidates (local maxima with mask != 0) and the masked maximum leaves the chip; the
// quality threshold needs the global maximum and is applied by the selection kernel.
//
// The strip's pixels are staged once in shared memory (cp.async straight from the padded pyramid level, whose REFLECT_101
// border is the one cornerMinEigenVal wants; images without a border are gathered with reflected indices), so the march
// never waits for global memory.  The detection mask is never materialised: the warp builds its strip of it as one 32-bit
// word per row -- region mask (or all ones) minus the filled discs dx^2+dy^2 <= r^2 around the tracked points that reach
// into the strip -- and the march SKIPS every image row that no unmasked pixel depends on: lambda is needed at mask != 0
// (masked maximum, candidates) and one row around it (3x3 local maximum), i.e. image rows within 3 of an unmasked row.
// Skipped rows leave stale values in the carried registers; every value a kept row consumes is recomputed first (a run of
// kept rows starts 3 rows above its first unmasked row), so the outputs are exactly those of the full march.
#define RS_COLS 28
#ifndef RS_ROWS
#define RS_ROWS 48                // <= 58: the strip's row bitmap has 64 bits
#endif
#define RS_ROWS_SMALL 16
#ifndef RS_WARPS
#define RS_WARPS 1
#endif
#ifndef RS_UNROLL
#define RS_UNROLL 1
#endif
constexpr int kRsUnroll = RS_UNROLL;
#define RS_TILE_ROWS (RS_ROWS + 6)
#define RS_TILE_PITCH 64          // 34 columns (28 + 2 x 3 halo) starting at any byte of a 16-byte aligned row segment: the TMA box
                                  // must START on a 16-byte boundary of the tensor row (a misaligned start coordinate raises an
                                  // illegal-instruction fault: scripts/probes/tma_probe.cu), so the box is 64 bytes wide

#ifndef RS_OPT_I2F
#define RS_OPT_I2F 1
#endif
// u8 -> float without the XU-pipe I2F: 2^23 + v as bits, minus 2^23 (exact)
__device__ __forceinline__ float u8_to_float(unsigned v) {
#if RS_OPT_I2F
    return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f);
#else
    return (float)v;
#endif
}

struct __align__(128) RespSmem {
    uint8_t img[RS_TILE_ROWS * RS_TILE_PITCH];   // rows y0-3 .. y1+2 of the strip, columns xb-3 .. xb+30 (+ alignment); TMA box
    unsigned bits[64];                           // detection mask of rows y0 .. y1-1: bit l = column xb + l - 2 is unmasked
    unsigned long long mbar;                     // completion barrier of the TMA tile load
};

struct RespCtx {                 // per-strip constants of the marching stencil
    int w, h, y0, y1, lane, x;
    bool owned_col, cand_col, x_border, max_all;
    float* __restrict__ eig;
    int* __restrict__ counters; unsigned long long* __restrict__ cand; int cand_cap;
};
struct RespState {               // registers carried from row to row
    float dxr0 = 0.f, dxr1 = 0.f, smr0 = 0.f, smr1 = 0.f;             // row-filter outputs of rows y-2, y-1
    double h0x = 0.0, h0y = 0.0, h0z = 0.0, h1x = 0.0, h1y = 0.0, h1z = 0.0;   // H(y-3), H(y-2)
    float lam1 = 0.f, lr1 = 0.f, hm1 = 0.f, hm0 = 0.f;                 // lambda row y-3 (lam, left/right max), hm of y-3, y-4
    int best = INT_MIN;
    unsigned mk1 = 0;                                                  // mask of row y-3 at this column
};

// One step of the marching stencil: take image row y (this lane's pixel `pc`, for lanes 0 / 31 the pixel outside the warp's
// window `pe`), finish the derivatives of row y-1, lambda of row y-2 and the local-maximum test of row y-3.  `pm` = the
// detection mask of row y-2 at this column (0 outside the strip's rows).  FAST = the step is interior: rows y-3..y inside the
// image and the strip (no box-filter border rule, every row owned) and the warp's 32 columns inside the image; all the index
// tests fold away.  The arithmetic is the same expression by expression in both variants.
template <bool WRITE_EIG, bool EMIT, bool FAST>
__device__ __forceinline__ void resp_step(const RespCtx& C, RespState& S, int y, unsigned pc, unsigned pe, unsigned pm) {
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float s2 = s * 2.0f;
    const int lane = C.lane, x = C.x, w = C.w, h = C.h;
    const float c = u8_to_float(pc);
    const int yl = y - 2;
    const unsigned mk2 = pm;
    float l = __shfl_up_sync(0xffffffffu, c, 1), r = __shfl_down_sync(0xffffffffu, c, 1);
    {
        const float e = u8_to_float(pe);
        if (lane == 0) l = e;
        if (lane == 31) r = e;
    }
    // row filters: [-1 0 1] exact; [1 2 1]*scale as fma(s, r, fma(2s, c, s*l))
    const float dxr2 = r - l;
    const float smr2 = __fmaf_rn(s, r, __fmaf_rn(s2, c, __fmul_rn(s, l)));
    // ---- derivatives of row y-1: column filters [1 2 1]*scale -> fma(s, d0 + d2, (2s)*d1) ; [-1 0 1] ----
    const float dx = __fmaf_rn(s, __fadd_rn(S.dxr0, dxr2), __fmul_rn(s2, S.dxr1));
    const float dy = __fsub_rn(smr2, S.smr0);
    S.dxr0 = S.dxr1; S.dxr1 = dxr2; S.smr0 = S.smr1; S.smr1 = smr2;
    float pxx = __fmul_rn(dx, dx), pxy = __fmul_rn(dx, dy), pyy = __fmul_rn(dy, dy);
    // box-filter border rule (REFLECT_101 on the cov image): column -1 takes column 1, column w takes w-2
    if (!FAST && C.x_border) {
        const float axx = __shfl_down_sync(0xffffffffu, pxx, 2), axy = __shfl_down_sync(0xffffffffu, pxy, 2),
                    ayy = __shfl_down_sync(0xffffffffu, pyy, 2);
        const float bxx = __shfl_up_sync(0xffffffffu, pxx, 2), bxy = __shfl_up_sync(0xffffffffu, pxy, 2),
                    byy = __shfl_up_sync(0xffffffffu, pyy, 2);
        if (x == -1) { pxx = axx; pxy = axy; pyy = ayy; }
        if (x == w) { pxx = bxx; pxy = bxy; pyy = byy; }
    }
    // ---- H(y-1): horizontal 3-sum in double, left to right ----
    // widen once and move the doubles: 3 conversions + 12 shuffles instead of 9 conversions + 6 shuffles (a conversion holds
    // the 4-lane XU pipe of the sub-partition for 8 cycles; with 9 + 3 + sqrt of them per step the pipe bounds the march)
    const double qxx = (double)pxx, qxy = (double)pxy, qyy = (double)pyy;
    double h2x = (__shfl_up_sync(0xffffffffu, qxx, 1) + qxx) + __shfl_down_sync(0xffffffffu, qxx, 1);
    double h2y = (__shfl_up_sync(0xffffffffu, qxy, 1) + qxy) + __shfl_down_sync(0xffffffffu, qxy, 1);
    double h2z = (__shfl_up_sync(0xffffffffu, qyy, 1) + qyy) + __shfl_down_sync(0xffffffffu, qyy, 1);
    // ---- lambda of row yl = y-2: vertical 3-sum top to bottom; rows -1 / h take rows 1 / h-2 ----
    double ax = S.h0x, ay = S.h0y, az = S.h0z;
    if (!FAST && (yl == 0 || yl == h - 1)) {
        if (yl == 0) { ax = h2x; ay = h2y; az = h2z; }
        if (yl == h - 1) { h2x = S.h0x; h2y = S.h0y; h2z = S.h0z; }
    }
    const float cxx = (float)((ax + S.h1x) + h2x);
    const float cxy = (float)((ay + S.h1y) + h2y);
    const float cyy = (float)((az + S.h1z) + h2z);
    S.h0x = S.h1x; S.h0y = S.h1y; S.h0z = S.h1z; S.h1x = h2x; S.h1y = h2y; S.h1z = h2z;
    float lam = 0.f;
    if (FAST || (yl >= 0 && yl < h && x >= 0 && x < w)) {
        const float a = __fmul_rn(cxx, 0.5f), b = cxy, cc = __fmul_rn(cyy, 0.5f);
        const float amc = __fsub_rn(a, cc);
        lam = __fsub_rn(__fadd_rn(a, cc), sqrtf(__fadd_rn(__fmul_rn(amc, amc), __fmul_rn(b, b))));
        if (C.owned_col && (FAST || (yl >= C.y0 && yl < C.y1))) {
            if (WRITE_EIG) C.eig[(size_t)yl * w + x] = lam;
            if (EMIT && (mk2 != 0 || C.max_all)) S.best = max(S.best, f2ord(lam));
        }
    }
    if (!EMIT) return;
    // ---- 3x3 local maximum of row yc = y-3 ----
    const float ll = __shfl_up_sync(0xffffffffu, lam, 1), rl = __shfl_down_sync(0xffffffffu, lam, 1);
    const float lr2 = fmaxf(ll, rl), hm2 = fmaxf(lr2, lam);
    const int yc = y - 3;
    bool is_cand = false;
    if (C.cand_col && (FAST || (yc >= C.y0 && yc < C.y1 && yc >= 1 && yc < h - 1)) && S.lam1 != 0.f) {
        const float m = fmaxf(fmaxf(S.hm0, hm2), S.lr1);
        if (!(m > S.lam1)) is_cand = S.mk1 != 0;
    }
    const float v = S.lam1;
    S.hm0 = S.hm1; S.hm1 = hm2; S.lam1 = lam; S.lr1 = lr2; S.mk1 = mk2;
    const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
    if (ballot) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&C.counters[0], __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_cand) {
            const int pos = base + __popc(ballot & ((1u << lane) - 1));
            if (pos < C.cand_cap)
                C.cand[pos] = ((unsigned long long)((unsigned)f2ord(v) ^ 0x80000000u) << 32) | (unsigned)(yc * w + x);
            else
                C.counters[2] = 1;
        }
    }
}

__device__ __forceinline__ void rs_cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}

// What a strip is detected on: image (+ whether it carries a readable REFLECT_101 border of >= 3 px and 4-byte aligned
// rows), region mask, the points whose discs are cut out of it.
struct RespIn {
    const uint8_t* __restrict__ img; int pitch, w, h; bool bordered;
    const CUtensorMap* tmap; int tma_x, tma_y;      // non-null: the tile is box (tma_x, tma_y) of this 2-D tensor (TMA load)
    const uint8_t* __restrict__ region; int region_pitch;     // nullable
    const float2* __restrict__ pts; int n_pts, radius;        // discs (n_pts = 0: none)
    bool max_all;                                             // the maximum is taken over every pixel (GfttJob::max_unmasked)
};

template <bool WRITE_EIG, bool EMIT>
__device__ __forceinline__ void resp_strip(RespSmem& sm, const RespIn& in, int xb, int y0, int rows, float* __restrict__ eig,
                                           int* __restrict__ counters, unsigned long long* __restrict__ cand, int cand_cap) {
    const int w = in.w, h = in.h;
    const int lane = threadIdx.x & 31;
    const int y1 = min(y0 + rows, h), rows_n = y1 - y0, n_t = rows_n + 6;
    // ---- 1. the strip's pixels: rows y0-3 .. y1+2, columns xb-3 .. xb+30 ----
    int coff;                                      // tile column of image column xb-3
    if (in.tmap != nullptr) {
        // One TMA box load (cp.async.bulk.tensor, 64 x 54 bytes) instead of 17 cp.async per lane: lane 0 arms the barrier
        // with the box size and issues it; everybody waits on the barrier after the mask is built.
        coff = in.tma_x & 15;
        const unsigned mbar = (unsigned)__cvta_generic_to_shared(&sm.mbar);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(RS_TILE_ROWS * RS_TILE_PITCH) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"((unsigned)__cvta_generic_to_shared(sm.img)), "l"((unsigned long long)in.tmap), "r"(in.tma_x & ~15), "r"(in.tma_y),
                           "r"(mbar) : "memory");
        }
    } else if (in.bordered) {
        const int x_al = (xb - 3) & ~3;
        coff = (xb - 3) - x_al;
        const uint8_t* __restrict__ src = in.img + (ptrdiff_t)(y0 - 3) * in.pitch + x_al;
        for (int t = lane; t < n_t * 10; t += 32) {            // 10 aligned words hold the 34 columns
            const int r = t / 10, w4 = (t - r * 10) * 4;
            rs_cp_async4(sm.img + r * RS_TILE_PITCH + w4, src + (ptrdiff_t)r * in.pitch + w4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        coff = 0;
        for (int t = lane; t < n_t * 34; t += 32) {
            const int r = t / 34, c = t - r * 34;
            sm.img[r * RS_TILE_PITCH + c] = __ldg(in.img + (size_t)reflect101(y0 - 3 + r, h) * in.pitch + reflect101(xb - 3 + c, w));
        }
    }
    const int x = xb + lane - 2;
    const bool owned_col = lane >= 2 && lane < 2 + RS_COLS && x < w;
    unsigned long long need;
    if (EMIT) {
        // ---- 2. the strip's detection mask, one word per row ----
        sm.bits[lane] = 0u; sm.bits[lane + 32] = 0u;
        __syncwarp();
        if (in.region != nullptr) {
            for (int r = 0; r < rows_n; r++) {
                const unsigned v = owned_col ? __ldg(in.region + (size_t)(y0 + r) * in.region_pitch + x) : 0u;
                const unsigned b = __ballot_sync(0xffffffffu, v != 0u);
                if (lane == 0) sm.bits[r] = b;
            }
        } else {
            const unsigned b = __ballot_sync(0xffffffffu, owned_col);
            for (int r = lane; r < rows_n; r += 32) sm.bits[r] = b;
        }
        __syncwarp();
        const int rad = in.radius, r2 = rad * rad;
        for (int i0 = 0; i0 < in.n_pts; i0 += 32) {
            const int i = i0 + lane;
            int cx = 0, cy = 0;
            bool hit = false;
            if (i < in.n_pts) {
                const float2 p = in.pts[i];
                cx = __float2int_rn(p.x); cy = __float2int_rn(p.y);
                hit = cx + rad >= xb && cx - rad <= xb + RS_COLS - 1 && cy + rad >= y0 && cy - rad < y1;
            }
            unsigned bal = __ballot_sync(0xffffffffu, hit);
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const int dcx = __shfl_sync(0xffffffffu, cx, src), dcy = __shfl_sync(0xffffffffu, cy, src);
                for (int dy = -rad + lane; dy <= rad; dy += 32) {        // a lane clears whole row spans [cx - hw, cx + hw]
                    const int yy = dcy + dy;
                    if (yy < y0 || yy >= y1) continue;
                    const int rem = r2 - dy * dy;
                    int hw = (int)sqrtf((float)rem);
                    while (hw * hw > rem) hw--;
                    while ((hw + 1) * (hw + 1) <= rem) hw++;
                    const int xa = max(dcx - hw, xb), xe = min(dcx + hw, xb + RS_COLS - 1);      // inclusive
                    if (xa <= xe) sm.bits[yy - y0] &= ~(((2u << (xe - xb + 2)) - 1u) & ~((1u << (xa - xb + 2)) - 1u));
                }
                __syncwarp();
            }
        }
        __syncwarp();
        // ---- 3. rows to march: image row y0-3+j is needed iff an unmasked pixel lies within 3 rows of it ----
        const unsigned lo = __ballot_sync(0xffffffffu, sm.bits[lane] != 0u), hi = __ballot_sync(0xffffffffu, sm.bits[lane + 32] != 0u);
        const unsigned long long occ = ((unsigned long long)hi << 32) | lo;
        need = occ | (occ << 1) | (occ << 2) | (occ << 3) | (occ << 4) | (occ << 5) | (occ << 6);
        if (in.max_all) need = (1ull << (rows_n + 6)) - 1ull;      // the whole-image maximum needs every row (y0-3 .. y1+2)
    } else {
        need = (1ull << (rows_n + 5)) - 1ull;          // every row: y0-3 .. y1+1
    }
    if (in.tmap != nullptr) {
        __syncwarp();                      // the barrier initialised by lane 0 is visible
        const unsigned mbar = (unsigned)__cvta_generic_to_shared(&sm.mbar);
        asm volatile("{\n\t.reg .pred p;\n\tRS_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra RS_DONE;\n\tbra RS_WAIT;\n\tRS_DONE:\n\t}"
                     ::"r"(mbar) : "memory");
    } else if (in.bordered) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    if (need == 0ull) return;

    RespCtx C;
    C.w = w; C.h = h; C.y0 = y0; C.y1 = y1; C.lane = lane; C.x = x;
    C.owned_col = owned_col;
    C.cand_col = owned_col && x >= 1 && x < w - 1;
    C.x_border = xb < 2 || xb + 30 > w;
    C.max_all = in.max_all;
    C.eig = eig; C.counters = counters; C.cand = cand; C.cand_cap = cand_cap;
    RespState S;
    // interior steps: rows y-3..y inside the image and owned by the strip, all 32 columns (plus the edge columns) inside
    const int f0 = max(y0 + 3, 4), f1 = C.x_border ? f0 - 1 : min(y1 + 1, h - 1);      // [f0, f1]
    const uint8_t* __restrict__ pc = sm.img + coff + lane + 1;
    const uint8_t* __restrict__ pe = sm.img + coff + (lane == 0 ? 0 : (lane == 31 ? 33 : lane + 1));
    const int jf0 = f0 - (y0 - 3), jf1 = f1 - (y0 - 3);          // steps [jf0, jf1] are interior
    while (need) {
        // one run of consecutive needed rows [j, jend)
        int j = __ffsll((long long)need) - 1;
        const unsigned long long rest = ~(need >> j);
        const int len = rest == 0ull ? 64 - j : __ffsll((long long)rest) - 1;
        const int jend = j + len;
        need = jend >= 64 ? 0ull : (need >> jend) << jend;
        const uint8_t* __restrict__ qc = pc + j * RS_TILE_PITCH;
        const uint8_t* __restrict__ qe = pe + j * RS_TILE_PITCH;
        // ONE copy of each step variant in the instruction stream: the kernel lives in the instruction cache (6x unrolled
        // variants at three call sites made 25 % of the stall samples instruction fetches; measured on the B200 at 64 x 720p:
        // unroll 1 / 2 / 4 -> 0.245 / 0.252 / 0.278 ms)
#pragma unroll kRsUnroll
        for (; j < jend; j++) {
            unsigned vm = 0u;
            if (EMIT && j >= 5 && j < rows_n + 5) vm = (sm.bits[j - 5] >> lane) & 1u;      // mask of row y-2
            if (j >= jf0 && j <= jf1) resp_step<WRITE_EIG, EMIT, true>(C, S, y0 - 3 + j, *qc, *qe, vm);
            else resp_step<WRITE_EIG, EMIT, false>(C, S, y0 - 3 + j, *qc, *qe, vm);
            qc += RS_TILE_PITCH; qe += RS_TILE_PITCH;
        }
    }
    if (EMIT) {
        S.best = __reduce_max_sync(0xffffffffu, S.best);
        if (lane == 0 && S.best != INT_MIN) atomicMax(&counters[1], S.best);
    }
}

// rows = strip height: RS_ROWS when the launch has enough warps to fill the GPU, RS_ROWS_SMALL for small batches (a single
// camera), where three times as many, shorter strips cut the latency of the serial march
// tmap (use_tma != 0): level 0 of the padded pyramids of ALL jobs of the launch as one 2-D u8 tensor (rows of `img_pitch` bytes,
// job j's pixel (0,0) at row j * tma_rows_per_job + DVFE_PADY, column DVFE_PADX)
__global__ void __launch_bounds__(RS_WARPS * 32, 32 / RS_WARPS) k_gftt_response(const GfttJob* __restrict__ jobs, int rows,
                                                                                 const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                                                 int tma_rows_per_job) {
    __shared__ RespSmem s_resp[RS_WARPS];
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J) || J.eig_in != nullptr) return;
    const int xb = (blockIdx.x * RS_WARPS + (threadIdx.x >> 5)) * RS_COLS, y0 = blockIdx.y * rows;
    if (xb >= J.w || y0 >= J.h) return;
    RespIn in;
    in.img = J.img; in.pitch = J.img_pitch; in.w = J.w; in.h = J.h; in.bordered = J.img_bordered != 0;
    in.tmap = (use_tma && in.bordered) ? &tmap : nullptr;
    in.tma_x = DVFE_PADX + xb - 3; in.tma_y = (int)blockIdx.z * tma_rows_per_job + DVFE_PADY + y0 - 3;
    in.region = J.region_mask; in.region_pitch = J.region_pitch;
    in.pts = J.pts; in.n_pts = *J.n; in.radius = J.disc_radius;
    in.max_all = J.max_unmasked != 0;
    resp_strip<false, true>(s_resp[threadIdx.x >> 5], in, xb, y0, rows, nullptr, J.counters, J.cand, J.cand_cap);
}

__global__ void __launch_bounds__(RS_WARPS * 32, 32 / RS_WARPS) k_min_eigen_val(const uint8_t* img, int pitch, int w, int h, float* eig) {
    __shared__ RespSmem s_resp[RS_WARPS];
    const int xb = (blockIdx.x * RS_WARPS + (threadIdx.x >> 5)) * RS_COLS, y0 = blockIdx.y * RS_ROWS;
    if (xb >= w || y0 >= h) return;
    RespIn in;
    in.img = img; in.pitch = pitch; in.w = w; in.h = h; in.bordered = false;
    in.tmap = nullptr; in.tma_x = 0; in.tma_y = 0;
    in.region = nullptr; in.region_pitch = 0; in.pts = nullptr; in.n_pts = 0; in.radius = 0; in.max_all = false;
    resp_strip<true, false>(s_resp[threadIdx.x >> 5], in, xb, y0, RS_ROWS, eig, nullptr, nullptr, 0);
}

// ---- externally supplied response map (seam op): masked max, then the same pre-candidates ---------------
__global__ void __launch_bounds__(256) k_gftt_max_ext(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J) || J.eig_in == nullptr) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    int best = INT_MIN;
    if (x < J.w && y < J.h && (J.max_unmasked != 0 || J.mask[(size_t)y * J.mask_pitch + x] != 0)) best = f2ord(J.eig_in[(size_t)y * J.w + x]);
    best = __reduce_max_sync(0xffffffffu, best);
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && best != INT_MIN) atomicMax(&J.counters[1], best);
}

__global__ void __launch_bounds__(256) k_gftt_candidates_ext(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J) || J.eig_in == nullptr) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const float* __restrict__ eig = J.eig_in;
    bool is_cand = false;
    float v = 0.f;
    if (x >= 1 && x < J.w - 1 && y >= 1 && y < J.h - 1) {
        v = eig[(size_t)y * J.w + x];
        if (v != 0.f && J.mask[(size_t)y * J.mask_pitch + x] != 0) {
            const float* p = eig + (size_t)y * J.w + x;
            float m = fmaxf(fmaxf(p[-J.w - 1], p[-J.w]), fmaxf(p[-J.w + 1], p[-1]));
            m = fmaxf(m, fmaxf(fmaxf(p[1], p[J.w - 1]), fmaxf(p[J.w], p[J.w + 1])));
            is_cand = !(m > v);
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
    if (ballot == 0) return;
    const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&J.counters[0], __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (is_cand) {
        const int pos = base + __popc(ballot & ((1u << lane) - 1));
        if (pos < J.cand_cap)
            J.cand[pos] = ((unsigned long long)((unsigned)f2ord(v) ^ 0x80000000u) << 32) | (unsigned)(y * J.w + x);
        else
            J.counters[2] = 1;
    }
}

// ---- selection: quality threshold, parallel greedy min-distance suppression, top-K; one CTA per job -------------
__device__ __forceinline__ int block_excl_scan_inplace(int* data, int n, int* s_part /* [NMS_THREADS] */) {
    // exclusive scan of data[0..n) in place, returns the total; all threads of the block participate
    const int tid = threadIdx.x;
    const int per = (n + NMS_THREADS - 1) / NMS_THREADS;
    const int b = min(tid * per, n), e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += data[i];
    // block-wide inclusive scan of the partial sums: warp shuffles + one pass over the 32 warp totals
    const int lane = tid & 31, warp = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    __syncthreads();
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int wsum = s_part[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wsum, off);
            if (lane >= off) wsum += v;
        }
        s_part[32 + lane] = wsum;      // inclusive totals per warp
    }
    __syncthreads();
    const int total = s_part[32 + 31];
    int run = incl - sum + (warp > 0 ? s_part[32 + warp - 1] : 0);
    for (int i = b; i < e; i++) {
        const int v = data[i];
        data[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

// K-th largest key (1-based k) among keys[0..n) that satisfy pred; 8 radix passes of 8 bits from the top.
template <typename Pred>
__device__ __forceinline__ unsigned long long block_select_kth(const unsigned long long* __restrict__ keys, int n, int k,
                                                               Pred pred, int* s_hist, unsigned long long* s_prefix,
                                                               int* s_remaining) {
    const int tid = threadIdx.x;
    if (tid == 0) { *s_prefix = 0ull; *s_remaining = k; }
    __syncthreads();
    for (int pass = 0; pass < 8 && *s_remaining > 0; pass++) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = *s_prefix;
        const unsigned long long himask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
        for (int i = tid; i < n; i += NMS_THREADS) {
            const unsigned long long key = keys[i];
            if ((key & himask) == prefix && pred(i, key)) atomicAdd(&s_hist[(int)((key >> shift) & 255)], 1);
        }
        __syncthreads();
        if (tid < 32) {
            // warp 0: lane l owns bins 255-8l .. 248-8l (descending); find the bin where the count from the top
            // reaches the remaining rank
            const int rem = *s_remaining;
            int c[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { c[q] = s_hist[255 - (8 * tid + q)]; sum += c[q]; }
            int incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (tid >= off) incl += v;
            }
            const int before = incl - sum;
            const bool mine = before < rem && incl >= rem;
            const unsigned who = __ballot_sync(0xffffffffu, mine);
            if (who == 0) {
                // fewer keys than the rank (cannot happen for k <= population); keep everything
                if (tid == 0) { *s_remaining = 0; *s_prefix = prefix; }
            } else if (mine) {
                int r = rem - before, q = 0;
                while (q < 7 && c[q] < r) { r -= c[q]; q++; }
                const int b = 255 - (8 * tid + q);
                // every key of the selected bucket is wanted: all keys >= this prefix are exactly the k largest
                *s_remaining = (c[q] == r) ? 0 : r;
                *s_prefix = prefix | ((unsigned long long)b << shift);
            }
        }
        __syncthreads();
    }
    return *s_prefix;
}

#define NMS_SMEM_STATE 16384
// The cell table and the working keys live in (dynamic) shared memory when they fit: every round of the parallel
// greedy pass chases cell -> key -> state, and from global memory each hop is an L2 round trip.
#define NMS_SMEM_CELLS 8192
#define NMS_SMEM_KEYS 8192
#define NMS_DYN_SMEM (2 * (NMS_SMEM_CELLS + 2) * 4 + NMS_SMEM_KEYS * 8)

__global__ void __launch_bounds__(NMS_THREADS) k_gftt_select(const GfttJob* __restrict__ jobs) {
    __shared__ int s_part[64];
    __shared__ unsigned long long s_sel[NMS_MAX_K];
    __shared__ int s_hist[256];
    __shared__ int s_cnt;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining;
    __shared__ uint8_t s_state[NMS_SMEM_STATE];
    extern __shared__ __align__(16) unsigned char s_dyn[];
    int* const s_cells = reinterpret_cast<int*>(s_dyn);                                               // 2 x (cells + 1)
    unsigned long long* const s_keys = reinterpret_cast<unsigned long long*>(s_dyn + 2 * (NMS_SMEM_CELLS + 2) * 4);

    const GfttJob& J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    if (tid == 0) J.counters[4] = 0;       // number of new corners
    if (!gftt_job_active(J)) return;
    const int n_old = *J.n;
    int K = J.max_cnt - n_old;
    if (K > NMS_MAX_K) K = NMS_MAX_K;
    int nc = J.counters[0];
    const int mo = J.counters[1], overflow = J.counters[2];
    __syncthreads();
    if (tid == 0) {
        // hand the counters back reset for the next response launch (no separate clearing kernel); the overflow flag of this
        // launch stays readable in counters[7]
        J.counters[0] = 0; J.counters[1] = INT_MIN; J.counters[2] = 0; J.counters[7] = overflow;
        if (overflow && J.err) atomicOr(J.err, 1);                  // more local maxima than the buffer holds
    }
    if (nc > J.cand_cap) nc = J.cand_cap;
    if (nc <= 0) return;

    // quality threshold: keep lambda > (float)(maxVal * quality)   (cv::threshold THRESH_TOZERO, strict)
    const double maxVal = (mo == INT_MIN) ? 0.0 : (double)ord2f(mo);
    const float thr = (float)(maxVal * J.quality);
    const unsigned long long thr_key = ((unsigned long long)((unsigned)f2ord(thr) ^ 0x80000000u) << 32) | 0xffffffffull;
    const unsigned long long* all = J.cand;

    // candidates above the threshold, compacted into cand3 (unordered)
    unsigned long long* __restrict__ valid = J.cand3;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nc; i0 += NMS_THREADS) {
        const int i = i0 + tid;
        const unsigned long long key = i < nc ? all[i] : 0ull;
        const bool keep = i < nc && key > thr_key;
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        int base = 0;
        if ((tid & 31) == 0 && ballot) base = atomicAdd(&s_cnt, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) valid[base + __popc(ballot & ((1u << (tid & 31)) - 1))] = key;
    }
    __syncthreads();
    const int nv = s_cnt;                   // candidates above the threshold (cv: tmpCorners.size())
    __syncthreads();
    if (tid == 0) J.counters[3] = nv;
    if (nv == 0) return;

    const int w = J.w;
    const int cell = (int)lrintf(J.min_dist) > 0 ? (int)lrintf(J.min_dist) : 1;
    const int gw = (J.w + cell - 1) / cell, gh = (J.h + cell - 1) / cell;
    const int ncell = gw * gh;
    const bool cells_in_smem = ncell + 1 <= NMS_SMEM_CELLS + 1;
    int* cstart = cells_in_smem ? s_cells : J.cell_count;                  // [ncell + 1]
    int* ccur = cstart + (ncell + 1);                                      // [ncell + 1]
    const double md2 = (double)J.min_dist * (double)J.min_dist;
    const bool use_nms = J.min_dist >= 1.f;

    unsigned long long* work = J.cand2;          // the M strongest candidates (unordered)
    unsigned long long* cs = J.cand;             // the same, grouped by cell (the pre-candidate list is dead by now)
    // Only the first K accepted corners in rank order are wanted, and whether a candidate is accepted depends only
    // on stronger candidates: the greedy result restricted to the M strongest candidates is a prefix of the full
    // result.  Start with a small M and grow it until K corners are accepted or every candidate is in.
    int M = min(nv, max(512, 16 * K));
    int n_acc = 0;
    volatile uint8_t* state = nullptr;
    const unsigned long long* fin = work;        // the examined candidates, indexed like `state`
    for (;;) {
        unsigned long long tkey = 0ull;                 // keep keys >= tkey
        if (M < nv) tkey = block_select_kth(valid, nv, M, [&](int, unsigned long long) { return true; }, s_hist, &s_prefix,
                                            &s_remaining);
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int i0 = 0; i0 < nv; i0 += NMS_THREADS) {
            const int i = i0 + tid;
            const unsigned long long key = i < nv ? valid[i] : 0ull;
            const bool keep = i < nv && key >= tkey;
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            int base = 0;
            if ((tid & 31) == 0 && ballot) base = atomicAdd(&s_cnt, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) work[base + __popc(ballot & ((1u << (tid & 31)) - 1))] = key;
        }
        __syncthreads();
        // (s_cnt == M: keys are unique)
        state = (M <= NMS_SMEM_STATE) ? (volatile uint8_t*)s_state : (volatile uint8_t*)J.state;
        if (use_nms) {
            for (int i = tid; i < 2 * (ncell + 1); i += NMS_THREADS) cstart[i] = 0;
            __syncthreads();
            for (int i = tid; i < M; i += NMS_THREADS) {
                const unsigned idx = (unsigned)work[i];
                const int y = idx / w, x = idx - y * w;
                atomicAdd(&cstart[(y / cell) * gw + x / cell], 1);
            }
            __syncthreads();
            block_excl_scan_inplace(cstart, ncell + 1, s_part);
            for (int i = tid; i < M; i += NMS_THREADS) {
                const unsigned long long key = work[i];
                const unsigned idx = (unsigned)key;
                const int y = idx / w, x = idx - y * w;
                const int c = (y / cell) * gw + x / cell;
                const int pos = atomicAdd(&ccur[c], 1);
                cs[cstart[c] + pos] = key;
            }
            __syncthreads();
            // rank-sort every cell (descending) from cs into the sorted array (shared memory when it fits)
            unsigned long long* srt_w = (M <= NMS_SMEM_KEYS) ? s_keys : work;
            {
                const int warp = tid >> 5, lane = tid & 31;
                for (int c = warp; c < ncell; c += NMS_THREADS / 32) {
                    const int b = cstart[c], m = cstart[c + 1] - b;
                    for (int e = lane; e < m; e += 32) {
                        const unsigned long long key = cs[b + e];
                        int rank = 0;
                        for (int o = 0; o < m; o++) rank += (cs[b + o] > key) ? 1 : 0;
                        srt_w[b + rank] = key;
                    }
                }
            }
            __syncthreads();
            const unsigned long long* srt = srt_w;
            fin = srt_w;
            for (int i = tid; i < M; i += NMS_THREADS) state[i] = 0;
            __syncthreads();
            // Decisions are final once written, so no barrier is needed between rounds: every thread keeps sweeping
            // its own undecided candidates until none is left (the strongest undecided candidate of the block can
            // always be decided by its owner, so the sweep terminates).  The flags are single bytes that only ever go
            // 0 -> 1 or 0 -> 2 and are read through a volatile pointer (a stale 0 means "look again next round"): the
            // read/write overlap compute-sanitizer's racecheck reports here is this protocol (profiles/sanitizer_r2.txt).
            volatile uint8_t* vstate = state;
            for (;;) {
                int undecided = 0;
                for (int i = tid; i < M; i += NMS_THREADS) {
                    if (vstate[i] != 0) continue;
                    const unsigned long long key = srt[i];
                    const unsigned idx = (unsigned)key;
                    const int y = idx / w, x = idx - y * w;
                    const int xc = x / cell, yc = y / cell;
                    const int x1 = max(xc - 1, 0), x2 = min(xc + 1, gw - 1);
                    const int y1 = max(yc - 1, 0), y2 = min(yc + 1, gh - 1);
                    bool rejected = false, blocked = false;
                    for (int yy = y1; yy <= y2 && !rejected; yy++)
                        for (int xx = x1; xx <= x2 && !rejected; xx++) {
                            const int c = yy * gw + xx;
                            const int e = cstart[c + 1];
                            for (int j = cstart[c]; j < e; j++) {
                                const unsigned long long kj = srt[j];
                                if (kj <= key) break;                 // only stronger candidates matter
                                const unsigned ij = (unsigned)kj;
                                const int yj = ij / w, xj = ij - yj * w;
                                const int dx = x - xj, dy = y - yj;
                                if ((double)(dx * dx + dy * dy) < md2) {
                                    const uint8_t sj = vstate[j];
                                    if (sj == 1) { rejected = true; break; }
                                    if (sj == 0) blocked = true;
                                }
                            }
                        }
                    if (rejected) vstate[i] = 2;
                    else if (!blocked) vstate[i] = 1;
                    else undecided = 1;
                }
                if (!__any_sync(0xffffffffu, undecided)) break;      // warp-level: lanes of a warp sweep together
            }
            __syncthreads();
        } else {
            fin = work;
            for (int i = tid; i < M; i += NMS_THREADS) state[i] = 1;
            __syncthreads();
        }
        // count the accepted
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        {
            int c = 0;
            for (int i = tid; i < M; i += NMS_THREADS) c += state[i] == 1 ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
        }
        __syncthreads();
        n_acc = s_cnt;
        __syncthreads();
        if (n_acc >= K || M >= nv) break;
        M = min(nv, 4 * M);
    }

    // ---- the K strongest accepted keys (all of them if fewer), sorted descending ----
    int n_sel = n_acc;
    unsigned long long kth = 0ull;
    if (n_acc > K) {
        kth = block_select_kth(fin, M, K, [&](int i, unsigned long long) { return state[i] == 1; }, s_hist, &s_prefix,
                               &s_remaining);
        n_sel = K;
    }
    int npow = 1;
    while (npow < n_sel) npow <<= 1;
    for (int i = tid; i < npow; i += NMS_THREADS) s_sel[i] = 0ull;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int i = tid; i < M; i += NMS_THREADS) {
        const unsigned long long key = fin[i];
        if (state[i] == 1 && key >= kth) {
            const int pos = atomicAdd(&s_cnt, 1);
            if (pos < NMS_MAX_K) s_sel[pos] = key;
        }
    }
    __syncthreads();
    for (int size = 2; size <= npow; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < npow / 2; t += NMS_THREADS) {
                const int lo = (t / stride) * 2 * stride + (t % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = s_sel[lo], b = s_sel[hi];
                if ((a < b) == desc) { s_sel[lo] = b; s_sel[hi] = a; }
            }
            __syncthreads();
        }
    // ---- append in acceptance order: (float)x, (float)y ----
    for (int r = tid; r < n_sel; r += NMS_THREADS) {
        const unsigned idx = (unsigned)s_sel[r];
        const int y = idx / w, x = idx - y * w;
        J.pts[n_old + r] = make_float2((float)x, (float)y);
    }
    if (tid == 0) { J.counters[4] = n_sel; J.counters[5] = M; J.counters[6] = n_acc; }
}

// ids = global_id_count++ in acceptance order; jobs that share an id counter are served in job order
__global__ void k_gftt_assign_ids(const GfttJob* __restrict__ jobs, int n_jobs) {
    const int j0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (j0 >= n_jobs) return;
    if (j0 > 0 && jobs[j0 - 1].next_id == jobs[j0].next_id && jobs[j0].next_id != nullptr) return;
    for (int j = j0; j < n_jobs && (j == j0 || (jobs[j].next_id == jobs[j0].next_id && jobs[j0].next_id != nullptr)); j++) {
        const GfttJob& J = jobs[j];
        const int cnt = J.counters[4];
        const int n_old = *J.n;
        unsigned next = J.next_id ? *J.next_id : 0u;
        for (int r = 0; r < cnt; r++) {
            if (J.ids) J.ids[n_old + r] = next + (unsigned)r;
            if (J.track_cnt) J.track_cnt[n_old + r] = 1;
        }
        if (J.next_id) *J.next_id = next + (unsigned)cnt;
        *J.n = n_old + cnt;
    }
}

// per-device function attribute of the selection kernel; called at tracker creation so that it never falls inside a graph
// capture, and (cheaply) before every launch for the seam ops
int gftt_prepare_device() {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        DVFE_CUDA(cudaFuncSetAttribute(k_gftt_select, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_DYN_SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    return DVFE_OK;
}
#define DVFE_CHECK_RC(call) do { int rc__ = (call); if (rc__ != DVFE_OK) return rc__; } while (0)

int launch_gftt(const GfttJob* d_jobs, const GfttJob* h_jobs, int n_jobs, int max_w, int max_h, int max_pts,
                cudaStream_t st, cudaEvent_t* marks, cudaEvent_t after_response, const void* level0_tmap, int tma_rows_per_job) {
    (void)h_jobs;
    if (n_jobs <= 0) return DVFE_OK;
    int mi = 0;
#define GFTT_MARK() do { if (marks) cudaEventRecord(marks[mi++], st); } while (0)
    const bool ext = h_jobs != nullptr && h_jobs[0].eig_in != nullptr;     // seam op with an external response map
    if (ext) {
        // the materialised detection mask is only needed by the kernels that read an external response map
        dim3 blk(32, 8), grid(((max_w + 15) / 16 + 31) / 32, (max_h + 7) / 8, n_jobs);
        DVFE_LAUNCH(k_gftt_mask_fill, grid, blk, 0, st, d_jobs);
    }
    GFTT_MARK();
    if (ext && max_pts > 0) {
        dim3 grid((max_pts + 3) / 4, n_jobs);
        DVFE_LAUNCH(k_gftt_discs, grid, 128, 0, st, d_jobs);
    }
    GFTT_MARK();
    if (ext) {
        dim3 blk(32, 8), grid((max_w + 31) / 32, (max_h + 7) / 8, n_jobs);
        DVFE_LAUNCH(k_gftt_max_ext, grid, blk, 0, st, d_jobs);
        DVFE_LAUNCH(k_gftt_candidates_ext, grid, blk, 0, st, d_jobs);
    } else {
        // fused: detection mask (region minus discs, built per strip in shared memory) + response + pre-candidates
        const int gx = (max_w + RS_COLS * RS_WARPS - 1) / (RS_COLS * RS_WARPS);
        const long warps = (long)gx * RS_WARPS * ((max_h + RS_ROWS - 1) / RS_ROWS) * n_jobs;
        const int rows = warps >= 148L * 16 ? RS_ROWS : RS_ROWS_SMALL;       // fewer than 16 warps per SM: shorter strips
        dim3 blk(RS_WARPS * 32), grid(gx, (max_h + rows - 1) / rows, n_jobs);
        CUtensorMap tm;
        memset(&tm, 0, sizeof(tm));
        if (level0_tmap != nullptr) memcpy(&tm, level0_tmap, sizeof(tm));
        DVFE_LAUNCH(k_gftt_response, grid, blk, 0, st, d_jobs, rows, tm, level0_tmap != nullptr ? 1 : 0, tma_rows_per_job);
    }
    GFTT_MARK();
    if (after_response) cudaEventRecord(after_response, st);
    DVFE_CHECK_RC(gftt_prepare_device());
    DVFE_LAUNCH(k_gftt_select, n_jobs, NMS_THREADS, NMS_DYN_SMEM, st, d_jobs);
    DVFE_LAUNCH(k_gftt_assign_ids, (n_jobs + 127) / 128, 128, 0, st, d_jobs, n_jobs);
#undef GFTT_MARK
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libdvfe does not link libcuda)
int dvfe_make_level0_tmap(void* out, const uint8_t* base, int pitch, long n_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            dvfe_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return DVFE_ERR_CUDA;
        }
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};                       // bytes between rows (dimension 1)
    const cuuint32_t box[2] = {RS_TILE_PITCH, RS_TILE_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dvfe_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DVFE_ERR_CUDA; }
    return DVFE_OK;
}

int launch_min_eigen_val(const uint8_t* img, int pitch, int w, int h, float* eig, cudaStream_t st) {
    dim3 blk(RS_WARPS * 32), grid((w + RS_COLS * RS_WARPS - 1) / (RS_COLS * RS_WARPS), (h + RS_ROWS - 1) / RS_ROWS);
    DVFE_LAUNCH(k_min_eigen_val, grid, blk, 0, st, img, pitch, w, h, eig);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

int launch_disc_mask(uint8_t* mask, int pitch, int w, int h, const float2* pts, const int* n, int max_pts, int radius,
                     cudaStream_t st) {
    if (max_pts <= 0) return DVFE_OK;
    DVFE_LAUNCH(k_disc_mask_op, (max_pts + 3) / 4, 128, 0, st, mask, pitch, w, h, pts, n, radius);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
