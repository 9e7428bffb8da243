// Pyramidal Lucas-Kanade with forward-backward check — replaces FeatureTrackByLK
// (dynamic_vins/src/front_end/feature_utils.cpp:35-69), i.e. two cv::calcOpticalFlowPyrLK calls
// (:43 forward, 21x21, maxLevel 3; :50-53 backward, maxLevel 1, OPTFLOW_USE_INITIAL_FLOW), the
// 0.5 px round-trip test (:55-60), InBorder (:63-66, feature_utils.h:68-74) and, when a region mask is
// given, the mask test of InstFeat::TrackLeft (front_end/instance_feature.cpp:166-171).
//
// Arithmetic follows cv::detail::LKTrackerInvoker (OpenCV 3.4.16 modules/video/src/lkpyramid.cpp,
// restated in SURVEY.md Appendix A and oracle/spec.c):  14-bit fixed-point bilinear weights,
// int16 template I (5 fractional bits) and Scharr derivatives, fp32 2x2 solve.  The normal-equation
// sums are accumulated EXACTLY in integers and converted to float once (OpenCV accumulates in float
// SIMD lanes; the exact sum is the value those approximate).  Compiled with -fmad=false: every float
// expression below must round exactly like the scalar C++ it restates.
//
// Mapping: one warp per point, all pyramid levels, forward then backward, in one launch.
//   * the 21x21 window is cut into 63 horizontal runs of 7 pixels; a lane owns runs `lane` and `lane+32`
//     (14 pixels), whose template values (I, Ix, Iy) stay in registers for all iterations of a level;
//   * per iteration a run needs 2 rows x 8 bytes of J: three aligned 32-bit loads per row, funnel-shifted to
//     the window origin, instead of 4 byte loads per pixel; the bilinear sample of a pixel is two dp2a (16-bit
//     weights x 8-bit pixels) on the row registers and their 1-byte-shifted copies (dp2a.lo / .hi select the byte pair);
//   * the Scharr derivatives are computed on the fly from a 24x24 u8 window staged in shared memory (the
//     reference materialises a 4 B/px derivative image per level and per call);
//   * the 2x2 sums are reduced exactly with redux.sync on 16-bit halves.
#include <limits.h>

#include "kernels.cuh"

#ifndef LK_WARPS
#define LK_WARPS 4
#endif
#ifndef LK_MIN_BLOCKS
#define LK_MIN_BLOCKS 6          // resident blocks per SM the register allocation targets
#endif
#ifndef LK_JCACHE
#define LK_JCACHE 1            // keep the J row registers while the window origin stays on the same integer pixel
#endif
#define LK_RUN 7                  // pixels per run; 3 runs per window row
#define W_BITS 14

__device__ __forceinline__ long long warp_sum_i64(int v) {
    // exact 64-bit sum of 32 int32 lanes with two 32-bit redux ops
    const int lo = v & 0xffff;
    const int hi = v >> 16;
    const int slo = __reduce_add_sync(0xffffffffu, lo);
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    return (long long)shi * 65536ll + (long long)slo;
}

// Exact sum of 32 int32 lanes as a float (one rounding), scaled.  When every lane is below 2^26 in magnitude the sum fits
// an int32: one redux and a 32-bit conversion; otherwise the two-halves path.  `small` is warp-uniform.
__device__ __forceinline__ float warp_sum_f(int v, bool small, float scale) {
    if (small) return __int2float_rn(__reduce_add_sync(0xffffffffu, v)) * scale;
    return __ll2float_rn(warp_sum_i64(v)) * scale;
}
__device__ __forceinline__ unsigned lk_big(int v) { return (unsigned)(v + (1 << 26)) >> 27; }      // != 0 iff v outside [-2^26, 2^26)

// 8 consecutive bytes at byte offset `o` from the 4-byte aligned `base` (any alignment of o): lo = bytes 0..3, hi = bytes 4..7
__device__ __forceinline__ void load8o(const uint8_t* __restrict__ base, int o, unsigned& lo, unsigned& hi) {
    const unsigned* wp = reinterpret_cast<const unsigned*>(base + (o & ~3));
    const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    const int sh = (o & 3) * 8;
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

__device__ __forceinline__ void lk_weights(float a, float b, int& iw00, int& iw01, int& iw10, int& iw11) {
    iw00 = __float2int_rn((1.f - a) * (1.f - b) * (float)(1 << W_BITS));
    iw01 = __float2int_rn(a * (1.f - b) * (float)(1 << W_BITS));
    iw10 = __float2int_rn((1.f - a) * b * (float)(1 << W_BITS));
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
}

// d = a.s16[0] * b.u8[2h] + a.s16[1] * b.u8[2h+1] + c     (h = 0: lo, 1: hi)
__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 8 consecutive bytes starting at (row pointer + x), any alignment: lo = bytes 0..3, hi = bytes 4..7
__device__ __forceinline__ void load8(const uint8_t* __restrict__ row, int x, unsigned& lo, unsigned& hi) {
    const unsigned* wp = reinterpret_cast<const unsigned*>(row + (x & ~3));
    const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    const int sh = (x & 3) * 8;
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

// 4 consecutive bytes starting at (row pointer + x), any alignment
__device__ __forceinline__ unsigned load4(const uint8_t* __restrict__ row, int x) {
    const unsigned* wp = reinterpret_cast<const unsigned*>(row + (x & ~3));
    return __funnelshift_r(__ldg(wp), __ldg(wp + 1), (x & 3) * 8);
}

// bilinear samples of the 7 pixels of a run from its two rows of 8 bytes (A: upper, B: lower) and MAC with the template:
// pixel j needs the byte pairs (A[j], A[j+1]) and (B[j], B[j+1]); dp2a.lo / dp2a.hi pick the pair at bytes 0-1 / 2-3 of a
// register, so the rows and their 1-byte-shifted copies serve all 7 pixels without assembling a tap word per pixel
#define LK_RUN_MAC(A_lo, A_hi, B_lo, B_hi, IX, IY)                                                         \
    {                                                                                                      \
        const unsigned A_m1 = __funnelshift_r(A_lo, A_hi, 8), B_m1 = __funnelshift_r(B_lo, B_hi, 8);       \
        const unsigned A_m2 = A_hi >> 8, B_m2 = B_hi >> 8;                                                 \
        const int rnd = 1 << (W_BITS - 5 - 1);                                                             \
        int v;                                                                                             \
        v = dp2a_lo_su(W23, B_lo, dp2a_lo_su(W01, A_lo, rnd)) >> (W_BITS - 5); sb1 += v * IX[0]; sb2 += v * IY[0]; \
        v = dp2a_lo_su(W23, B_m1, dp2a_lo_su(W01, A_m1, rnd)) >> (W_BITS - 5); sb1 += v * IX[1]; sb2 += v * IY[1]; \
        v = dp2a_hi_su(W23, B_lo, dp2a_hi_su(W01, A_lo, rnd)) >> (W_BITS - 5); sb1 += v * IX[2]; sb2 += v * IY[2]; \
        v = dp2a_hi_su(W23, B_m1, dp2a_hi_su(W01, A_m1, rnd)) >> (W_BITS - 5); sb1 += v * IX[3]; sb2 += v * IY[3]; \
        v = dp2a_lo_su(W23, B_hi, dp2a_lo_su(W01, A_hi, rnd)) >> (W_BITS - 5); sb1 += v * IX[4]; sb2 += v * IY[4]; \
        v = dp2a_lo_su(W23, B_m2, dp2a_lo_su(W01, A_m2, rnd)) >> (W_BITS - 5); sb1 += v * IX[5]; sb2 += v * IY[5]; \
        v = dp2a_hi_su(W23, B_hi, dp2a_hi_su(W01, A_hi, rnd)) >> (W_BITS - 5); sb1 += v * IX[6]; sb2 += v * IY[6]; \
    }

// d = sum_i a.u8[i] * b.s8[i] + c
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Staged template window: the 24 x 24 bytes around the patch (rows ipy-1.., columns ipx-1..) are copied as the 7 aligned
// words per row that contain them, so the copy is pure cp.async (no registers, no shifts); the consumer adds the column
// misalignment (ipx-1) & 3 to its own.  Pitch 7 words (odd: the 21 window rows fall into different banks).
#define LK_WIN_PITCH 28
#define LK_WIN_BYTES (24 * LK_WIN_PITCH + 16)

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// window origin of the template at `level` for source point `src` (cv::detail::LKTrackerInvoker: prevPt * scale - halfWin)
__device__ __forceinline__ bool lk_window_origin(float2 src, int level, int lw, int lh, float& prevx, float& prevy, int& ipx, int& ipy) {
    const float scale = __int_as_float((127 - level) << 23);      // (float)(1./(1 << level))
    prevx = src.x * scale - DVFE_HALF_WIN;
    prevy = src.y * scale - DVFE_HALF_WIN;
    ipx = __float2int_rd(prevx); ipy = __float2int_rd(prevy);
    return !(ipx < -DVFE_WIN || ipx >= lw || ipy < -DVFE_WIN || ipy >= lh);
}

// enqueue the asynchronous copy of one level's template window (all lanes; no wait)
__device__ __forceinline__ void lk_stage_window_async(uint8_t* __restrict__ win, const uint8_t* __restrict__ Ipx, int pitch, int ipx, int ipy,
                                                      int lane) {
    const uint8_t* __restrict__ src = Ipx + (ptrdiff_t)(ipy - 1) * pitch + ((ipx - 1) & ~3);
#pragma unroll
    for (int t0 = 0; t0 < 24 * 7; t0 += 32) {
        const int t = t0 + lane;
        if (t < 24 * 7) {
            const int r = t / 7, w4 = (t - r * 7) * 4;
            cp_async4(win + r * LK_WIN_PITCH + w4, src + r * pitch + w4);
        }
    }
}

// Template of one run (7 pixels of window row `ry`, columns x0..x0+6) from the staged 24x24 window:
// Scharr derivatives at the 8x2 taps the run's bilinear samples touch, computed with dp4a on 4-byte windows
// (row filter and the vertical 3/10/3 resp. -1/+1 weights folded into the int8 tap weights), then the 14-bit
// fixed-point bilinear samples of I, Ix, Iy.
template <bool INTERIOR>
__device__ __forceinline__ void lk_template_run(const uint8_t* __restrict__ win, int ry, int x0, int xs,
                                                int ipx, int ipy, int lw, int lh, int iw00, int iw01, int iw10, int iw11,
                                                int (&Ix)[LK_RUN], int (&Iy)[LK_RUN], int& c1, int& c2, int& sA11,
                                                int& sA12, int& sA22) {
    int gx0[8], gx1[8], gy0[8], gy1[8];      // d/dx and d/dy at tap rows ry, ry+1 ; tap columns x0..x0+7
#pragma unroll
    for (int j = 0; j < 8; j++) { gx0[j] = 0; gx1[j] = 0; gy0[j] = 0; gy1[j] = 0; }
    unsigned Wa[8], Wb[8];                   // 4-byte windows of image rows ry+1 (= window row of the pixels) and ry+2
    const int sh = (xs & 3) * 8;           // xs = x0 + the staged window's column misalignment
#pragma unroll
    for (int r = 0; r < 4; r++) {
        // image row (ipy - 1 + ry + r): 12 bytes from window column x0 (= image column ipx - 1 + x0)
        const unsigned* wp = reinterpret_cast<const unsigned*>(win + (ry + r) * LK_WIN_PITCH + (xs & ~3));
        const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        unsigned R[3];
        R[0] = __funnelshift_r(w0, w1, sh); R[1] = __funnelshift_r(w1, w2, sh); R[2] = __funnelshift_r(w2, w3, sh);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned W = (j & 3) ? __funnelshift_r(R[j >> 2], R[(j >> 2) + 1], 8 * (j & 3)) : R[j >> 2];
            // bytes of W: I(x-1), I(x), I(x+1), - for tap column x = x0 + j of this image row
            // d/dx = 3*hd(y-1) + 10*hd(y) + 3*hd(y+1), hd = I(x+1) - I(x-1);  d/dy = hs(y+1) - hs(y-1), hs = 3,10,3
            if (r == 0) { gx0[j] = dp4a_us(W, 0x000300FD, gx0[j]); gy0[j] = dp4a_us(W, 0x00FDF6FD, gy0[j]); }
            if (r == 1) { gx0[j] = dp4a_us(W, 0x000A00F6, gx0[j]); gx1[j] = dp4a_us(W, 0x000300FD, gx1[j]);
                          gy1[j] = dp4a_us(W, 0x00FDF6FD, gy1[j]); Wa[j] = W; }
            if (r == 2) { gx0[j] = dp4a_us(W, 0x000300FD, gx0[j]); gx1[j] = dp4a_us(W, 0x000A00F6, gx1[j]);
                          gy0[j] = dp4a_us(W, 0x00030A03, gy0[j]); Wb[j] = W; }
            if (r == 3) { gx1[j] = dp4a_us(W, 0x000300FD, gx1[j]); gy1[j] = dp4a_us(W, 0x00030A03, gy1[j]); }
        }
    }
    if (!INTERIOR) {
        // the derivative image has a ZERO border (cv::copyMakeBorder BORDER_CONSTANT): taps outside the image are 0
        const bool r0 = (unsigned)(ipy + ry) < (unsigned)lh, r1 = (unsigned)(ipy + ry + 1) < (unsigned)lh;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const bool cv = (unsigned)(ipx + x0 + j) < (unsigned)lw;
            if (!(cv && r0)) { gx0[j] = 0; gy0[j] = 0; }
            if (!(cv && r1)) { gx1[j] = 0; gy1[j] = 0; }
        }
    }
    const int W01 = (iw00 & 0xffff) | (iw01 << 16);
    const int W23 = (iw10 & 0xffff) | (iw11 << 16);
#pragma unroll
    for (int j = 0; j < LK_RUN; j++) {
        // pixel (x0 + j, ry): I taps are bytes 1,2 of the windows at column j of image rows ry+1 / ry+2 of the window
        const unsigned T = __byte_perm(Wa[j], Wb[j], 0x6521);
        const int ival = dp2a_hi_su(W23, T, dp2a_lo_su(W01, T, 1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
        int ixv = (gx0[j] * iw00 + gx0[j + 1] * iw01 + gx1[j] * iw10 + gx1[j + 1] * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
        int iyv = (gy0[j] * iw00 + gy0[j + 1] * iw01 + gy1[j] * iw10 + gy1[j + 1] * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
        Ix[j] = ixv; Iy[j] = iyv;
        c1 += ival * ixv; c2 += ival * iyv;
        sA11 += ixv * ixv; sA12 += ixv * iyv; sA22 += iyv * iyv;
    }
}

__global__ void __launch_bounds__(LK_WARPS * 32, LK_MIN_BLOCKS) k_lk_track(const LkGroup* __restrict__ groups, int max_level, int flow_back,
                                                                                    int back_max_level, double fb_threshold, int tcache_flags,
                                                                                    int reuse_max_level) {
    __shared__ __align__(16) uint8_t s_win[LK_WARPS][2][LK_WIN_BYTES];      // two windows per warp: this level / the next one
    const LkGroup& G = groups[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * LK_WARPS + warp;
    if (i >= *G.n) return;
    const float FLT_SCALE = 1.f / (1 << 20);

    // this lane's two runs: run r -> window row r / 3, first column 7 * (r % 3)
    const int ry0 = lane / 3, rx0 = (lane - ry0 * 3) * LK_RUN;
    const int r1 = lane + 32;
    const bool has1 = r1 < 63;
    const int ry1 = has1 ? r1 / 3 : 0, rx1 = has1 ? (r1 - (r1 / 3) * 3) * LK_RUN : 0;

    float2 p1 = G.ptsA[i];
    p1.x += G.offx; p1.y += G.offy;
    const int top = G.desc.n_levels - 1;

    float2 p2 = make_float2(0.f, 0.f), rev = p1;
    int status = 1;
    // index of this point in the temporal call of the same step, whose backward pass left its templates in tcache_bwd
    const int old_i = ((tcache_flags & LK_TCACHE_READ_BWD) && G.old_idx != nullptr && G.tcache_bwd != nullptr) ? G.old_idx[i] : -1;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: forward  img1 -> img2 from p1;  pass 1: backward img2 -> img1 from p2, initial guess p1
        const uint8_t* __restrict__ pyrI = pass ? G.pyrB : G.pyrA;
        const uint8_t* __restrict__ pyrJ = pass ? G.pyrA : G.pyrB;
        const float2 src = pass ? p2 : p1;
        const int lmax = pass ? (back_max_level < top ? back_max_level : top) : (max_level < top ? max_level : top);
        float outx = pass ? p1.x : 0.f, outy = pass ? p1.y : 0.f;      // nextPts[ptidx]
        int st = 1;
        // The template windows of a pass depend only on the source point: the window of the next level is fetched with
        // cp.async while this level iterates, so only the first window of a pass is waited for.
        const bool stage = !(pass == 0 && G.tcache != nullptr && (tcache_flags & LK_TCACHE_READ));
        // forward levels whose template comes from the temporal call's backward pass need no window either
        const int cached_below = (pass == 0 && old_i >= 0) ? reuse_max_level : -1;
        if (stage) {
            const PyrLevel L = G.desc.lv[lmax];
            float fx, fy; int wx, wy;
            __syncwarp();
            if (lmax > cached_below && lk_window_origin(src, lmax, L.w, L.h, fx, fy, wx, wy))
                lk_stage_window_async(s_win[warp][0], pyrI + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX, L.pitch, wx, wy, lane);
            cp_async_commit();
        } else if (lane < 17) {
            // cached templates: 17 lines of 128 B per level; pull the lower levels towards L2 while the top level iterates
            const unsigned* __restrict__ tb = G.tcache + (size_t)i * DVFE_MAX_PYR_LEVELS * LK_TCACHE_WORDS + lane * 32;
            for (int l = lmax - 1; l >= 0; --l) asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + l * LK_TCACHE_WORDS));
        }
#pragma unroll 1
        for (int level = lmax; level >= 0; --level) {
            const PyrLevel L = G.desc.lv[level];
            const uint8_t* __restrict__ Ipx = pyrI + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX;
            const uint8_t* __restrict__ Jpx = pyrJ + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX;
            const float scale = __int_as_float((127 - level) << 23);      // (float)(1./(1 << level))
            uint8_t* __restrict__ win = s_win[warp][(lmax - level) & 1];
            if (stage) {
                if (level > 0) {
                    const PyrLevel Ln = G.desc.lv[level - 1];
                    float fx, fy; int wx, wy;
                    __syncwarp();                // every lane has finished reading the buffer the copy below overwrites
                    if (level - 1 > cached_below && lk_window_origin(src, level - 1, Ln.w, Ln.h, fx, fy, wx, wy))
                        lk_stage_window_async(s_win[warp][(lmax - level + 1) & 1], pyrI + Ln.offset + (size_t)DVFE_PADY * Ln.pitch + DVFE_PADX,
                                              Ln.pitch, wx, wy, lane);
                    cp_async_commit();
                    cp_async_wait<1>();          // everything but the group just committed: this level's window is in
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
            }
            float prevx, prevy;
            int ipx, ipy;
            const bool in_range = lk_window_origin(src, level, L.w, L.h, prevx, prevy, ipx, ipy);
            float nextx, nexty;
            if (level == lmax) {
                if (pass) { nextx = outx * scale; nexty = outy * scale; }   // OPTFLOW_USE_INITIAL_FLOW
                else { nextx = src.x * scale; nexty = src.y * scale; }
            } else {
                nextx = outx * 2.f; nexty = outy * 2.f;
            }
            outx = nextx; outy = nexty;
            if (!in_range) {
                if (level == 0) st = 0;
                continue;
            }
            int Ix0[LK_RUN], Iy0[LK_RUN], Ix1[LK_RUN], Iy1[LK_RUN];
            int c1 = 0, c2 = 0;          // sum I*Ix, sum I*Iy over this lane's pixels (constant over the iterations)
            float A11, A12, A22;
            // The forward template of a stereo call (current left image at the current points) is bit for bit the
            // forward template of the next temporal call (same image, now `prev`, same points): the stereo call
            // stores it, the temporal call loads it instead of rebuilding it (4 of the 12 templates of a frame).
            // Where this (pass, level)'s template comes from / goes to.  src_tc: load instead of building; dst_tc: store.
            const size_t tco = ((size_t)i * DVFE_MAX_PYR_LEVELS + level) * LK_TCACHE_WORDS + lane;
            const unsigned* __restrict__ src_tc = nullptr;
            unsigned* __restrict__ dst_tc = nullptr;
            if (pass == 0) {
                if (G.tcache != nullptr && (tcache_flags & LK_TCACHE_READ)) src_tc = G.tcache + tco;
                else if (level <= cached_below)
                    src_tc = G.tcache_bwd + ((size_t)old_i * DVFE_MAX_PYR_LEVELS + level) * LK_TCACHE_WORDS + lane;
                if (G.tcache != nullptr && (tcache_flags & LK_TCACHE_WRITE)) dst_tc = G.tcache + tco;
            } else if (G.tcache_bwd != nullptr && (tcache_flags & LK_TCACHE_WRITE_BWD)) {
                dst_tc = G.tcache_bwd + tco;
            }
            if (src_tc != nullptr) {
#pragma unroll
                for (int j = 0; j < LK_RUN; j++) {
                    const int w0 = (int)src_tc[j * 32], w1 = (int)src_tc[(LK_RUN + j) * 32];
                    Ix0[j] = (w0 << 16) >> 16; Iy0[j] = w0 >> 16;
                    Ix1[j] = (w1 << 16) >> 16; Iy1[j] = w1 >> 16;
                }
                c1 = (int)src_tc[14 * 32]; c2 = (int)src_tc[15 * 32];
                const unsigned m = src_tc[16 * 32];
                A11 = __uint_as_float(__shfl_sync(0xffffffffu, m, 0));
                A12 = __uint_as_float(__shfl_sync(0xffffffffu, m, 1));
                A22 = __uint_as_float(__shfl_sync(0xffffffffu, m, 2));
            } else {
                const int wm = (ipx - 1) & 3;          // column misalignment of the staged window
                const float a = prevx - (float)ipx, b = prevy - (float)ipy;
                int iw00, iw01, iw10, iw11;
                lk_weights(a, b, iw00, iw01, iw10, iw11);
                // warp-uniform: the patch and its derivative taps lie inside the image (no zero-border rule to apply); the two
                // variants are separate code so that interior patches, nearly all of them, skip the 64 border selects
                const bool interior = ipx >= 0 && ipy >= 0 && ipx + 22 <= L.w && ipy + 22 <= L.h;
                int sA11 = 0, sA12 = 0, sA22 = 0;
                // lane 31 owns one run only: its second run is computed with zero weights, which makes I, Ix and Iy zero
                const int v00 = has1 ? iw00 : 0, v01 = has1 ? iw01 : 0, v10 = has1 ? iw10 : 0, v11 = has1 ? iw11 : 0;
                if (interior) {
                    lk_template_run<true>(win, ry0, rx0, rx0 + wm, ipx, ipy, L.w, L.h, iw00, iw01, iw10, iw11, Ix0, Iy0, c1, c2, sA11, sA12, sA22);
                    lk_template_run<true>(win, ry1, rx1, rx1 + wm, ipx, ipy, L.w, L.h, v00, v01, v10, v11, Ix1, Iy1, c1, c2, sA11, sA12, sA22);
                } else {
                    lk_template_run<false>(win, ry0, rx0, rx0 + wm, ipx, ipy, L.w, L.h, iw00, iw01, iw10, iw11, Ix0, Iy0, c1, c2, sA11, sA12, sA22);
                    lk_template_run<false>(win, ry1, rx1, rx1 + wm, ipx, ipy, L.w, L.h, v00, v01, v10, v11, Ix1, Iy1, c1, c2, sA11, sA12, sA22);
                }
                {
                    const bool small = !__any_sync(0xffffffffu, (lk_big(sA11) | lk_big(sA12) | lk_big(sA22)) != 0u);
                    A11 = warp_sum_f(sA11, small, FLT_SCALE);
                    A12 = warp_sum_f(sA12, small, FLT_SCALE);
                    A22 = warp_sum_f(sA22, small, FLT_SCALE);
                }
            }
            if (dst_tc != nullptr && dst_tc != src_tc) {
#pragma unroll
                for (int j = 0; j < LK_RUN; j++) {
                    dst_tc[j * 32] = __byte_perm((unsigned)Ix0[j], (unsigned)Iy0[j], 0x5410);
                    dst_tc[(LK_RUN + j) * 32] = __byte_perm((unsigned)Ix1[j], (unsigned)Iy1[j], 0x5410);
                }
                dst_tc[14 * 32] = (unsigned)c1; dst_tc[15 * 32] = (unsigned)c2;
                dst_tc[16 * 32] = __float_as_uint(lane == 0 ? A11 : (lane == 1 ? A12 : A22));
            }
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * DVFE_WIN * DVFE_WIN);
            if ((double)minEig < 1e-4 || D < 1.1920928955078125e-07f) {
                if (level == 0) st = 0;
                continue;
            }
            D = 1.f / D;
            nextx -= DVFE_HALF_WIN; nexty -= DVFE_HALF_WIN;
            // byte offsets of this lane's two runs inside the search window (rows are 16-byte aligned, so the alignment of a
            // linear offset is the alignment of its column)
            const int off0 = ry0 * L.pitch + rx0, off1 = ry1 * L.pitch + rx1;
            float pdx = 0.f, pdy = 0.f;
#if LK_JCACHE
            // The bytes of J an iteration reads depend only on the integer part of the window origin.  Once the Gauss-Newton
            // steps fall below a pixel, consecutive iterations mostly stay on the same integer origin and differ in the
            // bilinear weights only: the 8 row registers are kept, the gather (12 loads + address arithmetic + funnel shifts)
            // and its latency are skipped.  Same bytes, same arithmetic: bit-identical.
            int c_inx = INT_MIN, c_iny = INT_MIN;
            unsigned A0_lo = 0, A0_hi = 0, B0_lo = 0, B0_hi = 0, A1_lo = 0, A1_hi = 0, B1_lo = 0, B1_hi = 0;
#endif
#pragma unroll 1
            for (int it = 0; it < 30; it++) {
                const int inx = __float2int_rd(nextx), iny = __float2int_rd(nexty);
                if (inx < -DVFE_WIN || inx >= L.w || iny < -DVFE_WIN || iny >= L.h) {
                    if (level == 0) st = 0;
                    break;
                }
                const float a = nextx - (float)inx, b = nexty - (float)iny;
                int iw00, iw01, iw10, iw11;
                lk_weights(a, b, iw00, iw01, iw10, iw11);
                const int W01 = (iw00 & 0xffff) | (iw01 << 16);
                const int W23 = (iw10 & 0xffff) | (iw11 << 16);
                // sum (J - I) * Ix = sum J * Ix - sum I * Ix
                int sb1 = -c1, sb2 = -c2;
#if LK_JCACHE
                if (inx != c_inx || iny != c_iny) {
                    const int o = iny * L.pitch + inx;
                    load8o(Jpx, o + off0, A0_lo, A0_hi);
                    load8o(Jpx, o + off0 + L.pitch, B0_lo, B0_hi);
                    load8o(Jpx, o + off1, A1_lo, A1_hi);
                    load8o(Jpx, o + off1 + L.pitch, B1_lo, B1_hi);
                    c_inx = inx; c_iny = iny;
                }
                LK_RUN_MAC(A0_lo, A0_hi, B0_lo, B0_hi, Ix0, Iy0);
                LK_RUN_MAC(A1_lo, A1_hi, B1_lo, B1_hi, Ix1, Iy1);
#else
                const int o = iny * L.pitch + inx;
                {
                    unsigned A_lo, A_hi, B_lo, B_hi;
                    load8o(Jpx, o + off0, A_lo, A_hi);
                    load8o(Jpx, o + off0 + L.pitch, B_lo, B_hi);
                    LK_RUN_MAC(A_lo, A_hi, B_lo, B_hi, Ix0, Iy0);
                }
                {
                    unsigned A_lo, A_hi, B_lo, B_hi;
                    load8o(Jpx, o + off1, A_lo, A_hi);
                    load8o(Jpx, o + off1 + L.pitch, B_lo, B_hi);
                    LK_RUN_MAC(A_lo, A_hi, B_lo, B_hi, Ix1, Iy1);
                }
#endif
                const bool small = !__any_sync(0xffffffffu, (lk_big(sb1) | lk_big(sb2)) != 0u);
                const float b1 = warp_sum_f(sb1, small, FLT_SCALE);
                const float b2 = warp_sum_f(sb2, small, FLT_SCALE);
                const float dx = (A12 * b2 - A22 * b1) * D;
                const float dy = (A12 * b1 - A11 * b2) * D;
                nextx += dx; nexty += dy;
                outx = nextx + DVFE_HALF_WIN; outy = nexty + DVFE_HALF_WIN;
                if ((double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01) break;
                // std::abs(delta.x + prevDelta.x) < 0.01 (float sum against the double 0.01): the floats below 0.01 are exactly
                // those below the first float above it, 0x3C23D70B
                if (it > 0 && fabsf(dx + pdx) < __int_as_float(0x3C23D70B) && fabsf(dy + pdy) < __int_as_float(0x3C23D70B)) {
                    outx -= dx * 0.5f; outy -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (st && level == 0) {
                const int qx = __float2int_rd(outx - DVFE_HALF_WIN), qy = __float2int_rd(outy - DVFE_HALF_WIN);
                if (qx < -DVFE_WIN || qx >= L.w || qy < -DVFE_WIN || qy >= L.h) st = 0;
            }
        }
        if (pass == 0) {
            p2 = make_float2(outx, outy);
            status = st;
            if (!flow_back || !st) break;      // the backward result cannot change a failed status
        } else {
            rev = make_float2(outx, outy);
            const float ddx = p1.x - rev.x, ddy = p1.y - rev.y;
            const float dist = sqrtf(ddx * ddx + ddy * ddy);
            status = (st && (double)dist <= fb_threshold) ? 1 : 0;
        }
    }
    if (status) {
        const int W = G.desc.lv[0].w, H = G.desc.lv[0].h;
        const int rx = __float2int_rn(p2.x), ry = __float2int_rn(p2.y);
        if (!(1 <= rx && rx < W - 1 && 1 <= ry && ry < H - 1)) status = 0;                        // InBorder
        else if (G.mask != nullptr && G.mask[(size_t)ry * G.mask_pitch + rx] == 0) status = 0;   // region mask
    }
    if (lane == 0) {
        G.ptsB[i] = p2;
        G.status[i] = (uint8_t)status;
        if (G.rev) G.rev[i] = rev;
    }
}

int launch_lk(const LkGroup* d_groups, int n_groups, int max_pts, int max_level, int flow_back, cudaStream_t st,
              int back_max_level, double fb_threshold, int tcache_flags, int reuse_max_level) {
    if (n_groups <= 0 || max_pts <= 0) return DVFE_OK;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, n_groups);
    DVFE_LAUNCH(k_lk_track, grid, LK_WARPS * 32, 0, st, d_groups, max_level, flow_back, back_max_level, fb_threshold, tcache_flags,
                reuse_max_level);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
