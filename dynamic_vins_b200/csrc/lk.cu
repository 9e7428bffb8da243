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
#include "kernels.cuh"

#ifndef LK_WARPS
#define LK_WARPS 4
#endif
#ifndef LK_MIN_BLOCKS
#define LK_MIN_BLOCKS 6          // resident blocks per SM the register allocation targets
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

#define LK_WIN_BYTES (24 * 24 + 16)

// Template of one run (7 pixels of window row `ry`, columns x0..x0+6) from the staged 24x24 window:
// Scharr derivatives at the 8x2 taps the run's bilinear samples touch, computed with dp4a on 4-byte windows
// (row filter and the vertical 3/10/3 resp. -1/+1 weights folded into the int8 tap weights), then the 14-bit
// fixed-point bilinear samples of I, Ix, Iy.
__device__ __forceinline__ void lk_template_run(const uint8_t* __restrict__ win, int ry, int x0, bool valid, bool interior,
                                                int ipx, int ipy, int lw, int lh, int iw00, int iw01, int iw10, int iw11,
                                                int (&Ix)[LK_RUN], int (&Iy)[LK_RUN], int& c1, int& c2, int& sA11,
                                                int& sA12, int& sA22) {
    int gx0[8], gx1[8], gy0[8], gy1[8];      // d/dx and d/dy at tap rows ry, ry+1 ; tap columns x0..x0+7
#pragma unroll
    for (int j = 0; j < 8; j++) { gx0[j] = 0; gx1[j] = 0; gy0[j] = 0; gy1[j] = 0; }
    unsigned Wa[8], Wb[8];                   // 4-byte windows of image rows ry+1 (= window row of the pixels) and ry+2
    const int sh = (x0 & 3) * 8;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        // image row (ipy - 1 + ry + r): 12 bytes from window column x0 (= image column ipx - 1 + x0)
        const unsigned* wp = reinterpret_cast<const unsigned*>(win + (ry + r) * 24 + (x0 & ~3));
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
    if (!interior) {
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
        if (!valid) { ixv = 0; iyv = 0; }
        Ix[j] = ixv; Iy[j] = iyv;
        c1 += ival * ixv; c2 += ival * iyv;
        sA11 += ixv * ixv; sA12 += ixv * iyv; sA22 += iyv * iyv;
    }
}

__global__ void __launch_bounds__(LK_WARPS * 32, LK_MIN_BLOCKS) k_lk_track(const LkGroup* __restrict__ groups, int max_level, int flow_back,
                                                                                    int back_max_level, double fb_threshold, int tcache_flags) {
    __shared__ __align__(16) uint8_t s_win[LK_WARPS][LK_WIN_BYTES];
    const LkGroup& G = groups[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * LK_WARPS + warp;
    if (i >= *G.n) return;
    uint8_t* __restrict__ win = s_win[warp];
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
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: forward  img1 -> img2 from p1;  pass 1: backward img2 -> img1 from p2, initial guess p1
        const uint8_t* __restrict__ pyrI = pass ? G.pyrB : G.pyrA;
        const uint8_t* __restrict__ pyrJ = pass ? G.pyrA : G.pyrB;
        const float2 src = pass ? p2 : p1;
        const int lmax = pass ? (back_max_level < top ? back_max_level : top) : (max_level < top ? max_level : top);
        float outx = pass ? p1.x : 0.f, outy = pass ? p1.y : 0.f;      // nextPts[ptidx]
        int st = 1;
#pragma unroll 1
        for (int level = lmax; level >= 0; --level) {
            const PyrLevel L = G.desc.lv[level];
            const uint8_t* __restrict__ Ipx = pyrI + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX;
            const uint8_t* __restrict__ Jpx = pyrJ + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX;
            const float scale = __int_as_float((127 - level) << 23);      // (float)(1./(1 << level))
            float prevx = src.x * scale, prevy = src.y * scale;
            float nextx, nexty;
            if (level == lmax) {
                if (pass) { nextx = outx * scale; nexty = outy * scale; }   // OPTFLOW_USE_INITIAL_FLOW
                else { nextx = prevx; nexty = prevy; }
            } else {
                nextx = outx * 2.f; nexty = outy * 2.f;
            }
            outx = nextx; outy = nexty;

            prevx -= DVFE_HALF_WIN; prevy -= DVFE_HALF_WIN;
            const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
            if (ipx < -DVFE_WIN || ipx >= L.w || ipy < -DVFE_WIN || ipy >= L.h) {
                if (level == 0) st = 0;
                continue;
            }
            int Ix0[LK_RUN], Iy0[LK_RUN], Ix1[LK_RUN], Iy1[LK_RUN];
            int c1 = 0, c2 = 0;          // sum I*Ix, sum I*Iy over this lane's pixels (constant over the iterations)
            float A11, A12, A22;
            // The forward template of a stereo call (current left image at the current points) is bit for bit the
            // forward template of the next temporal call (same image, now `prev`, same points): the stereo call
            // stores it, the temporal call loads it instead of rebuilding it (4 of the 12 templates of a frame).
            unsigned* __restrict__ tc = (pass == 0 && G.tcache != nullptr && level < DVFE_MAX_PYR_LEVELS)
                                            ? G.tcache + ((size_t)i * DVFE_MAX_PYR_LEVELS + level) * LK_TCACHE_WORDS + lane : nullptr;
            if (tc != nullptr && (tcache_flags & LK_TCACHE_READ)) {
#pragma unroll
                for (int j = 0; j < LK_RUN; j++) {
                    const int w0 = (int)tc[j * 32], w1 = (int)tc[(LK_RUN + j) * 32];
                    Ix0[j] = (w0 << 16) >> 16; Iy0[j] = w0 >> 16;
                    Ix1[j] = (w1 << 16) >> 16; Iy1[j] = w1 >> 16;
                }
                c1 = (int)tc[14 * 32]; c2 = (int)tc[15 * 32];
                const unsigned m = tc[16 * 32];
                A11 = __uint_as_float(__shfl_sync(0xffffffffu, m, 0));
                A12 = __uint_as_float(__shfl_sync(0xffffffffu, m, 1));
                A22 = __uint_as_float(__shfl_sync(0xffffffffu, m, 2));
            } else {
                // ---- stage the 24x24 window of I around the patch (rows ipy-1.., columns ipx-1..) ----
                __syncwarp();
                for (int t = lane; t < 24 * 6; t += 32) {          // 24 rows x 6 words, rows are 4-byte aligned in smem
                    const int r = t / 6, c4 = (t - r * 6) * 4;
                    *reinterpret_cast<unsigned*>(win + r * 24 + c4) = load4(Ipx + (ipy - 1 + r) * L.pitch, ipx - 1 + c4);
                }
                __syncwarp();

                const float a = prevx - (float)ipx, b = prevy - (float)ipy;
                int iw00, iw01, iw10, iw11;
                lk_weights(a, b, iw00, iw01, iw10, iw11);
                const bool interior = ipx >= 0 && ipy >= 0 && ipx + 22 <= L.w && ipy + 22 <= L.h;
                int sA11 = 0, sA12 = 0, sA22 = 0;
                lk_template_run(win, ry0, rx0, true, interior, ipx, ipy, L.w, L.h, iw00, iw01, iw10, iw11, Ix0, Iy0, c1, c2,
                                sA11, sA12, sA22);
                lk_template_run(win, ry1, rx1, has1, interior, ipx, ipy, L.w, L.h, iw00, iw01, iw10, iw11, Ix1, Iy1, c1, c2,
                                sA11, sA12, sA22);
                A11 = __ll2float_rn(warp_sum_i64(sA11)) * FLT_SCALE;
                A12 = __ll2float_rn(warp_sum_i64(sA12)) * FLT_SCALE;
                A22 = __ll2float_rn(warp_sum_i64(sA22)) * FLT_SCALE;
                if (tc != nullptr && (tcache_flags & LK_TCACHE_WRITE)) {
#pragma unroll
                    for (int j = 0; j < LK_RUN; j++) {
                        tc[j * 32] = __byte_perm((unsigned)Ix0[j], (unsigned)Iy0[j], 0x5410);
                        tc[(LK_RUN + j) * 32] = __byte_perm((unsigned)Ix1[j], (unsigned)Iy1[j], 0x5410);
                    }
                    tc[14 * 32] = (unsigned)c1; tc[15 * 32] = (unsigned)c2;
                    tc[16 * 32] = __float_as_uint(lane == 0 ? A11 : (lane == 1 ? A12 : A22));
                }
            }
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * DVFE_WIN * DVFE_WIN);
            if ((double)minEig < 1e-4 || D < 1.1920928955078125e-07f) {
                if (level == 0) st = 0;
                continue;
            }
            D = 1.f / D;
            nextx -= DVFE_HALF_WIN; nexty -= DVFE_HALF_WIN;
            const int roff0 = ry0 * L.pitch, roff1 = ry1 * L.pitch;
            float pdx = 0.f, pdy = 0.f;
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
                const uint8_t* __restrict__ Jw = Jpx + iny * L.pitch;
                // sum (J - I) * Ix = sum J * Ix - sum I * Ix
                int sb1 = -c1, sb2 = -c2;
                {
                    unsigned A_lo, A_hi, B_lo, B_hi;
                    load8(Jw + roff0, inx + rx0, A_lo, A_hi);
                    load8(Jw + roff0 + L.pitch, inx + rx0, B_lo, B_hi);
                    LK_RUN_MAC(A_lo, A_hi, B_lo, B_hi, Ix0, Iy0);
                }
                {
                    unsigned A_lo, A_hi, B_lo, B_hi;
                    load8(Jw + roff1, inx + rx1, A_lo, A_hi);
                    load8(Jw + roff1 + L.pitch, inx + rx1, B_lo, B_hi);
                    LK_RUN_MAC(A_lo, A_hi, B_lo, B_hi, Ix1, Iy1);
                }
                const float b1 = __ll2float_rn(warp_sum_i64(sb1)) * FLT_SCALE;
                const float b2 = __ll2float_rn(warp_sum_i64(sb2)) * FLT_SCALE;
                const float dx = (A12 * b2 - A22 * b1) * D;
                const float dy = (A12 * b1 - A11 * b2) * D;
                nextx += dx; nexty += dy;
                outx = nextx + DVFE_HALF_WIN; outy = nexty + DVFE_HALF_WIN;
                if ((double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01) break;
                if (it > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
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
              int back_max_level, double fb_threshold, int tcache_flags) {
    if (n_groups <= 0 || max_pts <= 0) return DVFE_OK;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, n_groups);
    DVFE_LAUNCH(k_lk_track, grid, LK_WARPS * 32, 0, st, d_groups, max_level, flow_back, back_max_level, fb_threshold, tcache_flags);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
