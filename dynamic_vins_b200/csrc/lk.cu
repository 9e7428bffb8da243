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
//     the window origin, instead of 4 byte loads per pixel; the four bilinear taps of a pixel are gathered
//     with one PRMT and reduced with two dp2a (16-bit weights x 8-bit pixels);
//   * the Scharr derivatives are computed on the fly from a 24x24 u8 window staged in shared memory (the
//     reference materialises a 4 B/px derivative image per level and per call);
//   * the 2x2 sums are reduced exactly with redux.sync on 16-bit halves.
#include "kernels.cuh"

#define LK_WARPS 4
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

// the 7 tap words (t00, t01, t10, t11) of a run from its two rows of 8 bytes
#define LK_TAPS(A_lo, A_hi, B_lo, B_hi, T)                                   \
    {                                                                        \
        const unsigned A_mid = __funnelshift_r(A_lo, A_hi, 16);              \
        const unsigned B_mid = __funnelshift_r(B_lo, B_hi, 16);              \
        T[0] = __byte_perm(A_lo, B_lo, 0x5410);                              \
        T[1] = __byte_perm(A_lo, B_lo, 0x6521);                              \
        T[2] = __byte_perm(A_lo, B_lo, 0x7632);                              \
        T[3] = __byte_perm(A_mid, B_mid, 0x6521);                            \
        T[4] = __byte_perm(A_hi, B_hi, 0x5410);                              \
        T[5] = __byte_perm(A_hi, B_hi, 0x6521);                              \
        T[6] = __byte_perm(A_hi, B_hi, 0x7632);                              \
    }

__global__ void __launch_bounds__(LK_WARPS * 32, 4) k_lk_track(const LkGroup* __restrict__ groups, int max_level, int flow_back) {
    __shared__ __align__(16) uint8_t s_win[LK_WARPS][24 * 24];
    __shared__ __align__(16) short2 s_der[LK_WARPS][22 * 22];
    const LkGroup& G = groups[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * LK_WARPS + warp;
    if (i >= *G.n) return;
    uint8_t* __restrict__ win = s_win[warp];
    short2* __restrict__ der = s_der[warp];
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
        const int lmax = pass ? (1 < top ? 1 : top) : (max_level < top ? max_level : top);
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
            // ---- stage the 24x24 window of I around the patch; Scharr taps (zero outside the image) ----
            __syncwarp();
            for (int t = lane; t < 24 * 6; t += 32) {          // 24 rows x 6 words, rows are 4-byte aligned in smem
                const int r = t / 6, c4 = (t - r * 6) * 4;
                *reinterpret_cast<unsigned*>(win + r * 24 + c4) = load4(Ipx + (ipy - 1 + r) * L.pitch, ipx - 1 + c4);
            }
            __syncwarp();
            for (int t = lane; t < 22 * 22; t += 32) {
                const int r = t / 22, c = t - r * 22;
                const int gx = ipx + c, gy = ipy + r;
                short2 d = make_short2(0, 0);
                if (gx >= 0 && gx < L.w && gy >= 0 && gy < L.h) {
                    const uint8_t* w0 = win + r * 24 + c;
                    const int a00 = w0[0], a01 = w0[1], a02 = w0[2];
                    const int a10 = w0[24], a12 = w0[26];
                    const int a20 = w0[48], a21 = w0[49], a22 = w0[50];
                    const int t0m = 3 * (a00 + a20) + 10 * a10, t0p = 3 * (a02 + a22) + 10 * a12;
                    const int t1m = a20 - a00, t1c = a21 - a01, t1p = a22 - a02;
                    d.x = (short)(t0p - t0m);
                    d.y = (short)(3 * (t1p + t1m) + 10 * t1c);
                }
                der[t] = d;
            }
            __syncwarp();

            float a = prevx - (float)ipx, b = prevy - (float)ipy;
            int iw00, iw01, iw10, iw11;
            lk_weights(a, b, iw00, iw01, iw10, iw11);

            int Iw[2][LK_RUN], Ix[2][LK_RUN], Iy[2][LK_RUN];
            int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int y = q ? ry1 : ry0, x0 = q ? rx1 : rx0;
                const bool valid = q ? has1 : true;
#pragma unroll
                for (int j = 0; j < LK_RUN; j++) {
                    const int x = x0 + j;
                    const uint8_t* w0 = win + (y + 1) * 24 + x + 1;
                    const int ival = (w0[0] * iw00 + w0[1] * iw01 + w0[24] * iw10 + w0[25] * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
                    const short2 d00 = der[y * 22 + x], d01 = der[y * 22 + x + 1];
                    const short2 d10 = der[(y + 1) * 22 + x], d11 = der[(y + 1) * 22 + x + 1];
                    int ixv = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                    int iyv = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                    if (!valid) { ixv = 0; iyv = 0; }
                    Iw[q][j] = ival; Ix[q][j] = ixv; Iy[q][j] = iyv;
                    sA11 += ixv * ixv; sA12 += ixv * iyv; sA22 += iyv * iyv;
                }
            }
            const float A11 = __ll2float_rn(warp_sum_i64(sA11)) * FLT_SCALE;
            const float A12 = __ll2float_rn(warp_sum_i64(sA12)) * FLT_SCALE;
            const float A22 = __ll2float_rn(warp_sum_i64(sA22)) * FLT_SCALE;
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
                a = nextx - (float)inx; b = nexty - (float)iny;
                lk_weights(a, b, iw00, iw01, iw10, iw11);
                const int W01 = (iw00 & 0xffff) | (iw01 << 16);
                const int W23 = (iw10 & 0xffff) | (iw11 << 16);
                const uint8_t* __restrict__ Jw = Jpx + iny * L.pitch;
                int sb1 = 0, sb2 = 0;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const uint8_t* rowA = Jw + (q ? roff1 : roff0);
                    const int x = inx + (q ? rx1 : rx0);
                    unsigned A_lo, A_hi, B_lo, B_hi, T[LK_RUN];
                    load8(rowA, x, A_lo, A_hi);
                    load8(rowA + L.pitch, x, B_lo, B_hi);
                    LK_TAPS(A_lo, A_hi, B_lo, B_hi, T);
#pragma unroll
                    for (int j = 0; j < LK_RUN; j++) {
                        const int v = dp2a_hi_su(W23, T[j], dp2a_lo_su(W01, T[j], 1 << (W_BITS - 5 - 1)));
                        const int diff = (v >> (W_BITS - 5)) - Iw[q][j];
                        sb1 += diff * Ix[q][j];
                        sb2 += diff * Iy[q][j];
                    }
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
            status = (st && (double)dist <= 0.5) ? 1 : 0;
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

int launch_lk(const LkGroup* d_groups, int n_groups, int max_pts, int max_level, int flow_back, cudaStream_t st) {
    if (n_groups <= 0 || max_pts <= 0) return DVFE_OK;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, n_groups);
    DVFE_LAUNCH(k_lk_track, grid, LK_WARPS * 32, 0, st, d_groups, max_level, flow_back);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
