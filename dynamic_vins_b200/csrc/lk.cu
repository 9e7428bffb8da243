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
// Mapping: one warp per point, all pyramid levels, forward then backward, in one launch.  The 21x21
// template (I, Ix, Iy) lives in registers (14 pixels per lane); the Scharr derivatives are computed on
// the fly from a 24x24 u8 window staged in shared memory (the reference materialises a 4 B/px derivative
// image per level per call); the 2x2 sums are reduced with redux.sync.
#include "kernels.cuh"

#define LK_WARPS 4
#define LK_PPL 14                 // pixels per lane: 14*32 = 448 >= 441
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

// cv::calcOpticalFlowPyrLK for one point, levels max_level..0.
//   pyrI/pyrJ : pyramids of the template / search image
//   p1        : point in the template image (level-0 coordinates)
//   p2        : in: initial guess (use_init), out: tracked point
// Returns status (0/1).
__device__ int lk_track_point(const uint8_t* __restrict__ pyrI, const uint8_t* __restrict__ pyrJ,
                              const PyrDesc& desc, int max_level, float2 p1, bool use_init, float2& p2,
                              uint8_t* __restrict__ win /* [24*24] */, short2* __restrict__ der /* [22*22] */,
                              int lane, const int (&pxy)[LK_PPL]) {
    const float FLT_SCALE = 1.f / (1 << 20);
    int status = 1;
    float nextx = p2.x, nexty = p2.y;
    float outx = nextx, outy = nexty;      // nextPts[ptidx]
    for (int level = max_level; level >= 0; --level) {
        const PyrLevel L = desc.lv[level];
        const uint8_t* __restrict__ Ipx = pyrI + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX;
        const uint8_t* __restrict__ Jpx = pyrJ + L.offset + (size_t)DVFE_PADY * L.pitch + DVFE_PADX;
        const float scale = __int_as_float((127 - level) << 23);      // (float)(1./(1 << level))
        float prevx = p1.x * scale, prevy = p1.y * scale;
        if (level == max_level) {
            if (use_init) { nextx = outx * scale; nexty = outy * scale; }
            else { nextx = prevx; nexty = prevy; }
        } else {
            nextx = outx * 2.f; nexty = outy * 2.f;
        }
        outx = nextx; outy = nexty;

        prevx -= DVFE_HALF_WIN; prevy -= DVFE_HALF_WIN;
        const int ipx = __float2int_rd(prevx), ipy = __float2int_rd(prevy);
        if (ipx < -DVFE_WIN || ipx >= L.w || ipy < -DVFE_WIN || ipy >= L.h) {
            if (level == 0) status = 0;
            continue;
        }
        // ---- stage the 24x24 window of I around the patch, derive Scharr taps (zero outside the image)
        __syncwarp();
        for (int i = lane; i < 24 * 24; i += 32) {
            const int r = i / 24, c = i - r * 24;
            win[i] = __ldg(Ipx + (ipy - 1 + r) * L.pitch + (ipx - 1 + c));
        }
        __syncwarp();
        for (int i = lane; i < 22 * 22; i += 32) {
            const int r = i / 22, c = i - r * 22;
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
            der[i] = d;
        }
        __syncwarp();

        float a = prevx - (float)ipx, b = prevy - (float)ipy;
        int iw00, iw01, iw10, iw11;
        lk_weights(a, b, iw00, iw01, iw10, iw11);

        int Iw[LK_PPL], Ix[LK_PPL], Iy[LK_PPL];
        int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
        for (int k = 0; k < LK_PPL; k++) {
            const int x = pxy[k] & 255, y = pxy[k] >> 8;
            if (pxy[k] >= 0) {
                const uint8_t* w0 = win + (y + 1) * 24 + x + 1;
                const int ival = (w0[0] * iw00 + w0[1] * iw01 + w0[24] * iw10 + w0[25] * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
                const short2 d00 = der[y * 22 + x], d01 = der[y * 22 + x + 1];
                const short2 d10 = der[(y + 1) * 22 + x], d11 = der[(y + 1) * 22 + x + 1];
                const int ixv = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                const int iyv = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                Iw[k] = ival; Ix[k] = ixv; Iy[k] = iyv;
                sA11 += ixv * ixv; sA12 += ixv * iyv; sA22 += iyv * iyv;
            } else {
                Iw[k] = 0; Ix[k] = 0; Iy[k] = 0;
            }
        }
        const float A11 = __ll2float_rn(warp_sum_i64(sA11)) * FLT_SCALE;
        const float A12 = __ll2float_rn(warp_sum_i64(sA12)) * FLT_SCALE;
        const float A22 = __ll2float_rn(warp_sum_i64(sA22)) * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * DVFE_WIN * DVFE_WIN);
        if ((double)minEig < 1e-4 || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        nextx -= DVFE_HALF_WIN; nexty -= DVFE_HALF_WIN;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < 30; j++) {
            const int inx = __float2int_rd(nextx), iny = __float2int_rd(nexty);
            if (inx < -DVFE_WIN || inx >= L.w || iny < -DVFE_WIN || iny >= L.h) {
                if (level == 0) status = 0;
                break;
            }
            a = nextx - (float)inx; b = nexty - (float)iny;
            lk_weights(a, b, iw00, iw01, iw10, iw11);
            const uint8_t* __restrict__ Jw = Jpx + iny * L.pitch + inx;
            int sb1 = 0, sb2 = 0;
#pragma unroll
            for (int k = 0; k < LK_PPL; k++) {
                const int x = pxy[k] & 255, y = (pxy[k] >> 8) & 255;     // invalid slots read pixel (0,0): Ix=Iy=0
                const uint8_t* q = Jw + y * L.pitch + x;
                const int v = (int)__ldg(q) * iw00 + (int)__ldg(q + 1) * iw01 + (int)__ldg(q + L.pitch) * iw10 +
                              (int)__ldg(q + L.pitch + 1) * iw11;
                const int diff = ((v + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - Iw[k];
                sb1 += diff * Ix[k];
                sb2 += diff * Iy[k];
            }
            const float b1 = __ll2float_rn(warp_sum_i64(sb1)) * FLT_SCALE;
            const float b2 = __ll2float_rn(warp_sum_i64(sb2)) * FLT_SCALE;
            const float dx = (A12 * b2 - A22 * b1) * D;
            const float dy = (A12 * b1 - A11 * b2) * D;
            nextx += dx; nexty += dy;
            outx = nextx + DVFE_HALF_WIN; outy = nexty + DVFE_HALF_WIN;
            if ((double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                outx -= dx * 0.5f; outy -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            const int qx = __float2int_rd(outx - DVFE_HALF_WIN), qy = __float2int_rd(outy - DVFE_HALF_WIN);
            if (qx < -DVFE_WIN || qx >= L.w || qy < -DVFE_WIN || qy >= L.h) status = 0;
        }
    }
    p2.x = outx; p2.y = outy;
    return status;
}

__global__ void __launch_bounds__(LK_WARPS * 32) k_lk_track(const LkGroup* __restrict__ groups, int max_level, int flow_back) {
    __shared__ __align__(16) uint8_t s_win[LK_WARPS][24 * 24];
    __shared__ __align__(16) short2 s_der[LK_WARPS][22 * 22];
    const LkGroup& G = groups[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * LK_WARPS + warp;
    const int n = *G.n;
    if (i >= n) return;

    int pxy[LK_PPL];
#pragma unroll
    for (int k = 0; k < LK_PPL; k++) {
        const int p = k * 32 + lane;
        const int y = p / DVFE_WIN, x = p - y * DVFE_WIN;
        pxy[k] = (p < DVFE_WIN * DVFE_WIN) ? (x | (y << 8)) : (int)0x80000000;
    }

    float2 p1 = G.ptsA[i];
    p1.x += G.offx; p1.y += G.offy;
    const int top = G.desc.n_levels - 1;
    const int lf = max_level < top ? max_level : top;
    const int lb = 1 < top ? 1 : top;

    float2 p2 = make_float2(0.f, 0.f);
    int status = lk_track_point(G.pyrA, G.pyrB, G.desc, lf, p1, false, p2, s_win[warp], s_der[warp], lane, pxy);
    float2 rev = p1;
    if (flow_back && status) {
        const int sb = lk_track_point(G.pyrB, G.pyrA, G.desc, lb, p2, true, rev, s_win[warp], s_der[warp], lane, pxy);
        const float ddx = p1.x - rev.x, ddy = p1.y - rev.y;
        const float dist = sqrtf(ddx * ddx + ddy * ddy);
        status = (sb && (double)dist <= 0.5) ? 1 : 0;
    }
    if (status) {
        const int W = G.desc.lv[0].w, H = G.desc.lv[0].h;
        const int rx = __float2int_rn(p2.x), ry = __float2int_rn(p2.y);
        if (!(1 <= rx && rx < W - 1 && 1 <= ry && ry < H - 1)) status = 0;
        else if (G.mask != nullptr && G.mask[(size_t)ry * G.mask_pitch + rx] == 0) status = 0;
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
