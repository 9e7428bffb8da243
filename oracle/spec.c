/* ORACLE (test infrastructure, NOT product code).
 *
 * PARITY STATUS: the stages below are OpenCV's (a third-party dependency of the reference, absent from /root/reference); they are
 * pinned against cv2 4.13.0 itself at the reference's call sites (tests/test_oracle_pinned.py, tests/test_epipolar.py) and, where
 * the reference's own code is involved (liftProjective, FeatureTrackByLK's glue, ErodeMask, InstanceImagePadding,
 * DetectExtraPoints, RejectWithF), against the reference's sources compiled unmodified into oracle/_ref/libdvref.so
 * (tests/test_ref_compiled.py, tests/test_epipolar.py).  What stays unpinned is the OpenCV version: the reference pins 3.4.16,
 * the image has 4.13.0.
 *
 * Plain-C arithmetic restatement of the OpenCV stages the reference's CPU front-end
 * calls.  The arithmetic lives in a third-party dependency that is NOT under
 * /root/reference: OpenCV, pinned 3.4.16 by dynamic_vins/CMakeLists.txt:36.  The
 * published algorithms (modules/video/src/lkpyramid.cpp, modules/imgproc/src/
 * {pyramids,corner,featureselect,morph,drawing}.cpp) are restated here from
 * SURVEY.md Appendix A/B and are pinned in tests/test_oracle_spec.py against the
 * `cv2` 4.13.0 build in this image, at the reference's own call sites:
 *   - cv::calcOpticalFlowPyrLK(img1,img2,pts1,pts2,status,err,Size(21,21),3)
 *       dynamic_vins/src/front_end/feature_utils.cpp:43
 *   - cv::calcOpticalFlowPyrLK(..., Size(21,21), 1, TermCriteria(COUNT+EPS,30,0.01), OPTFLOW_USE_INITIAL_FLOW)
 *       dynamic_vins/src/front_end/feature_utils.cpp:50-53
 *   - cv::goodFeaturesToTrack(gray, pts, n, 0.01, min_dist, mask)
 *       dynamic_vins/src/front_end/background_tracker.cpp:85, instance_feature.cpp:381, dynamic_tracker.cpp:435
 *   - cv::circle(mask, pt, r, 0, -1)      background_tracker.cpp:80, instance_feature.cpp:368, dynamic_tracker.cpp:430
 *   - cv::erode(rect k x k)               feature_utils.h:142-146
 *   - camodocal PinholeCamera::liftProjective  camera_models/src/camera_models/PinholeCamera.cc:450-510
 *
 * Two accumulation modes exist for the LK normal equations:
 *   exact_int = 0 : float accumulators (what OpenCV's scalar loop does)
 *   exact_int = 1 : exact 64-bit integer sums converted to float once (what the
 *                   CUDA kernels do; the value OpenCV's float sums approximate)
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define WIN 21
#define HALF 10.0f
#define BORDER 24          /* >= 22: patch taps reach [-21, w+20] */

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* ---- cv::pyrDown, 8-bit: separable [1 4 6 4 1], REFLECT_101, (s+128)>>8, size ((w+1)/2,(h+1)/2) ---- */
void spec_pyr_down(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* row = (int*)malloc(sizeof(int) * (size_t)w * 5);
    for (int y = 0; y < dh; y++) {
        int acc_rows[5];
        for (int k = 0; k < 5; k++) acc_rows[k] = reflect101(2 * y + k - 2, h);
        for (int x = 0; x < dw; x++) {
            int s = 0;
            static const int kw[5] = {1, 4, 6, 4, 1};
            for (int j = 0; j < 5; j++) {
                const uint8_t* r = src + (size_t)acc_rows[j] * sstride;
                int hs = 0;
                for (int i = 0; i < 5; i++) hs += kw[i] * r[reflect101(2 * x + i - 2, w)];
                s += kw[j] * hs;
            }
            dst[(size_t)y * dstride + x] = (uint8_t)((s + 128) >> 8);
        }
    }
    free(row);
}

/* number of the last pyramid level cv::buildOpticalFlowPyramid keeps (winSize 21x21) */
int spec_pyr_levels(int w, int h, int max_level) {
    int lvl = 0;
    while (lvl < max_level) {
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= WIN || nh <= WIN) break;
        w = nw; h = nh; lvl++;
    }
    return lvl;
}

/* ---- calcSharrDeriv: int16 (dx,dy) interleaved, REFLECT_101 inside the image ---- */
void spec_scharr(const uint8_t* src, int w, int h, int sstride, int16_t* dst /* h*w*2 */) {
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * sstride;
        const uint8_t* r1 = src + (size_t)y * sstride;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * sstride;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
            int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
            int t1m = r2[xm] - r0[xm];
            int t1c = r2[x] - r0[x];
            int t1p = r2[xp] - r0[xp];
            dst[((size_t)y * w + x) * 2 + 0] = (int16_t)(t0p - t0m);
            dst[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
        }
    }
}

typedef struct {
    int w, h;
    int pitch;            /* of the padded u8 image */
    uint8_t* img;         /* padded, origin at img + BORDER*pitch + BORDER */
    int16_t* der;         /* padded (dx,dy), zero border, pitch*2 int16 per row */
} level_t;

static void make_level(level_t* L, const uint8_t* src, int w, int h, int sstride, int with_deriv) {
    L->w = w; L->h = h; L->pitch = w + 2 * BORDER;
    size_t rows = (size_t)h + 2 * BORDER;
    L->img = (uint8_t*)malloc(rows * L->pitch);
    for (int y = -BORDER; y < h + BORDER; y++) {
        const uint8_t* r = src + (size_t)reflect101(y, h) * sstride;
        uint8_t* d = L->img + (size_t)(y + BORDER) * L->pitch + BORDER;
        for (int x = -BORDER; x < w + BORDER; x++) d[x] = r[reflect101(x, w)];
    }
    L->der = NULL;
    if (with_deriv) {
        L->der = (int16_t*)calloc(rows * L->pitch * 2, sizeof(int16_t));
        int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * h * 2);
        spec_scharr(src, w, h, sstride, tmp);
        for (int y = 0; y < h; y++)
            memcpy(L->der + ((size_t)(y + BORDER) * L->pitch + BORDER) * 2, tmp + (size_t)y * w * 2,
                   sizeof(int16_t) * (size_t)w * 2);
        free(tmp);
    }
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }   /* round-half-even (default FE mode) */
static inline int cv_floor_f(float v) { return (int)floorf(v); }
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

/* One LK level for one point: LKTrackerInvoker::operator() body (lkpyramid.cpp). */
static void lk_point_level(const level_t* I, const level_t* J, int level, int max_level, int use_init,
                           const float* prev_pt_in, float* next_pt_io, uint8_t* status, int exact_int) {
    const float FLT_SCALE = 1.f / (1 << 20);
    const int W_BITS = 14;
    float scale = (float)(1. / (1 << level));
    float prevx = prev_pt_in[0] * scale, prevy = prev_pt_in[1] * scale;
    float nextx, nexty;
    if (level == max_level) {
        if (use_init) { nextx = next_pt_io[0] * scale; nexty = next_pt_io[1] * scale; }
        else { nextx = prevx; nexty = prevy; }
    } else { nextx = next_pt_io[0] * 2.f; nexty = next_pt_io[1] * 2.f; }
    next_pt_io[0] = nextx; next_pt_io[1] = nexty;

    prevx -= HALF; prevy -= HALF;
    int ipx = cv_floor_f(prevx), ipy = cv_floor_f(prevy);
    if (ipx < -WIN || ipx >= I->w || ipy < -WIN || ipy >= I->h) {
        if (level == 0) *status = 0;
        return;
    }
    float a = prevx - ipx, b = prevy - ipy;
    int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
    int iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
    int iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

    int16_t Iw[WIN * WIN], dIx[WIN * WIN], dIy[WIN * WIN];
    float fA11 = 0, fA12 = 0, fA22 = 0;
    int64_t iA11 = 0, iA12 = 0, iA22 = 0;
    int stepI = I->pitch, dstep = I->pitch * 2;
    for (int y = 0; y < WIN; y++) {
        const uint8_t* src = I->img + (size_t)(y + ipy + BORDER) * stepI + ipx + BORDER;
        const int16_t* dsrc = I->der + ((size_t)(y + ipy + BORDER) * I->pitch + ipx + BORDER) * 2;
        for (int x = 0; x < WIN; x++, dsrc += 2) {
            int ival = DESCALE(src[x] * iw00 + src[x + 1] * iw01 + src[x + stepI] * iw10 + src[x + stepI + 1] * iw11,
                               W_BITS - 5);
            int ixval = DESCALE(dsrc[0] * iw00 + dsrc[2] * iw01 + dsrc[dstep] * iw10 + dsrc[dstep + 2] * iw11, W_BITS);
            int iyval = DESCALE(dsrc[1] * iw00 + dsrc[3] * iw01 + dsrc[dstep + 1] * iw10 + dsrc[dstep + 3] * iw11,
                                W_BITS);
            Iw[y * WIN + x] = (int16_t)ival;
            dIx[y * WIN + x] = (int16_t)ixval;
            dIy[y * WIN + x] = (int16_t)iyval;
            fA11 += (float)(ixval * ixval); fA12 += (float)(ixval * iyval); fA22 += (float)(iyval * iyval);
            iA11 += (int64_t)ixval * ixval; iA12 += (int64_t)ixval * iyval; iA22 += (int64_t)iyval * iyval;
        }
    }
    float A11, A12, A22;
    if (exact_int) { A11 = (float)iA11 * FLT_SCALE; A12 = (float)iA12 * FLT_SCALE; A22 = (float)iA22 * FLT_SCALE; }
    else { A11 = fA11 * FLT_SCALE; A12 = fA12 * FLT_SCALE; A22 = fA22 * FLT_SCALE; }
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
    if ((double)minEig < 1e-4 || D < FLT_EPSILON) {
        if (level == 0) *status = 0;
        return;
    }
    D = 1.f / D;
    nextx -= HALF; nexty -= HALF;
    float pdx = 0, pdy = 0;
    for (int j = 0; j < 30; j++) {
        int inx = cv_floor_f(nextx), iny = cv_floor_f(nexty);
        if (inx < -WIN || inx >= J->w || iny < -WIN || iny >= J->h) {
            if (level == 0) *status = 0;
            break;
        }
        a = nextx - inx; b = nexty - iny;
        iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
        iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
        iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
        iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        float fb1 = 0, fb2 = 0;
        int64_t ib1 = 0, ib2 = 0;
        int stepJ = J->pitch;
        for (int y = 0; y < WIN; y++) {
            const uint8_t* Jp = J->img + (size_t)(y + iny + BORDER) * stepJ + inx + BORDER;
            for (int x = 0; x < WIN; x++) {
                int diff = DESCALE(Jp[x] * iw00 + Jp[x + 1] * iw01 + Jp[x + stepJ] * iw10 + Jp[x + stepJ + 1] * iw11,
                                   W_BITS - 5) - Iw[y * WIN + x];
                fb1 += (float)(diff * dIx[y * WIN + x]);
                fb2 += (float)(diff * dIy[y * WIN + x]);
                ib1 += (int64_t)diff * dIx[y * WIN + x];
                ib2 += (int64_t)diff * dIy[y * WIN + x];
            }
        }
        float b1, b2;
        if (exact_int) { b1 = (float)ib1 * FLT_SCALE; b2 = (float)ib2 * FLT_SCALE; }
        else { b1 = fb1 * FLT_SCALE; b2 = fb2 * FLT_SCALE; }
        float dx = (A12 * b2 - A22 * b1) * D;
        float dy = (A12 * b1 - A11 * b2) * D;
        nextx += dx; nexty += dy;
        next_pt_io[0] = nextx + HALF; next_pt_io[1] = nexty + HALF;
        if ((double)dx * dx + (double)dy * dy <= 0.01 * 0.01) break;
        if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
            next_pt_io[0] -= dx * 0.5f; next_pt_io[1] -= dy * 0.5f;
            break;
        }
        pdx = dx; pdy = dy;
    }
    if (*status && level == 0) {
        float qx = next_pt_io[0] - HALF, qy = next_pt_io[1] - HALF;
        int ix = cv_floor_f(qx), iy = cv_floor_f(qy);
        if (ix < -WIN || ix >= J->w || iy < -WIN || iy >= J->h) *status = 0;
    }
}

/* cv::calcOpticalFlowPyrLK(img1,img2,pts1,pts2,status,err,Size(21,21),max_level,(COUNT+EPS,30,0.01),flags)
 * pts2 is in/out (initial flow when use_init).  Returns the effective max level. */
int spec_lk(const uint8_t* img1, const uint8_t* img2, int w, int h, int stride1, int stride2,
            const float* pts1, float* pts2, uint8_t* status, int n, int max_level, int use_init, int exact_int) {
    int L = spec_pyr_levels(w, h, max_level);
    level_t* P1 = (level_t*)malloc(sizeof(level_t) * (L + 1));
    level_t* P2 = (level_t*)malloc(sizeof(level_t) * (L + 1));
    uint8_t *c1 = NULL, *c2 = NULL;
    const uint8_t *s1 = img1, *s2 = img2;
    int cw = w, ch = h, st1 = stride1, st2 = stride2;
    for (int l = 0; l <= L; l++) {
        make_level(&P1[l], s1, cw, ch, st1, 1);
        make_level(&P2[l], s2, cw, ch, st2, 0);
        if (l < L) {
            int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
            uint8_t* d1 = (uint8_t*)malloc((size_t)nw * nh);
            uint8_t* d2 = (uint8_t*)malloc((size_t)nw * nh);
            spec_pyr_down(s1, cw, ch, st1, d1, nw);
            spec_pyr_down(s2, cw, ch, st2, d2, nw);
            free(c1); free(c2);
            c1 = d1; c2 = d2; s1 = d1; s2 = d2; cw = nw; ch = nh; st1 = st2 = nw;
        }
    }
    free(c1); free(c2);
    for (int i = 0; i < n; i++) status[i] = 1;
    for (int l = L; l >= 0; l--)
        for (int i = 0; i < n; i++)
            lk_point_level(&P1[l], &P2[l], l, L, use_init, pts1 + 2 * i, pts2 + 2 * i, status + i, exact_int);
    for (int l = 0; l <= L; l++) { free(P1[l].img); free(P1[l].der); free(P2[l].img); }
    free(P1); free(P2);
    return L;
}

/* FeatureTrackByLK  (dynamic_vins/src/front_end/feature_utils.cpp:35-69) */
void spec_feature_track_by_lk(const uint8_t* img1, const uint8_t* img2, int w, int h, int stride1, int stride2,
                              const float* pts1, float* pts2, uint8_t* status, int n, int flow_back,
                              int max_level, int exact_int, float* rev_out /* nullable, n*2 */) {
    spec_lk(img1, img2, w, h, stride1, stride2, pts1, pts2, status, n, max_level, 0, exact_int);
    if (flow_back) {
        float* rev = (float*)malloc(sizeof(float) * 2 * (size_t)n);
        uint8_t* rst = (uint8_t*)malloc((size_t)n);
        memcpy(rev, pts1, sizeof(float) * 2 * (size_t)n);
        spec_lk(img2, img1, w, h, stride2, stride1, pts2, rev, rst, n, 1, 1, exact_int);
        for (int i = 0; i < n; i++) {
            float dx = pts1[2 * i] - rev[2 * i], dy = pts1[2 * i + 1] - rev[2 * i + 1];
            float d = sqrtf(dx * dx + dy * dy);
            status[i] = (status[i] && rst[i] && (double)d <= 0.5) ? 1 : 0;
        }
        if (rev_out) memcpy(rev_out, rev, sizeof(float) * 2 * (size_t)n);
        free(rev); free(rst);
    }
    for (int i = 0; i < n; i++) {
        if (!status[i]) continue;
        int x = cv_round_f(pts2[2 * i]), y = cv_round_f(pts2[2 * i + 1]);
        if (!(1 <= x && x < w - 1 && 1 <= y && y < h - 1)) status[i] = 0;      /* InBorder, feature_utils.h:68-74 */
    }
}

/* ---- cv::cornerMinEigenVal(blockSize 3, ksize 3) ----  */
void spec_min_eigen_val(const uint8_t* img, int w, int h, int stride, float* out /* h*w */) {
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));   /* scale = 1/((1<<(ksize-1))*blockSize*255) */
    const float s2 = s * 2.0f;
    size_t n = (size_t)w * h;
    float* Dx = (float*)malloc(sizeof(float) * n);
    float* Dy = (float*)malloc(sizeof(float) * n);
    /* Dx: row kernel [-1 0 1] (exact ints), column kernel [1 2 1]*scale : fma(s, d0+d2, s2*d1)
     * Dy: row kernel [1 2 1]*scale: fma(s, r, fma(s2, c, s*l)), column kernel [-1 0 1] */
    float* dxr = (float*)malloc(sizeof(float) * n);
    float* smr = (float*)malloc(sizeof(float) * n);
    for (int y = 0; y < h; y++) {
        const uint8_t* r = img + (size_t)y * stride;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            dxr[(size_t)y * w + x] = (float)((int)r[xp] - (int)r[xm]);
            smr[(size_t)y * w + x] = fmaf(s, (float)r[xp], fmaf(s2, (float)r[x], s * (float)r[xm]));
        }
    }
    for (int y = 0; y < h; y++) {
        int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
        for (int x = 0; x < w; x++) {
            float d0 = dxr[(size_t)ym * w + x], d1 = dxr[(size_t)y * w + x], d2 = dxr[(size_t)yp * w + x];
            Dx[(size_t)y * w + x] = fmaf(s, d0 + d2, s2 * d1);
            Dy[(size_t)y * w + x] = smr[(size_t)yp * w + x] - smr[(size_t)ym * w + x];
        }
    }
    free(dxr); free(smr);
    /* cv::boxFilter(cov, cov, CV_32F, Size(3,3), normalize=false, BORDER_REFLECT_101): RowSum<float,double>
     * (horizontal 3-sum, left to right) then ColumnSum<double,float> (vertical 3-sum, top to bottom) */
    double* H = (double*)malloc(sizeof(double) * n * 3);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            double sxx = 0, sxy = 0, syy = 0;
            for (int i = -1; i <= 1; i++) {
                int xx = reflect101(x + i, w);
                float dx = Dx[(size_t)y * w + xx], dy = Dy[(size_t)y * w + xx];
                sxx += (double)(dx * dx); sxy += (double)(dx * dy); syy += (double)(dy * dy);
            }
            H[((size_t)y * w + x) * 3 + 0] = sxx; H[((size_t)y * w + x) * 3 + 1] = sxy; H[((size_t)y * w + x) * 3 + 2] = syy;
        }
    for (int y = 0; y < h; y++) {
        int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
        for (int x = 0; x < w; x++) {
            const double* h0 = H + ((size_t)ym * w + x) * 3;
            const double* h1 = H + ((size_t)y * w + x) * 3;
            const double* h2 = H + ((size_t)yp * w + x) * 3;
            float a = (float)((h0[0] + h1[0]) + h2[0]) * 0.5f, b = (float)((h0[1] + h1[1]) + h2[1]);
            float c = (float)((h0[2] + h1[2]) + h2[2]) * 0.5f;
            out[(size_t)y * w + x] = (a + c) - sqrtf((a - c) * (a - c) + b * b);
        }
    }
    free(H);
    free(Dx); free(Dy);
}

typedef struct { float v; int idx; } cand_t;
static int cand_cmp(const void* pa, const void* pb) {
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->idx > b->idx) ? -1 : (a->idx < b->idx ? 1 : 0);
}

/* cv::goodFeaturesToTrack stages 2..6 on a given response map `eig` (h*w floats).
 * Returns number of corners written to out_xy (x,y floats, acceptance order). */
/* unmasked_max != 0: the quality threshold comes from the maximum over the WHOLE response map (what
 * cv::cuda::GoodFeaturesToTrackDetector::detect does: cuda::minMax(eig_, 0, &maxVal) without the mask,
 * opencv/modules/cudaimgproc/src/gftt.cpp; candidates, ordering and the distance grid are the CPU detector's). */
int spec_gftt_select_ex(const float* eig, int w, int h, const uint8_t* mask, int mstride, int max_corners,
                        double quality, double min_dist, float* out_xy, int* n_cand_out, int unmasked_max) {
    float maxv = 0; int any = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (unmasked_max || !mask || mask[(size_t)y * mstride + x]) {
                float v = eig[(size_t)y * w + x];
                if (!any || v > maxv) { maxv = v; any = 1; }
            }
    double maxVal = any ? (double)maxv : 0.0;
    float thr = (float)(maxVal * quality);
    size_t cap = 1024, nc = 0;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * cap);
#define TZ(v) ((v) > thr ? (v) : 0.f)
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            float v = TZ(eig[(size_t)y * w + x]);
            if (v == 0) continue;
            if (mask && !mask[(size_t)y * mstride + x]) continue;
            float m = v;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    float t = TZ(eig[(size_t)(y + j) * w + x + i]);
                    if (t > m) m = t;
                }
            if (v != m) continue;
            if (nc == cap) { cap *= 2; c = (cand_t*)realloc(c, sizeof(cand_t) * cap); }
            c[nc].v = v; c[nc].idx = y * w + x; nc++;
        }
    if (n_cand_out) *n_cand_out = (int)nc;
    qsort(c, nc, sizeof(cand_t), cand_cmp);
    int ncorners = 0;
    if (min_dist >= 1) {
        int cell = (int)lrint(min_dist);
        int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        int* head = (int*)malloc(sizeof(int) * (size_t)gw * gh);
        int* next = (int*)malloc(sizeof(int) * (nc + 1));
        float* ax = (float*)malloc(sizeof(float) * (nc + 1));
        float* ay = (float*)malloc(sizeof(float) * (nc + 1));
        for (int i = 0; i < gw * gh; i++) head[i] = -1;
        double md2 = min_dist * min_dist;
        for (size_t i = 0; i < nc; i++) {
            int y = c[i].idx / w, x = c[i].idx - y * w;
            int xc = x / cell, yc = y / cell;
            int x1 = xc - 1 < 0 ? 0 : xc - 1, y1 = yc - 1 < 0 ? 0 : yc - 1;
            int x2 = xc + 1 > gw - 1 ? gw - 1 : xc + 1, y2 = yc + 1 > gh - 1 ? gh - 1 : yc + 1;
            int good = 1;
            for (int yy = y1; yy <= y2 && good; yy++)
                for (int xx = x1; xx <= x2 && good; xx++)
                    for (int k = head[yy * gw + xx]; k >= 0; k = next[k]) {
                        float dx = x - ax[k], dy = y - ay[k];
                        if ((double)(dx * dx + dy * dy) < md2) { good = 0; break; }
                    }
            if (good) {
                ax[ncorners] = (float)x; ay[ncorners] = (float)y;
                next[ncorners] = head[yc * gw + xc]; head[yc * gw + xc] = ncorners;
                out_xy[2 * ncorners] = (float)x; out_xy[2 * ncorners + 1] = (float)y;
                ncorners++;
                if (max_corners > 0 && ncorners == max_corners) break;
            }
        }
        free(head); free(next); free(ax); free(ay);
    } else {
        for (size_t i = 0; i < nc; i++) {
            int y = c[i].idx / w, x = c[i].idx - y * w;
            out_xy[2 * ncorners] = (float)x; out_xy[2 * ncorners + 1] = (float)y;
            ncorners++;
            if (max_corners > 0 && ncorners == max_corners) break;
        }
    }
    free(c);
    return ncorners;
}

int spec_gftt_select(const float* eig, int w, int h, const uint8_t* mask, int mstride, int max_corners,
                     double quality, double min_dist, float* out_xy, int* n_cand_out) {
    return spec_gftt_select_ex(eig, w, h, mask, mstride, max_corners, quality, min_dist, out_xy, n_cand_out, 0);
}

int spec_good_features(const uint8_t* img, int w, int h, int stride, const uint8_t* mask, int mstride,
                       int max_corners, double quality, double min_dist, float* out_xy) {
    float* eig = (float*)malloc(sizeof(float) * (size_t)w * h);
    spec_min_eigen_val(img, w, h, stride, eig);
    int n = spec_gftt_select(eig, w, h, mask, mstride, max_corners, quality, min_dist, out_xy, NULL);
    free(eig);
    return n;
}

/* cv::circle(mask, cvRound(pt), r, 0, -1): pixel cleared  <=>  dx^2+dy^2 <= r^2 (clipped) */
void spec_disc_mask(uint8_t* mask, int w, int h, int stride, const float* pts, int n, int r) {
    for (int i = 0; i < n; i++) {
        int cx = cv_round_f(pts[2 * i]), cy = cv_round_f(pts[2 * i + 1]);
        for (int dy = -r; dy <= r; dy++) {
            int y = cy + dy;
            if (y < 0 || y >= h) continue;
            for (int dx = -r; dx <= r; dx++) {
                int x = cx + dx;
                if (x < 0 || x >= w) continue;
                if (dx * dx + dy * dy <= r * r) mask[(size_t)y * stride + x] = 0;
            }
        }
    }
}

/* cv::erode with MORPH_RECT k x k, anchor k/2, border = +inf */
void spec_erode_rect(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int k) {
    int a = k / 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int m = 255;
            for (int j = 0; j < k; j++) {
                int yy = y - a + j;
                if (yy < 0 || yy >= h) continue;
                for (int i = 0; i < k; i++) {
                    int xx = x - a + i;
                    if (xx < 0 || xx >= w) continue;
                    int v = src[(size_t)yy * sstride + xx];
                    if (v < m) m = v;
                }
            }
            dst[(size_t)y * dstride + x] = (uint8_t)m;
        }
}

/* camodocal PinholeCamera::liftProjective + the b.x/b.z float narrowing of InstFeat::UndistortedPts */
void spec_lift(const double* cam /* fx fy cx cy k1 k2 p1 p2 */, const float* pts, int n, float off_x, float off_y,
               float* out) {
    double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], k1 = cam[4], k2 = cam[5], p1 = cam[6], p2 = cam[7];
    double iK11 = 1.0 / fx, iK13 = -cx / fx, iK22 = 1.0 / fy, iK23 = -cy / fy;
    int nod = (k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0);
    for (int i = 0; i < n; i++) {
        double u = (double)(pts[2 * i] + off_x), v = (double)(pts[2 * i + 1] + off_y);
        double mxd = iK11 * u + iK13, myd = iK22 * v + iK23, mxu = mxd, myu = myd;
        if (!nod) {
            for (int it = 0; it < 8; it++) {
                double x = (it == 0) ? mxd : mxu, y = (it == 0) ? myd : myu;
                double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
                double rad = k1 * rho2 + k2 * rho2 * rho2;
                double dux = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2);
                double duy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2);
                mxu = mxd - dux; myu = myd - duy;
            }
        }
        out[2 * i] = (float)(mxu / 1.0); out[2 * i + 1] = (float)(myu / 1.0);
    }
}

/* cv::remap(src, dst, map1 CV_16SC2, map2 CV_16UC1, INTER_LINEAR, BORDER_CONSTANT 0) as ImageProcessor::Run calls it
 * (image_process/image_process.cpp:109-122) [OpenCV imgwarp.cpp remapBilinear, fixed-point path]: weights from the
 * 32 x 32 bilinear table scaled by 2^15, taps outside the source read 0, rounding (v + 2^14) >> 15.  ch interleaved
 * channels; dst has the size of the maps = the size of src. */
void spec_remap(const uint8_t* src, int w, int h, int ch, int sstride, const int16_t* map1, const uint16_t* map2,
                uint8_t* dst, int dstride) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t m = (size_t)y * w + x;
            const int sx = map1[2 * m], sy = map1[2 * m + 1];
            const int fx = map2[m] & 31, fy = (map2[m] >> 5) & 31;
            const int wt[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
            for (int c = 0; c < ch; c++) {
                int acc = 1 << 14;
                for (int k = 0; k < 4; k++) {
                    const int xx = sx + (k & 1), yy = sy + (k >> 1);
                    if (xx >= 0 && xx < w && yy >= 0 && yy < h) acc += wt[k] * src[(size_t)yy * sstride + xx * ch + c];
                }
                dst[(size_t)y * dstride + x * ch + c] = (uint8_t)(acc >> 15);
            }
        }
}

/* cv::cvtColor(BGR2GRAY) with the 15-bit coefficients of cv2 4.13 (SemanticImage::SetGrayImage, basic/semantic_image.cpp:69-73) */
void spec_bgr_to_gray(const uint8_t* bgr, int w, int h, int sstride, uint8_t* dst, int dstride) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = bgr + (size_t)y * sstride + 3 * x;
            dst[(size_t)y * dstride + x] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
        }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * cv::findFundamentalMat(points1, points2, cv::FM_RANSAC, param1, param2, mask) as InstsFeatManager::RejectWithF calls it
 * (dynamic_vins/src/front_end/dynamic_tracker.cpp:846; FeatureTracker::RejectWithF, background_tracker.cpp:520-550, is the
 * same call, commented out).  The algorithm lives in OpenCV (modules/calib3d/src/{fundam,ptsetreg}.cpp), absent from
 * /root/reference; restated from its published source and pinned against cv2 4.13.0 in tests/test_oracle_pinned.py:
 *   n < 7        no model, the mask stays empty (returns 0)
 *   n == 7       the 7-point solver runs once, mask = all ones
 *   8 <= n < 15  LMedS  (createLMeDSPointSetRegistrator(cb, 7, confidence)): 7-point samples, the model with the smallest
 *                median (element count/2 after nth_element) error wins, inliers within sigma = 2.5*1.4826*(1+5/(n-7))*sqrt(med)
 *   n >= 15      RANSAC (createRANSACPointSetRegistrator(cb, 7, threshold, confidence), maxIters 1000)
 * Samples: cv::RNG(uint64(-1)) (multiply-with-carry, 4164903690), `uniform(0, n)` = next() % n, duplicates redrawn, a sample is
 * rejected when its last point is collinear with two earlier ones in either image (haveCollinearPoints).
 * 7-point solver (run7Point): F = lambda*f1 + (1-lambda)*f2 over the null space {f1, f2} of the 7 x 9 epipolar system, lambda
 * from det F = 0 (cv::solveCubic), each F scaled to F[8] = 1.
 * What is NOT reproducible about OpenCV here: it takes {f1, f2} from LAPACK's SVD, i.e. an arbitrary orthonormal basis of the
 * null space; any other basis gives the same F set (to rounding) but numbers the up-to-3 roots differently, so (a) F differs
 * from cv2's in the last digits and (b) a tie in inlier count between two roots of ONE sample may be won by the other root.
 * This restatement takes the null space from Gauss-Jordan elimination with complete pivoting.  For 8 <= n <= 13 the LMedS
 * median is one of the 7 sample points' own residuals (rounding noise), so OpenCV's winner is itself arbitrary there.
 * The mask it returns equals cv2's in every seeded trial of the pinned test for n == 14 and n >= 15. */
typedef struct { uint64_t state; } fm_rng;
static unsigned fm_rng_next(fm_rng* r) {
    r->state = (uint64_t)(unsigned)r->state * 4164903690u + (unsigned)(r->state >> 32);
    return (unsigned)r->state;
}
static int fm_rng_uniform(fm_rng* r, int a, int b) { return a == b ? a : (int)(fm_rng_next(r) % (unsigned)(b - a) + a); }

static int fm_solve_cubic(const double* c, double* x) {       /* c[0] x^3 + c[1] x^2 + c[2] x + c[3] = 0, cv::solveCubic */
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) return a3 == 0 ? -1 : 0;
            x[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d >= 0) {
            d = sqrt(d);
            double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
            if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
            else { x[0] = q2 / a1; x[1] = a3 / q2; }
            return d > 0 ? 2 : 1;
        }
        return 0;
    }
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
        double theta = acos(R / sqrt(Qcubed));
        double sqrtQ = sqrt(Q);
        double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2. * 3.1415926535897932384626433832795 / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * 3.1415926535897932384626433832795 / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
        else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
        return x[0] == x[1] ? 1 : 2;
    }
    d = sqrt(-d);
    double e = pow(d + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

/* null space of the 7 x 9 system by Gauss-Jordan elimination with complete pivoting: f1, f2 = the solutions with the two
 * free unknowns set to (1, 0) and (0, 1).  Returns 0 when the rank is below 7. */
static int fm_null_space(double A[7][9], double* f1, double* f2) {
    int perm[9];
    for (int j = 0; j < 9; j++) perm[j] = j;
    for (int k = 0; k < 7; k++) {
        int pr = k, pc = k;
        double best = -1.0;
        for (int i = k; i < 7; i++)
            for (int j = k; j < 9; j++)
                if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
        if (!(best > 0.0)) return 0;
        if (pr != k) for (int j = 0; j < 9; j++) { double t = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = t; }
        if (pc != k) {
            for (int i = 0; i < 7; i++) { double t = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = t; }
            int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t;
        }
        const double inv = 1.0 / A[k][k];
        for (int j = k; j < 9; j++) A[k][j] *= inv;
        for (int i = 0; i < 7; i++) {
            if (i == k) continue;
            const double m = A[i][k];
            if (m == 0.0) continue;
            for (int j = k; j < 9; j++) A[i][j] -= m * A[k][j];
        }
    }
    for (int i = 0; i < 7; i++) { f1[perm[i]] = -A[i][7]; f2[perm[i]] = -A[i][8]; }
    f1[perm[7]] = 1.0; f1[perm[8]] = 0.0;
    f2[perm[7]] = 0.0; f2[perm[8]] = 1.0;
    return 1;
}

static int fm_run7(const float* m1, const float* m2, double* F /* up to 3 x 9 */) {
    double A[7][9], f1[9], f2[9], c[4], r[3] = {0, 0, 0};
    for (int i = 0; i < 7; i++) {
        const double x0 = m1[2 * i], y0 = m1[2 * i + 1], x1 = m2[2 * i], y1 = m2[2 * i + 1];
        A[i][0] = x1 * x0; A[i][1] = x1 * y0; A[i][2] = x1;
        A[i][3] = y1 * x0; A[i][4] = y1 * y0; A[i][5] = y1;
        A[i][6] = x0; A[i][7] = y0; A[i][8] = 1;
    }
    if (!fm_null_space(A, f1, f2)) return 0;
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    const int n = fm_solve_cubic(c, r);
    if (n < 1 || n > 3) return n < 0 ? 0 : n;
    for (int k = 0; k < n; k++, F += 9) {
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; F[8] = 1.; }
        else F[8] = 0.;
        for (int i = 0; i < 8; i++) F[i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}

static int fm_collinear(const float* p, int count) {          /* haveCollinearPoints: only the last point is tested */
    const int i = count - 1;
    for (int j = 0; j < i; j++) {
        const double dx1 = p[2 * j] - p[2 * i], dy1 = p[2 * j + 1] - p[2 * i + 1];
        for (int k = 0; k < j; k++) {
            const double dx2 = p[2 * k] - p[2 * i], dy2 = p[2 * k + 1] - p[2 * i + 1];
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return 1;
        }
    }
    return 0;
}

static int fm_get_subset(const float* m1, const float* m2, int count, fm_rng* rng, int max_attempts, float* s1, float* s2) {
    int idx[7];
    for (int iters = 0; iters < max_attempts; iters++) {
        for (int i = 0; i < 7; i++) {
            int v, dup;
            do {
                v = fm_rng_uniform(rng, 0, count);
                dup = 0;
                for (int j = 0; j < i; j++) dup |= idx[j] == v;
            } while (dup);
            idx[i] = v;
            s1[2 * i] = m1[2 * v]; s1[2 * i + 1] = m1[2 * v + 1];
            s2[2 * i] = m2[2 * v]; s2[2 * i + 1] = m2[2 * v + 1];
        }
        if (!fm_collinear(s1, 7) && !fm_collinear(s2, 7)) return 1;
    }
    return 0;
}

static void fm_errors(const float* m1, const float* m2, int n, const double* F, float* err) {   /* FMEstimatorCallback::computeError */
    for (int i = 0; i < n; i++) {
        const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
        err[i] = (float)(e1 > e2 ? e1 : e2);
    }
}

static int fm_update_iters(double p, double ep, int model_points, int max_iters) {              /* RANSACUpdateNumIters */
    p = p > 0. ? p : 0.; p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.; ep = ep < 1. ? ep : 1.;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

static int fm_float_cmp(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* returns 0: no model (n < 7 or nothing found), else 1; F9 = the winning model (n == 7: the first root), mask[n] = inliers */
int spec_find_fundamental_mat(const float* m1, const float* m2, int n, double thresh, double conf, double* F9, uint8_t* mask) {
    const int max_iters = 1000;
    double F[27];
    if (n < 7) return 0;
    if (n == 7) {
        const int k = fm_run7(m1, m2, F);
        memset(mask, 1, (size_t)n);               /* mask.setTo(1) whatever the solver returns */
        if (k <= 0) return 0;
        memcpy(F9, F, sizeof(double) * 9);
        return 1;
    }
    if (thresh <= 0) thresh = 3;
    if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
    float* err = (float*)malloc(sizeof(float) * (size_t)n * 2);
    float* srt = err + n;
    float s1[14], s2[14];
    fm_rng rng = {(uint64_t)-1};
    int found = 0;
    if (n >= 15) {
        int niters = max_iters, max_good = 0;
        const float t = (float)(thresh * thresh);
        for (int iter = 0; iter < niters; iter++) {
            if (!fm_get_subset(m1, m2, n, &rng, 10000, s1, s2)) break;
            const int k = fm_run7(s1, s2, F);
            for (int i = 0; i < k; i++) {
                fm_errors(m1, m2, n, F + 9 * i, err);
                int good = 0;
                for (int j = 0; j < n; j++) good += err[j] <= t;
                if (good > (max_good > 6 ? max_good : 6)) {
                    for (int j = 0; j < n; j++) mask[j] = err[j] <= t;
                    memcpy(F9, F + 9 * i, sizeof(double) * 9);
                    max_good = good;
                    found = 1;
                    niters = fm_update_iters(conf, (double)(n - good) / n, 7, niters);
                }
            }
        }
    } else {
        int niters = fm_update_iters(conf, 0.45, 7, max_iters);
        if (niters < 3) niters = 3;
        double min_median = DBL_MAX;
        for (int iter = 0; iter < niters; iter++) {
            if (!fm_get_subset(m1, m2, n, &rng, 1000, s1, s2)) break;
            const int k = fm_run7(s1, s2, F);
            for (int i = 0; i < k; i++) {
                fm_errors(m1, m2, n, F + 9 * i, err);
                memcpy(srt, err, sizeof(float) * (size_t)n);
                qsort(srt, (size_t)n, sizeof(float), fm_float_cmp);
                const double median = srt[n / 2];
                if (median < min_median) { min_median = median; memcpy(F9, F + 9 * i, sizeof(double) * 9); found = 1; }
            }
        }
        if (found) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(min_median);
            if (sigma < 0.001) sigma = 0.001;
            const float t = (float)(sigma * sigma);
            fm_errors(m1, m2, n, F9, err);
            for (int j = 0; j < n; j++) mask[j] = err[j] <= t;
        }
    }
    free(err);
    return found;
}

/* InstsFeatManager::RejectWithF (dynamic_tracker.cpp:831-849): both point lists lifted with cam0 (liftProjective, z = 1),
 * re-projected with the virtual focal length kFocalLength = 460 (utils/parameters.h:42) about (col/2, row/2), narrowed to
 * float, then findFundamentalMat(un_cur, un_prev, FM_RANSAC, F_threshold, 0.99, status).  Returns the size of `status`
 * (n, or 0 when findFundamentalMat leaves it empty). */
int spec_reject_with_f(const double* cam, const float* cur_pts, const float* prev_pts, int n, int col, int row, double f_threshold,
                       uint8_t* status, float* un_out /* nullable: 4 n floats, un_cur then un_prev */) {
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], k1 = cam[4], k2 = cam[5], p1 = cam[6], p2 = cam[7];
    const double iK11 = 1.0 / fx, iK13 = -cx / fx, iK22 = 1.0 / fy, iK23 = -cy / fy;
    const int nod = (k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0);
    float* un = (float*)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
    for (int s = 0; s < 2; s++) {
        const float* pts = s ? prev_pts : cur_pts;
        for (int i = 0; i < n; i++) {
            const double u = (double)pts[2 * i], v = (double)pts[2 * i + 1];
            const double mxd = iK11 * u + iK13, myd = iK22 * v + iK23;
            double mxu = mxd, myu = myd;
            if (!nod) {
                for (int it = 0; it < 8; it++) {
                    const double x = mxu, y = myu;
                    const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
                    const double rad = k1 * rho2 + k2 * rho2 * rho2;
                    const double dux = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2);
                    const double duy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2);
                    mxu = mxd - dux; myu = myd - duy;
                }
            }
            un[(size_t)s * 2 * n + 2 * i] = (float)(460.0 * mxu / 1.0 + col / 2.0);
            un[(size_t)s * 2 * n + 2 * i + 1] = (float)(460.0 * myu / 1.0 + row / 2.0);
        }
    }
    if (un_out) memcpy(un_out, un, sizeof(float) * 4 * (size_t)n);
    double F9[9];
    const int ok = spec_find_fundamental_mat(un, un + 2 * (size_t)n, n, f_threshold, 0.99, F9, status);
    free(un);
    if (n < 7) return 0;
    if (!ok && n > 7) memset(status, 0, (size_t)n);   /* OpenCV creates the mask and leaves it unwritten when no model is found: 0 here */
    return n;
}

/* InstFeat::DetectExtraPoints (front_end/instance_feature.cpp:413-461): the instance mask is sampled on a `step` grid,
 * step = max(sqrt(0.8*rows*cols/1000), 2) truncated to int; a sample with mask > 0.5 and a finite positive disparity gives
 * depth = fx*baseline/disparity (float), kept when 0.1 < depth <= 100, x = (c-cx)*depth/fx, y = (r-cy)*depth/fy in float
 * arithmetic with the float camera constants of CameraInfo.  Returns the number of points (row-major sample order). */
int spec_detect_extra_points(const uint8_t* mask, int rows, int cols, int mask_pitch, const float* disp, int disp_pitch_elems,
                             int box_x, int box_y, float fx, float fy, float cx, float cy, float baseline, double* out /* 3 per point */) {
    double s = sqrt(0.8 * rows * cols / 1000.);
    const int step = (int)(s > 2. ? s : 2.);
    int n = 0;
    for (int i = 0; i < rows; i += step)
        for (int j = 0; j < cols; j += step) {
            if (mask[(size_t)i * mask_pitch + j] <= 0.5) continue;
            const int r = i + box_y, c = j + box_x;
            const float disparity = disp[(size_t)r * disp_pitch_elems + c];
            if (disparity <= 0) continue;
            if (disparity != disparity) continue;
            const float depth = fx * baseline / disparity;
            if (depth <= 0.1 || depth > 100) continue;
            const float x3 = (c - cx) * depth / fx, y3 = (r - cy) * depth / fy;
            out[3 * n] = x3; out[3 * n + 1] = y3; out[3 * n + 2] = depth;
            n++;
        }
    return n;
}
