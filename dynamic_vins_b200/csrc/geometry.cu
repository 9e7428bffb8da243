// Epipolar outlier rejection and the dense extra-point sampler of dynamic mode:
//   InstsFeatManager::RejectWithF   dynamic_vins/src/front_end/dynamic_tracker.cpp:831-849
//   FeatureTracker::RejectWithF     dynamic_vins/src/front_end/background_tracker.cpp:520-550 (same call, commented out there)
//   InstFeat::DetectExtraPoints     dynamic_vins/src/front_end/instance_feature.cpp:413-461
//
// RejectWithF = liftProjective of both point lists (fp64), re-projection with kFocalLength = 460 about the image centre,
// narrowing to float, then cv::findFundamentalMat(cur, prev, FM_RANSAC, F_threshold, 0.99, status).  OpenCV
// (modules/calib3d/src/{fundam,ptsetreg}.cpp) runs that as a SEQUENTIAL loop: draw a 7-point sample from cv::RNG(-1), solve the
// cubic of the 7-point algorithm (up to 3 models), count the inliers of each over all points, keep the best and shorten the
// loop from the inlier ratio (8 <= n < 15: LMedS, the model with the smallest median error).  Which sample is drawn at iteration
// k depends only on the RNG and the point coordinates, never on the models, so the loop splits into
//   (a) one thread replaying the RNG: a chunk of candidate samples (index 7-tuples),
//   (b) one thread per candidate sample: collinearity test, 7-point solve, inlier count / median of every root over all points,
//   (c) one thread replaying OpenCV's bookkeeping over the chunk in sample order (best model, iteration budget, give-up rule),
// repeated until (c) says the sequential loop would have stopped.  One CTA per problem; all arithmetic in fp64 without
// contraction (-fmad=false) in the operation order of the OpenCV source, so the status bytes equal the CPU loop's.
#include "kernels.cuh"
#include "tracker.h"

#define DVFE_CHECK(call)                  \
    do {                                  \
        int rc__ = (call);                \
        if (rc__ != DVFE_OK) return rc__; \
    } while (0)

#define FM_THREADS 256          // candidate samples per chunk = threads per CTA
#define FM_MODEL_POINTS 7
#define FM_MAX_ITERS 1000       // cv::findFundamentalMat's maxIters

namespace {

struct FmRng {                  // cv::RNG: multiply-with-carry
    unsigned long long state;
    __device__ unsigned next() {
        state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

__device__ int fm_solve_cubic(const double* c, double* x) {          // cv::solveCubic
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
            const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
            if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
            else { x[0] = q2 / a1; x[1] = a3 / q2; }
            return d > 0 ? 2 : 1;
        }
        return 0;
    }
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qcubed));
        const double sqrtQ = sqrt(Q);
        const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
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

// Null space of the 7 x 9 epipolar system by Gauss-Jordan elimination with complete pivoting; f1 / f2 = the solutions with the
// two free unknowns set to (1, 0) / (0, 1).  (OpenCV takes an orthonormal basis from its SVD; the set of F's is the same.)
__device__ bool fm_null_space(double (*A)[9], double* f1, double* f2) {
    int perm[9];
    for (int j = 0; j < 9; j++) perm[j] = j;
    for (int k = 0; k < 7; k++) {
        int pr = k, pc = k;
        double best = -1.0;
        for (int i = k; i < 7; i++)
            for (int j = k; j < 9; j++)
                if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
        if (!(best > 0.0)) return false;
        if (pr != k) for (int j = 0; j < 9; j++) { const double t = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = t; }
        if (pc != k) {
            for (int i = 0; i < 7; i++) { const double t = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = t; }
            const int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t;
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
    return true;
}

// run7Point on the sample (m1[idx], m2[idx]): up to 3 fundamental matrices, each scaled to F[8] = 1
__device__ int fm_run7(const float2* __restrict__ m1, const float2* __restrict__ m2, const int* idx, double* F) {
    double A[7][9], f1[9], f2[9], c[4], r[3] = {0, 0, 0};
    for (int i = 0; i < 7; i++) {
        const float2 a = m1[idx[i]], b = m2[idx[i]];
        const double x0 = a.x, y0 = a.y, x1 = b.x, y1 = b.y;
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
    if (n < 1 || n > 3) return 0;
    for (int k = 0; k < n; k++, F += 9) {
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > 2.220446049250313e-16) { mu = 1. / s; lambda *= mu; F[8] = 1.; }
        else F[8] = 0.;
        for (int i = 0; i < 8; i++) F[i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}

// haveCollinearPoints on the sample: only its last point is tested against the pairs before it
__device__ bool fm_collinear(const float2* __restrict__ m, const int* idx) {
    const float2 pi = m[idx[6]];
    for (int j = 0; j < 6; j++) {
        const float2 pj = m[idx[j]];
        const double dx1 = pj.x - pi.x, dy1 = pj.y - pi.y;      // float subtraction widened, as `double dx1 = ptr[j].x - ptr[i].x`
        for (int k = 0; k < j; k++) {
            const float2 pk = m[idx[k]];
            const double dx2 = pk.x - pi.x, dy2 = pk.y - pi.y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= 1.1920928955078125e-07 * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}

// FMEstimatorCallback::computeError of one correspondence (symmetric squared epipolar distance, narrowed to float)
__device__ __forceinline__ float fm_error(const double* F, float2 p1, float2 p2) {
    const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
    const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
    a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
    const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
    const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    return (float)(e1 > e2 ? e1 : e2);
}

__device__ int fm_update_iters(double p, double ep, int model_points, int max_iters) {      // RANSACUpdateNumIters
    p = p > 0. ? p : 0.; p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.; ep = ep < 1. ? ep : 1.;
    double num = 1. - p > 2.2250738585072014e-308 ? 1. - p : 2.2250738585072014e-308;
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num); denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)llrint(num / denom);
}

struct FmProblem {
    const float2* cur;       // distorted pixel positions, current frame  (findFundamentalMat's points1 after undistortion)
    const float2* prev;      // ... previous frame                         (points2)
    float2* un;              // scratch: 2 n undistorted / re-projected points (cur then prev)
    double* models;          // scratch: FM_THREADS * 27 doubles
    uint8_t* status;         // out: n bytes
    int n;
};

// One CTA per problem.
__global__ void __launch_bounds__(FM_THREADS) k_reject_with_f(const FmProblem* __restrict__ problems, CamParams cam, double half_col,
                                                             double half_row, double threshold, double confidence) {
    __shared__ int s_idx[FM_THREADS][FM_MODEL_POINTS];
    __shared__ int s_nmodels[FM_THREADS];            // -1: sample rejected (collinear), else number of roots
    __shared__ double s_score[FM_THREADS][3];        // RANSAC: inlier count; LMedS: median error
    __shared__ int s_done, s_found, s_best_slot, s_best_root;
    __shared__ double s_best[9], s_min_median;
    const FmProblem P = problems[blockIdx.x];
    const int n = P.n, tid = threadIdx.x;
    if (n < 7) return;                               // no model: the caller reports an empty status

    // undistort + re-project with the virtual focal length (float, as cv::Point2f((float)x, (float)y))
    for (int i = tid; i < 2 * n; i += FM_THREADS) {
        const float2 p = i < n ? P.cur[i] : P.prev[i - n];
        const double u = (double)p.x, v = (double)p.y;
        const double mx_d = cam.inv_K11 * u + cam.inv_K13, my_d = cam.inv_K22 * v + cam.inv_K23;
        double mx_u = mx_d, my_u = my_d;
        if (!cam.no_distortion) {
#pragma unroll 1
            for (int it = 0; it < 8; it++) {
                const double x = mx_u, y = my_u;
                const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
                const double rad = cam.k1 * rho2 + cam.k2 * rho2 * rho2;
                const double dux = x * rad + 2.0 * cam.p1 * mxy + cam.p2 * (rho2 + 2.0 * mx2);
                const double duy = y * rad + 2.0 * cam.p2 * mxy + cam.p1 * (rho2 + 2.0 * my2);
                mx_u = mx_d - dux; my_u = my_d - duy;
            }
        }
        P.un[i] = make_float2((float)(460.0 * mx_u / 1.0 + half_col), (float)(460.0 * my_u / 1.0 + half_row));
    }
    if (n == 7) {                                    // the solver runs once and the mask is set to all ones
        if (tid < 7) P.status[tid] = 1;
        return;
    }
    __syncthreads();
    const float2* __restrict__ m1 = P.un;
    const float2* __restrict__ m2 = P.un + n;
    const bool lmeds = n < 15;
    const float t_ransac = (float)(threshold * threshold);

    // thread 0 carries the sequential state of the OpenCV loop
    FmRng rng{~0ull};
    int niters = FM_MAX_ITERS, iter = 0, max_good = 0, failed_attempts = 0;
    bool found = false;
    if (tid == 0) {
        s_done = 0; s_min_median = 1.7976931348623157e308;
        if (lmeds) { niters = fm_update_iters(confidence, 0.45, FM_MODEL_POINTS, FM_MAX_ITERS); if (niters < 3) niters = 3; }
    }
    const int max_attempts = lmeds ? 1000 : 10000;   // getSubset's give-up bound per iteration
    double* my_models = P.models + (size_t)tid * 27;

    for (;;) {
        // (a) the RNG stream: FM_THREADS candidate samples, duplicates inside a sample redrawn
        if (tid == 0) {
            for (int a = 0; a < FM_THREADS; a++)
                for (int i = 0; i < FM_MODEL_POINTS; i++) {
                    int v;
                    bool dup;
                    do {
                        v = rng.uniform(0, n);
                        dup = false;
                        for (int j = 0; j < i; j++) dup |= s_idx[a][j] == v;
                    } while (dup);
                    s_idx[a][i] = v;
                }
        }
        __syncthreads();
        // (b) one candidate sample per thread
        {
            int idx[7];
            for (int i = 0; i < 7; i++) idx[i] = s_idx[tid][i];
            int k = -1;
            if (!fm_collinear(m1, idx) && !fm_collinear(m2, idx)) {
                double F[27];
                k = fm_run7(m1, m2, idx, F);
                for (int r = 0; r < k; r++) {
                    const double* Fr = F + 9 * r;
                    if (!lmeds) {
                        int good = 0;
                        for (int j = 0; j < n; j++) good += fm_error(Fr, m1[j], m2[j]) <= t_ransac;
                        s_score[tid][r] = (double)good;
                    } else {
                        float e[14];
                        for (int j = 0; j < n; j++) {            // insertion sort, n <= 14
                            const float v = fm_error(Fr, m1[j], m2[j]);
                            int q = j;
                            while (q > 0 && e[q - 1] > v) { e[q] = e[q - 1]; q--; }
                            e[q] = v;
                        }
                        s_score[tid][r] = (double)e[n / 2];       // std::nth_element(..., count / 2)
                    }
                    for (int q = 0; q < 9; q++) my_models[9 * r + q] = Fr[q];
                }
            }
            s_nmodels[tid] = k;
        }
        __syncthreads();
        // (c) OpenCV's loop bookkeeping replayed over the chunk in sample order
        if (tid == 0) {
            int best_slot = -1, best_root = 0;
            for (int a = 0; a < FM_THREADS && !s_done; a++) {
                if (iter >= niters) { s_done = 1; break; }
                const int k = s_nmodels[a];
                if (k < 0) {                                     // getSubset retries; gives up after max_attempts
                    if (++failed_attempts >= max_attempts) s_done = 1;
                    continue;
                }
                failed_attempts = 0;
                for (int r = 0; r < k; r++) {
                    if (!lmeds) {
                        const int good = (int)s_score[a][r];
                        if (good > (max_good > FM_MODEL_POINTS - 1 ? max_good : FM_MODEL_POINTS - 1)) {
                            max_good = good; best_slot = a; best_root = r; found = true;
                            niters = fm_update_iters(confidence, (double)(n - good) / n, FM_MODEL_POINTS, niters);
                        }
                    } else if (s_score[a][r] < s_min_median) {
                        s_min_median = s_score[a][r]; best_slot = a; best_root = r; found = true;
                    }
                }
                iter++;
            }
            if (iter >= niters) s_done = 1;
            s_best_slot = best_slot; s_best_root = best_root; s_found = found ? 1 : 0;
        }
        __syncthreads();
        if (s_best_slot >= 0 && tid < 9) s_best[tid] = P.models[(size_t)s_best_slot * 27 + 9 * s_best_root + tid];
        __syncthreads();          // s_best is taken before the next chunk overwrites the model scratch
        if (s_done) break;
    }
    if (!s_found) {                // no model: OpenCV leaves the mask unwritten; reported as all zero
        for (int j = tid; j < n; j += FM_THREADS) P.status[j] = 0;
        return;
    }
    float t = t_ransac;
    if (lmeds) {
        double sigma = 2.5 * 1.4826 * (1 + 5. / (n - FM_MODEL_POINTS)) * sqrt(s_min_median);
        if (sigma < 0.001) sigma = 0.001;
        t = (float)(sigma * sigma);
    }
    for (int j = tid; j < n; j += FM_THREADS) P.status[j] = fm_error(s_best, m1[j], m2[j]) <= t ? 1 : 0;
}

// InstFeat::DetectExtraPoints: one thread per grid sample, order-preserving compaction (row-major sample order) by a block scan
struct ExtraJob {
    const uint8_t* mask;     // ROI mask, rows x cols
    int mask_pitch;
    int rows, cols;
    int box_x, box_y;        // box2d->rect.tl()
    double* out;             // 3 doubles per point
    int* n_out;
};

__global__ void __launch_bounds__(256) k_detect_extra_points(const ExtraJob* __restrict__ jobs, const float* __restrict__ disp, int disp_pitch,
                                                              float fx, float fy, float cx, float cy, float baseline) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const ExtraJob J = jobs[blockIdx.x];
    const int step = (int)fmax(sqrt(0.8 * J.rows * J.cols / 1000.), 2.);
    const int gw = (J.cols + step - 1) / step, gh = (J.rows + step - 1) / step, total = gw * gh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 256) {
        const int s = base + tid;
        bool keep = false;
        float x3 = 0.f, y3 = 0.f, depth = 0.f;
        if (s < total) {
            const int i = (s / gw) * step, j = (s - (s / gw) * gw) * step;
            if (J.mask[(size_t)i * J.mask_pitch + j] > 0) {
                const int r = i + J.box_y, c = j + J.box_x;
                const float d = disp[(size_t)r * disp_pitch + c];
                if (d > 0.f && d == d) {
                    depth = fx * baseline / d;
                    if (!((double)depth <= 0.1 || depth > 100.f)) {
                        x3 = ((float)c - cx) * depth / fx;
                        y3 = ((float)r - cy) * depth / fy;
                        keep = true;
                    }
                }
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) { J.out[3 * (size_t)off] = x3; J.out[3 * (size_t)off + 1] = y3; J.out[3 * (size_t)off + 2] = depth; }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) *J.n_out = s_base;
}

struct DevMem {
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        DVFE_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return DVFE_OK;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

int need_device() {
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        dvfe_set_error("no CUDA device available (%s): libdvfe has no CPU fallback", cudaGetErrorString(e));
        return DVFE_ERR_NO_DEVICE;
    }
    return DVFE_OK;
}
}  // namespace

extern "C" int dvfe_op_reject_with_f(const dvfe_camera* cam, const float* cur_pts, const float* prev_pts, int n, int col, int row,
                                     double f_threshold, uint8_t* status, int* n_status) {
    if (!cam || n < 0 || !n_status || (n > 0 && (!cur_pts || !prev_pts || !status))) {
        dvfe_set_error("op_reject_with_f: bad argument");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(need_device());
    *n_status = 0;
    if (n < 7) return DVFE_OK;                       // cv::findFundamentalMat returns an empty matrix and leaves `status` empty
    DevMem d_cur, d_prev, d_un, d_models, d_status, d_prob;
    DVFE_CHECK(d_cur.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_prev.alloc(sizeof(float2) * n));
    DVFE_CHECK(d_un.alloc(sizeof(float2) * 2 * n));
    DVFE_CHECK(d_models.alloc(sizeof(double) * 27 * FM_THREADS));
    DVFE_CHECK(d_status.alloc(n));
    DVFE_CHECK(d_prob.alloc(sizeof(FmProblem)));
    DVFE_CUDA(cudaMemcpy(d_cur.p, cur_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    DVFE_CUDA(cudaMemcpy(d_prev.p, prev_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    FmProblem P{d_cur.as<float2>(), d_prev.as<float2>(), d_un.as<float2>(), d_models.as<double>(), d_status.as<uint8_t>(), n};
    DVFE_CUDA(cudaMemcpy(d_prob.p, &P, sizeof(P), cudaMemcpyHostToDevice));
    double thr = f_threshold;
    if (thr <= 0) thr = 3;                           // cv::findFundamentalMat: param1 <= 0 -> 3
    DVFE_LAUNCH(k_reject_with_f, 1, FM_THREADS, 0, 0, d_prob.as<FmProblem>(), make_cam(*cam), col / 2.0, row / 2.0, thr, 0.99);
    DVFE_CUDA(cudaGetLastError());
    DVFE_CUDA(cudaDeviceSynchronize());
    DVFE_CUDA(cudaMemcpy(status, d_status.p, n, cudaMemcpyDeviceToHost));
    *n_status = n;
    return DVFE_OK;
}

extern "C" int dvfe_op_detect_extra_points(const uint8_t* roi_mask, int rows, int cols, int mask_pitch, const float* disp, int disp_w,
                                           int disp_h, int disp_pitch, int box_x, int box_y, float fx, float fy, float cx, float cy,
                                           float baseline, double* out, int cap, int* n_out) {
    if (!roi_mask || !disp || !out || !n_out || rows < 1 || cols < 1 || mask_pitch < cols || disp_pitch < disp_w || box_x < 0 || box_y < 0 ||
        box_x + cols > disp_w || box_y + rows > disp_h) {
        dvfe_set_error("op_detect_extra_points: bad argument (the box must lie inside the disparity map)");
        return DVFE_ERR_INVALID;
    }
    DVFE_CHECK(need_device());
    const int step = (int)fmax(sqrt(0.8 * rows * cols / 1000.), 2.);
    const int total = ((cols + step - 1) / step) * ((rows + step - 1) / step);
    DevMem d_mask, d_disp, d_out, d_n, d_job;
    DVFE_CHECK(d_mask.alloc((size_t)rows * cols));
    DVFE_CHECK(d_disp.alloc(sizeof(float) * (size_t)disp_w * disp_h));
    DVFE_CHECK(d_out.alloc(sizeof(double) * 3 * total));
    DVFE_CHECK(d_n.alloc(sizeof(int)));
    DVFE_CHECK(d_job.alloc(sizeof(ExtraJob)));
    DVFE_CUDA(cudaMemcpy2D(d_mask.p, cols, roi_mask, mask_pitch, cols, rows, cudaMemcpyHostToDevice));
    DVFE_CUDA(cudaMemcpy2D(d_disp.p, sizeof(float) * disp_w, disp, sizeof(float) * disp_pitch, sizeof(float) * disp_w, disp_h,
                           cudaMemcpyHostToDevice));
    ExtraJob J{d_mask.as<uint8_t>(), cols, rows, cols, box_x, box_y, d_out.as<double>(), d_n.as<int>()};
    DVFE_CUDA(cudaMemcpy(d_job.p, &J, sizeof(J), cudaMemcpyHostToDevice));
    DVFE_LAUNCH(k_detect_extra_points, 1, 256, 0, 0, d_job.as<ExtraJob>(), d_disp.as<float>(), disp_w, fx, fy, cx, cy, baseline);
    DVFE_CUDA(cudaGetLastError());
    DVFE_CUDA(cudaDeviceSynchronize());
    int n = 0;
    DVFE_CUDA(cudaMemcpy(&n, d_n.p, sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > cap) {
        dvfe_set_error("op_detect_extra_points: %d points do not fit the output capacity %d", n, cap);
        return DVFE_ERR_CAPACITY;
    }
    DVFE_CUDA(cudaMemcpy(out, d_out.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    return DVFE_OK;
}
