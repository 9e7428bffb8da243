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
#include <math.h>

#include "kernels.cuh"

#define RESP_TW 32
#define RESP_TH 16
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
__global__ void __launch_bounds__(256) k_gftt_mask_fill(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J)) return;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 8 && threadIdx.y == 0) {
        // reset the per-job counters: n_cand, max, overflow, n_accepted, new_cnt ...
        J.counters[threadIdx.x] = (threadIdx.x == 1) ? INT_MIN : 0;
    }
    if (x >= J.w || y >= J.h) return;
    uint8_t* m = J.mask + (size_t)y * J.mask_pitch + x;
    if (J.region_mask == nullptr) {
        if (x + 3 < J.w && (J.mask_pitch & 3) == 0) *reinterpret_cast<uint32_t*>(m) = 0xffffffffu;
        else for (int i = 0; i < 4 && x + i < J.w; i++) m[i] = 255;
    } else {
        const uint8_t* r = J.region_mask + (size_t)y * J.region_pitch + x;
        for (int i = 0; i < 4 && x + i < J.w; i++) m[i] = r[i];
    }
}

// ---- ... minus a filled disc around every tracked point:  cleared <=> dx^2+dy^2 <= r^2 ------------------
__global__ void __launch_bounds__(128) k_gftt_discs(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.y];
    if (!gftt_job_active(J)) return;
    const int i = blockIdx.x;
    if (i >= *J.n) return;
    const float2 p = J.pts[i];
    const int cx = __float2int_rn(p.x), cy = __float2int_rn(p.y);
    const int r = J.disc_radius, d = 2 * r + 1, r2 = r * r;
    for (int t = threadIdx.x; t < d * d; t += blockDim.x) {
        const int dy = t / d - r, dx = t - (t / d) * d - r;
        const int x = cx + dx, y = cy + dy;
        if (x < 0 || x >= J.w || y < 0 || y >= J.h) continue;
        if (dx * dx + dy * dy <= r2) J.mask[(size_t)y * J.mask_pitch + x] = 0;
    }
}

__global__ void __launch_bounds__(128) k_disc_mask_op(uint8_t* mask, int pitch, int w, int h, const float2* pts, const int* n,
                                                      int r) {
    const int i = blockIdx.x;
    if (i >= *n) return;
    const float2 p = pts[i];
    const int cx = __float2int_rn(p.x), cy = __float2int_rn(p.y);
    const int d = 2 * r + 1, r2 = r * r;
    for (int t = threadIdx.x; t < d * d; t += blockDim.x) {
        const int dy = t / d - r, dx = t - (t / d) * d - r;
        const int x = cx + dx, y = cy + dy;
        if (x < 0 || x >= w || y < 0 || y >= h) continue;
        if (dx * dx + dy * dy <= r2) mask[(size_t)y * pitch + x] = 0;
    }
}

// ---- response map (cv::cornerMinEigenVal, blockSize 3, ksize 3) + masked max -----------------------
__device__ __forceinline__ void resp_tile(const uint8_t* __restrict__ img, int pitch, int w, int h, int tx0, int ty0,
                                          float* __restrict__ eig, const uint8_t* __restrict__ mask, int mask_pitch,
                                          int* max_out) {
    __shared__ uint8_t s_img[RESP_TH + 4][RESP_TW + 4];
    __shared__ float s_dx[RESP_TH + 2][RESP_TW + 2];
    __shared__ float s_dy[RESP_TH + 2][RESP_TW + 2];
    __shared__ int s_max[8];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float s2 = s * 2.0f;

    for (int i = tid; i < (RESP_TH + 4) * (RESP_TW + 4); i += nthr) {
        const int r = i / (RESP_TW + 4), c = i - r * (RESP_TW + 4);
        const int gx = reflect101(tx0 - 2 + c, w), gy = reflect101(ty0 - 2 + r, h);
        s_img[r][c] = __ldg(img + (size_t)gy * pitch + gx);
    }
    __syncthreads();
    // Sobel at in-image positions of the tile + 1 halo
    for (int i = tid; i < (RESP_TH + 2) * (RESP_TW + 2); i += nthr) {
        const int r = i / (RESP_TW + 2), c = i - r * (RESP_TW + 2);
        const int gx = tx0 - 1 + c, gy = ty0 - 1 + r;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            // image taps: s_img[r + dy + 1][c + dx + 1], dy,dx in {-1,0,1} -> rows r..r+2, cols c..c+2
            const float a00 = s_img[r][c], a01 = s_img[r][c + 1], a02 = s_img[r][c + 2];
            const float a10 = s_img[r + 1][c], a11 = s_img[r + 1][c + 1], a12 = s_img[r + 1][c + 2];
            const float a20 = s_img[r + 2][c], a21 = s_img[r + 2][c + 1], a22 = s_img[r + 2][c + 2];
            // Dx: row kernel [-1 0 1] exact, column kernel [1 2 1]*scale -> fma(s, d0 + d2, (2s)*d1)
            const float d0 = a02 - a00, d1 = a12 - a10, d2 = a22 - a20;
            s_dx[r][c] = __fmaf_rn(s, __fadd_rn(d0, d2), __fmul_rn(s2, d1));
            // Dy: row kernel [1 2 1]*scale -> fma(s, r, fma(2s, c, s*l)); column kernel [-1 0 1]
            const float top = __fmaf_rn(s, a02, __fmaf_rn(s2, a01, __fmul_rn(s, a00)));
            const float bot = __fmaf_rn(s, a22, __fmaf_rn(s2, a21, __fmul_rn(s, a20)));
            s_dy[r][c] = __fsub_rn(bot, top);
            (void)a11;
        }
    }
    __syncthreads();
    // positions outside the image take the derivative of their REFLECT_101 position (box filter border)
    for (int i = tid; i < (RESP_TH + 2) * (RESP_TW + 2); i += nthr) {
        const int r = i / (RESP_TW + 2), c = i - r * (RESP_TW + 2);
        const int gx = tx0 - 1 + c, gy = ty0 - 1 + r;
        if (!(gx >= 0 && gx < w && gy >= 0 && gy < h)) {
            const int rx = reflect101(gx, w), ry = reflect101(gy, h);
            const int rr = ry - (ty0 - 1), rc = rx - (tx0 - 1);
            if (rr >= 0 && rr < RESP_TH + 2 && rc >= 0 && rc < RESP_TW + 2) {
                s_dx[r][c] = s_dx[rr][rc];
                s_dy[r][c] = s_dy[rr][rc];
            } else {   // only reachable for positions that no in-image output pixel of this tile uses
                s_dx[r][c] = 0.f;
                s_dy[r][c] = 0.f;
            }
        }
    }
    __syncthreads();
    int best = INT_MIN;
    for (int i = tid; i < RESP_TH * RESP_TW; i += nthr) {
        const int r = i / RESP_TW, c = i - r * RESP_TW;
        const int gx = tx0 + c, gy = ty0 + r;
        if (gx < w && gy < h) {
            double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float dx = s_dx[r + j][c + k], dy = s_dy[r + j][c + k];
                    sxx += (double)__fmul_rn(dx, dx);
                    sxy += (double)__fmul_rn(dx, dy);
                    syy += (double)__fmul_rn(dy, dy);
                }
            const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, cc = __fmul_rn((float)syy, 0.5f);
            const float amc = __fsub_rn(a, cc);
            const float lam = __fsub_rn(__fadd_rn(a, cc), sqrtf(__fadd_rn(__fmul_rn(amc, amc), __fmul_rn(b, b))));
            eig[(size_t)gy * w + gx] = lam;
            if (max_out != nullptr && (mask == nullptr || mask[(size_t)gy * mask_pitch + gx] != 0)) {
                const int o = f2ord(lam);
                best = o > best ? o : best;
            }
        }
    }
    if (max_out != nullptr) {
        best = __reduce_max_sync(0xffffffffu, best);
        if ((tid & 31) == 0) s_max[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            int m = s_max[0];
            for (int i = 1; i < nthr / 32; i++) m = s_max[i] > m ? s_max[i] : m;
            if (m != INT_MIN) atomicMax(max_out, m);
        }
    }
}

__global__ void __launch_bounds__(256) k_gftt_response(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J)) return;
    const int tx0 = blockIdx.x * RESP_TW, ty0 = blockIdx.y * RESP_TH;
    if (tx0 >= J.w || ty0 >= J.h) return;
    if (J.eig_in != nullptr) {
        // externally supplied response map: only the masked max is needed
        int best = INT_MIN;
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        for (int i = tid; i < RESP_TH * RESP_TW; i += blockDim.x * blockDim.y) {
            const int gx = tx0 + i % RESP_TW, gy = ty0 + i / RESP_TW;
            if (gx < J.w && gy < J.h && J.mask[(size_t)gy * J.mask_pitch + gx] != 0) {
                const int o = f2ord(J.eig_in[(size_t)gy * J.w + gx]);
                best = o > best ? o : best;
            }
        }
        best = __reduce_max_sync(0xffffffffu, best);
        if ((tid & 31) == 0 && best != INT_MIN) atomicMax(&J.counters[1], best);
        return;
    }
    resp_tile(J.img, J.img_pitch, J.w, J.h, tx0, ty0, J.eig, J.mask, J.mask_pitch, &J.counters[1]);
}

__global__ void __launch_bounds__(256) k_min_eigen_val(const uint8_t* img, int pitch, int w, int h, float* eig) {
    resp_tile(img, pitch, w, h, blockIdx.x * RESP_TW, blockIdx.y * RESP_TH, eig, nullptr, 0, nullptr);
}

// ---- candidates: threshold, 3x3 local max, mask -----------------------------------------------------
__global__ void __launch_bounds__(256) k_gftt_candidates(const GfttJob* __restrict__ jobs) {
    const GfttJob& J = jobs[blockIdx.z];
    if (!gftt_job_active(J)) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const float* __restrict__ eig = J.eig_in ? J.eig_in : J.eig;
    const int mo = J.counters[1];
    const double maxVal = (mo == INT_MIN) ? 0.0 : (double)ord2f(mo);
    const float thr = (float)(maxVal * J.quality);
    bool is_cand = false;
    float v = 0.f;
    if (x >= 1 && x < J.w - 1 && y >= 1 && y < J.h - 1) {
        v = eig[(size_t)y * J.w + x];
        if (v > thr && v != 0.f && J.mask[(size_t)y * J.mask_pitch + x] != 0) {
            // dilate(3x3) of the thresholded map equals v  <=>  no neighbour above both thr and v
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

// ---- selection: parallel greedy min-distance suppression + top-K, one CTA per job ------------------------
__device__ __forceinline__ int block_excl_scan_inplace(int* data, int n, int* s_part /* [NMS_THREADS] */) {
    // exclusive scan of data[0..n) in place, returns the total; all threads of the block participate
    const int tid = threadIdx.x;
    const int per = (n + NMS_THREADS - 1) / NMS_THREADS;
    const int b = tid * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += data[i];
    s_part[tid] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the partial sums
    for (int off = 1; off < NMS_THREADS; off <<= 1) {
        const int v = (tid >= off) ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    const int total = s_part[NMS_THREADS - 1];
    int run = s_part[tid] - sum;
    for (int i = b; i < e; i++) {
        const int v = data[i];
        data[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(NMS_THREADS) k_gftt_select(const GfttJob* __restrict__ jobs) {
    __shared__ int s_part[NMS_THREADS];
    __shared__ unsigned long long s_sel[NMS_MAX_K];
    __shared__ int s_hist[256];
    __shared__ int s_cnt;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining;

    const GfttJob& J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    if (tid == 0) J.counters[4] = 0;       // new_cnt
    if (!gftt_job_active(J)) return;
    const int n_old = *J.n;
    int K = J.max_cnt - n_old;
    if (K > NMS_MAX_K) K = NMS_MAX_K;
    int nc = J.counters[0];
    if (nc > J.cand_cap) nc = J.cand_cap;
    if (nc <= 0) return;

    const int w = J.w;
    const int cell = (int)lrintf(J.min_dist) > 0 ? (int)lrintf(J.min_dist) : 1;
    const int gw = (J.w + cell - 1) / cell, gh = (J.h + cell - 1) / cell;
    const int ncell = gw * gh;
    int* cstart = J.cell_count;                 // [ncell + 1]
    int* ccur = J.cell_count + (ncell + 1);     // [ncell + 1]
    const double md2 = (double)J.min_dist * (double)J.min_dist;
    const bool use_nms = J.min_dist >= 1.f;

    unsigned long long* sorted = J.cand2;        // candidates grouped by cell, descending key inside a cell
    volatile uint8_t* state = J.state;           // 0 undecided, 1 accepted, 2 rejected

    if (use_nms) {
        for (int i = tid; i < 2 * (ncell + 1); i += NMS_THREADS) cstart[i] = 0;
        __syncthreads();
        for (int i = tid; i < nc; i += NMS_THREADS) {
            const unsigned idx = (unsigned)J.cand[i];
            const int y = idx / w, x = idx - y * w;
            atomicAdd(&cstart[(y / cell) * gw + x / cell], 1);
        }
        __syncthreads();
        block_excl_scan_inplace(cstart, ncell + 1, s_part);
        // scatter into cells (unordered), then rank-sort each cell back into J.cand
        for (int i = tid; i < nc; i += NMS_THREADS) {
            const unsigned long long key = J.cand[i];
            const unsigned idx = (unsigned)key;
            const int y = idx / w, x = idx - y * w;
            const int c = (y / cell) * gw + x / cell;
            const int pos = atomicAdd(&ccur[c], 1);
            sorted[cstart[c] + pos] = key;
        }
        __syncthreads();
        {
            const int warp = tid >> 5, lane = tid & 31;
            for (int c = warp; c < ncell; c += NMS_THREADS / 32) {
                const int b = cstart[c], m = cstart[c + 1] - b;
                for (int e = lane; e < m; e += 32) {
                    const unsigned long long key = sorted[b + e];
                    int rank = 0;
                    for (int o = 0; o < m; o++) rank += (sorted[b + o] > key) ? 1 : 0;
                    J.cand[b + rank] = key;
                }
            }
        }
        __syncthreads();
        // J.cand is now cell-grouped and sorted; decide by rounds
        const unsigned long long* __restrict__ cs = J.cand;
        for (int i = tid; i < nc; i += NMS_THREADS) state[i] = 0;
        __syncthreads();
        for (;;) {
            int undecided = 0;
            for (int i = tid; i < nc; i += NMS_THREADS) {
                if (state[i] != 0) continue;
                const unsigned long long key = cs[i];
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
                            const unsigned long long kj = cs[j];
                            if (kj <= key) break;                 // only stronger candidates matter
                            const unsigned ij = (unsigned)kj;
                            const int yj = ij / w, xj = ij - yj * w;
                            const int dx = x - xj, dy = y - yj;
                            if ((double)(dx * dx + dy * dy) < md2) {
                                const uint8_t sj = state[j];
                                if (sj == 1) { rejected = true; break; }
                                if (sj == 0) blocked = true;
                            }
                        }
                    }
                if (rejected) state[i] = 2;
                else if (!blocked) state[i] = 1;
                else undecided = 1;
            }
            if (!__syncthreads_or(undecided)) break;
        }
    } else {
        for (int i = tid; i < nc; i += NMS_THREADS) state[i] = 1;
        __syncthreads();
    }

    // ---- collect accepted keys into `sorted` (unordered) ----
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nc; i0 += NMS_THREADS) {
        const int i = i0 + tid;
        const bool acc = (i < nc) && state[i] == 1;
        const unsigned ballot = __ballot_sync(0xffffffffu, acc);
        int base = 0;
        if ((tid & 31) == 0 && ballot) base = atomicAdd(&s_cnt, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (acc) sorted[base + __popc(ballot & ((1u << (tid & 31)) - 1))] = J.cand[i];
    }
    __syncthreads();
    const int n_acc = s_cnt;
    int n_sel = n_acc;
    unsigned long long kth = 0;       // keep keys >= kth
    if (n_acc > K) {
        // radix select of the K-th largest 64-bit key, 8 bits per pass from the top
        if (tid == 0) { s_prefix = 0; s_remaining = K; }
        __syncthreads();
        for (int pass = 0; pass < 8; pass++) {
            const int shift = 56 - 8 * pass;
            if (tid < 256) s_hist[tid] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long himask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
            for (int i = tid; i < n_acc; i += NMS_THREADS) {
                const unsigned long long k = sorted[i];
                if ((k & himask) == prefix) atomicAdd(&s_hist[(int)((k >> shift) & 255)], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int rem = s_remaining, b = 255;
                for (; b > 0; b--) {
                    if (s_hist[b] >= rem) break;
                    rem -= s_hist[b];
                }
                s_remaining = rem;
                s_prefix = prefix | ((unsigned long long)b << shift);
            }
            __syncthreads();
        }
        kth = s_prefix;
        n_sel = K;
    }
    // ---- gather the selected keys into shared memory, bitonic sort descending ----
    int npow = 1;
    while (npow < n_sel) npow <<= 1;
    for (int i = tid; i < npow; i += NMS_THREADS) s_sel[i] = 0ull;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int i = tid; i < n_acc; i += NMS_THREADS) {
        const unsigned long long k = sorted[i];
        if (k >= kth) {
            const int pos = atomicAdd(&s_cnt, 1);
            if (pos < NMS_MAX_K) s_sel[pos] = k;
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
    if (tid == 0) { J.counters[4] = n_sel; J.counters[3] = n_acc; }
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

int launch_gftt(const GfttJob* d_jobs, const GfttJob* h_jobs, int n_jobs, int max_w, int max_h, int max_pts,
                cudaStream_t st) {
    (void)h_jobs;
    if (n_jobs <= 0) return DVFE_OK;
    {
        dim3 blk(32, 8), grid(((max_w + 3) / 4 + 31) / 32, (max_h + 7) / 8, n_jobs);
        DVFE_LAUNCH(k_gftt_mask_fill, grid, blk, 0, st, d_jobs);
    }
    if (max_pts > 0) {
        dim3 grid(max_pts, n_jobs);
        DVFE_LAUNCH(k_gftt_discs, grid, 128, 0, st, d_jobs);
    }
    {
        dim3 blk(32, 8), grid((max_w + RESP_TW - 1) / RESP_TW, (max_h + RESP_TH - 1) / RESP_TH, n_jobs);
        DVFE_LAUNCH(k_gftt_response, grid, blk, 0, st, d_jobs);
    }
    {
        dim3 blk(32, 8), grid((max_w + 31) / 32, (max_h + 7) / 8, n_jobs);
        DVFE_LAUNCH(k_gftt_candidates, grid, blk, 0, st, d_jobs);
    }
    DVFE_LAUNCH(k_gftt_select, n_jobs, NMS_THREADS, 0, st, d_jobs);
    DVFE_LAUNCH(k_gftt_assign_ids, (n_jobs + 127) / 128, 128, 0, st, d_jobs, n_jobs);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

int launch_min_eigen_val(const uint8_t* img, int pitch, int w, int h, float* eig, cudaStream_t st) {
    dim3 blk(32, 8), grid((w + RESP_TW - 1) / RESP_TW, (h + RESP_TH - 1) / RESP_TH);
    DVFE_LAUNCH(k_min_eigen_val, grid, blk, 0, st, img, pitch, w, h, eig);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

int launch_disc_mask(uint8_t* mask, int pitch, int w, int h, const float2* pts, const int* n, int max_pts, int radius,
                     cudaStream_t st) {
    if (max_pts <= 0) return DVFE_OK;
    DVFE_LAUNCH(k_disc_mask_op, max_pts, 128, 0, st, mask, pitch, w, h, pts, n, radius);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
