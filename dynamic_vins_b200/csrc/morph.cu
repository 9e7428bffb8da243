// Rectangular mask erosion — replaces ErodeMask / ErodeMaskGpu
// (dynamic_vins/src/front_end/feature_utils.h:130-146): cv::erode with a k x k MORPH_RECT element,
// anchor (-1,-1) -> k/2, i.e. window [x - k/2, x - k/2 + k - 1]; pixels outside the image never lower the
// minimum (the border is +inf).  Separable: horizontal min then vertical min.  Integer, bit-exact.
#include "kernels.cuh"

// Word-wise passes.  The masks this path sees are binary (0 / 255: inv_merge_mask, ROI masks, masks derived from a label
// image), and on binary bytes min == AND, one instruction per 4 pixels instead of one compare per pixel and tap.  Every
// thread checks the words it loads; a thread that meets a non-binary byte redoes its pixels with the byte-wise minimum, so
// the result is the exact cv::erode for any u8 input.
__device__ __forceinline__ bool word_is_binary(unsigned w) { return ((w & 0x01010101u) * 255u) == w; }
// bytes of a label-image word -> 255 where the label byte is 0, else 0   (inv_merge_mask = ~(any instance))
__device__ __forceinline__ unsigned label_word_to_inv(unsigned w) {
    const unsigned nz = ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w;       // bit 7 of each byte set iff the byte is non-zero
    return (((~nz) & 0x80808080u) >> 7) * 255u;
}
// bytes of a label-image word -> 255 where bit `bit` of the label byte is set, else 0   (full_mask of one instance)
__device__ __forceinline__ unsigned label_word_to_mask(unsigned w, int bit) { return ((w >> bit) & 0x01010101u) * 255u; }

// 4 consecutive bytes of a row starting at byte x (any alignment, x may be < 0 or reach past w): bytes outside [0, w) read as 255
// (the border of cv::erode never lowers the minimum).  `row` is 4-byte aligned or not: loads are byte-exact at the edges.
template <int MODE>      // 0: plain mask, 1: label image -> inv_merge_mask, 2: label image -> mask of bit `bit`
__device__ __forceinline__ unsigned load_mask_word(const uint8_t* __restrict__ row, int x, int w, int bit, bool aligned_row) {
    unsigned v;
    if (aligned_row && x >= 0 && x + 7 < w) {
        const unsigned* wp = reinterpret_cast<const unsigned*>(row + (x & ~3));
        v = __funnelshift_r(__ldg(wp), __ldg(wp + 1), (x & 3) * 8);
        if (MODE == 1) v = label_word_to_inv(v);
        if (MODE == 2) v = label_word_to_mask(v, bit);
        return v;
    }
    v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        unsigned b = 255u;
        if (x + i >= 0 && x + i < w) {
            b = __ldg(row + x + i);
            if (MODE == 1) b = b == 0 ? 255u : 0u;
            if (MODE == 2) b = ((b >> bit) & 1u) ? 255u : 0u;
        }
        v |= b << (8 * i);
    }
    return v;
}

// horizontal k-tap minimum of one output word (4 pixels at x .. x+3) of `row`
template <int MODE>
__device__ __forceinline__ unsigned erode_h_word(const uint8_t* __restrict__ row, int x, int w, int k, int bit, bool aligned_row) {
    const int a = k / 2;
    unsigned acc = 0xffffffffu;
    bool binary = true;
    for (int s = 0; s < k; s++) {
        const unsigned v = load_mask_word<MODE>(row, x - a + s, w, bit, aligned_row);
        if (MODE == 0) binary = binary && word_is_binary(v);
        acc &= v;
    }
    if (MODE == 0 && !binary) {          // exact byte-wise minimum for arbitrary u8 data
        acc = 0;
        for (int i = 0; i < 4; i++) {
            int m = 255;
            const int x0 = max(x + i - a, 0), x1 = min(x + i - a + k - 1, w - 1);
            for (int j = x0; j <= x1; j++) m = min(m, (int)__ldg(row + j));
            acc |= (unsigned)m << (8 * i);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(256) k_erode_h(const uint8_t* __restrict__ src, int spitch, uint8_t* __restrict__ tmp,
                                                  int w, int h, int k, size_t img_stride, size_t tmp_stride,
                                                  const int* __restrict__ enable, int label_mode) {
    const int img = blockIdx.z;
    if (enable != nullptr && !enable[img]) return;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* row = src + img * img_stride + (size_t)y * spitch;
    const bool aligned_row = (((uintptr_t)row) & 3) == 0;
    const unsigned v = label_mode ? erode_h_word<1>(row, x, w, k, 0, aligned_row) : erode_h_word<0>(row, x, w, k, 0, aligned_row);
    uint8_t* o = tmp + img * tmp_stride + (size_t)y * w + x;
    if (x + 3 < w && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<unsigned*>(o) = v;
    else for (int i = 0; i < 4 && x + i < w; i++) o[i] = (uint8_t)(v >> (8 * i));
}

// vertical k-tap minimum of one output word of the dense intermediate `tmp` (pitch w)
__device__ __forceinline__ unsigned erode_v_word(const uint8_t* __restrict__ tmp, int x, int y, int w, int h, int k) {
    const int a = k / 2;
    const int y0 = max(y - a, 0), y1 = min(y - a + k - 1, h - 1);
    const bool vec = x + 3 < w && (w & 3) == 0 && (((uintptr_t)tmp) & 3) == 0;
    unsigned acc = 0xffffffffu;
    bool binary = true;
    for (int j = y0; j <= y1; j++) {
        unsigned v;
        if (vec) v = __ldg(reinterpret_cast<const unsigned*>(tmp + (size_t)j * w + x));
        else {
            v = 0;
            for (int i = 0; i < 4; i++) v |= (x + i < w ? (unsigned)__ldg(tmp + (size_t)j * w + x + i) : 255u) << (8 * i);
        }
        binary = binary && word_is_binary(v);
        acc &= v;
    }
    if (!binary) {
        acc = 0;
        for (int i = 0; i < 4 && x + i < w; i++) {
            int m = 255;
            for (int j = y0; j <= y1; j++) m = min(m, (int)__ldg(tmp + (size_t)j * w + x + i));
            acc |= (unsigned)m << (8 * i);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(256) k_erode_v(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ dst, int dpitch,
                                                  int w, int h, int k, size_t tmp_stride, size_t dst_stride,
                                                  const int* __restrict__ enable) {
    const int img = blockIdx.z;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    uint8_t* out = dst + img * dst_stride + (size_t)y * dpitch + x;
    // no instances: mask = all 255
    const unsigned v = (enable != nullptr && !enable[img]) ? 0xffffffffu : erode_v_word(tmp + img * tmp_stride, x, y, w, h, k);
    if (x + 3 < w && (((uintptr_t)out) & 3) == 0) *reinterpret_cast<unsigned*>(out) = v;
    else for (int i = 0; i < 4 && x + i < w; i++) out[i] = (uint8_t)(v >> (8 * i));
}

// src: n_img images (stride img_stride, pitch spitch) -> dst (dense pitch dpitch, stride dpitch*h);
// tmp: n_img * w * h scratch.  enable[img] == 0 -> dst image is filled with 255 instead.
int launch_erode_rect(const uint8_t* src, int spitch, uint8_t* dst, int dpitch, uint8_t* tmp, int w, int h, int k,
                      int n_img, size_t img_stride, const int* enable, cudaStream_t st, int label_mode) {
    dim3 blk(32, 8), grid(((w + 3) / 4 + 31) / 32, (h + 7) / 8, n_img);
    DVFE_LAUNCH(k_erode_h, grid, blk, 0, st, src, spitch, tmp, w, h, k, img_stride, (size_t)w * h, enable, label_mode);
    DVFE_LAUNCH(k_erode_v, grid, blk, 0, st, tmp, dst, dpitch, w, h, k, (size_t)w * h, (size_t)dpitch * h, enable);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

// job-based variant: one image per job, individual sizes (instance ROI masks)
__global__ void __launch_bounds__(256) k_erode_h_jobs(const ErodeJob* __restrict__ jobs) {
    const ErodeJob& J = jobs[blockIdx.z];
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= J.w || y >= J.h) return;
    const uint8_t* row = J.src + (size_t)y * J.spitch;
    const bool aligned_row = (((uintptr_t)row) & 3) == 0;
    // label_bit >= 0: full_mask(rect) of instance `label_bit` (basic/semantic_image.cpp:48-56), read from the label image
    const unsigned v = J.label_bit >= 0 ? erode_h_word<2>(row, x, J.w, J.k, J.label_bit, aligned_row)
                                        : erode_h_word<0>(row, x, J.w, J.k, 0, aligned_row);
    uint8_t* o = J.tmp + (size_t)y * J.w + x;
    if (x + 3 < J.w && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<unsigned*>(o) = v;
    else for (int i = 0; i < 4 && x + i < J.w; i++) o[i] = (uint8_t)(v >> (8 * i));
}

__global__ void __launch_bounds__(256) k_erode_v_jobs(const ErodeJob* __restrict__ jobs) {
    const ErodeJob& J = jobs[blockIdx.z];
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= J.w || y >= J.h) return;
    const unsigned v = erode_v_word(J.tmp, x, y, J.w, J.h, J.k);
    uint8_t* o = J.dst + (size_t)y * J.w + x;
    if (x + 3 < J.w && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<unsigned*>(o) = v;
    else for (int i = 0; i < 4 && x + i < J.w; i++) o[i] = (uint8_t)(v >> (8 * i));
}

int launch_erode_jobs(const ErodeJob* d_jobs, int n_jobs, int max_w, int max_h, cudaStream_t st) {
    if (n_jobs <= 0) return DVFE_OK;
    dim3 blk(32, 8), grid(((max_w + 3) / 4 + 31) / 32, (max_h + 7) / 8, n_jobs);
    DVFE_LAUNCH(k_erode_h_jobs, grid, blk, 0, st, d_jobs);
    DVFE_LAUNCH(k_erode_v_jobs, grid, blk, 0, st, d_jobs);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
