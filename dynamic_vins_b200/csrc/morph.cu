// Rectangular mask erosion — replaces ErodeMask / ErodeMaskGpu
// (dynamic_vins/src/front_end/feature_utils.h:130-146): cv::erode with a k x k MORPH_RECT element,
// anchor (-1,-1) -> k/2, i.e. window [x - k/2, x - k/2 + k - 1]; pixels outside the image never lower the
// minimum (the border is +inf).  Separable: horizontal min then vertical min.  Integer, bit-exact.
#include "kernels.cuh"

__global__ void __launch_bounds__(256) k_erode_h(const uint8_t* __restrict__ src, int spitch, uint8_t* __restrict__ tmp,
                                                  int w, int h, int k, size_t img_stride, size_t tmp_stride,
                                                  const int* __restrict__ enable, int label_mode) {
    const int img = blockIdx.z;
    if (enable != nullptr && !enable[img]) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* row = src + img * img_stride + (size_t)y * spitch;
    const int a = k / 2;
    const int x0 = max(x - a, 0), x1 = min(x - a + k - 1, w - 1);
    int m = 255;
    if (label_mode) {      // SemanticImage::SetMaskAndRoi: inv_merge_mask = ~(any instance) (basic/semantic_image.cpp:31-38)
        for (int i = x0; i <= x1; i++) m = min(m, __ldg(row + i) == 0 ? 255 : 0);
    } else {
        for (int i = x0; i <= x1; i++) m = min(m, (int)__ldg(row + i));
    }
    tmp[img * tmp_stride + (size_t)y * w + x] = (uint8_t)m;
}

__global__ void __launch_bounds__(256) k_erode_v(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ dst, int dpitch,
                                                  int w, int h, int k, size_t tmp_stride, size_t dst_stride,
                                                  const int* __restrict__ enable) {
    const int img = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    uint8_t* out = dst + img * dst_stride + (size_t)y * dpitch + x;
    if (enable != nullptr && !enable[img]) { *out = 255; return; }   // no instances: mask = all 255
    const uint8_t* col = tmp + img * tmp_stride + x;
    const int a = k / 2;
    const int y0 = max(y - a, 0), y1 = min(y - a + k - 1, h - 1);
    int m = 255;
    for (int j = y0; j <= y1; j++) m = min(m, (int)__ldg(col + (size_t)j * w));
    *out = (uint8_t)m;
}

// src: n_img images (stride img_stride, pitch spitch) -> dst (dense pitch dpitch, stride dpitch*h);
// tmp: n_img * w * h scratch.  enable[img] == 0 -> dst image is filled with 255 instead.
int launch_erode_rect(const uint8_t* src, int spitch, uint8_t* dst, int dpitch, uint8_t* tmp, int w, int h, int k,
                      int n_img, size_t img_stride, const int* enable, cudaStream_t st, int label_mode) {
    dim3 blk(32, 8), grid((w + 31) / 32, (h + 7) / 8, n_img);
    DVFE_LAUNCH(k_erode_h, grid, blk, 0, st, src, spitch, tmp, w, h, k, img_stride, (size_t)w * h, enable, label_mode);
    DVFE_LAUNCH(k_erode_v, grid, blk, 0, st, tmp, dst, dpitch, w, h, k, (size_t)w * h, (size_t)dpitch * h, enable);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

// job-based variant: one image per job, individual sizes (instance ROI masks)
__global__ void __launch_bounds__(256) k_erode_h_jobs(const ErodeJob* __restrict__ jobs) {
    const ErodeJob& J = jobs[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= J.w || y >= J.h) return;
    const uint8_t* row = J.src + (size_t)y * J.spitch;
    const int a = J.k / 2;
    const int x0 = max(x - a, 0), x1 = min(x - a + J.k - 1, J.w - 1);
    int m = 255;
    if (J.label_bit >= 0) {      // full_mask(rect) of instance `label_bit` (basic/semantic_image.cpp:48-56), read from the label image
        for (int i = x0; i <= x1; i++) m = min(m, ((row[i] >> J.label_bit) & 1) ? 255 : 0);
    } else {
        for (int i = x0; i <= x1; i++) m = min(m, (int)row[i]);
    }
    J.tmp[(size_t)y * J.w + x] = (uint8_t)m;
}

__global__ void __launch_bounds__(256) k_erode_v_jobs(const ErodeJob* __restrict__ jobs) {
    const ErodeJob& J = jobs[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= J.w || y >= J.h) return;
    const uint8_t* col = J.tmp + x;
    const int a = J.k / 2;
    const int y0 = max(y - a, 0), y1 = min(y - a + J.k - 1, J.h - 1);
    int m = 255;
    for (int j = y0; j <= y1; j++) m = min(m, (int)col[(size_t)j * J.w]);
    J.dst[(size_t)y * J.w + x] = (uint8_t)m;
}

int launch_erode_jobs(const ErodeJob* d_jobs, int n_jobs, int max_w, int max_h, cudaStream_t st) {
    if (n_jobs <= 0) return DVFE_OK;
    dim3 blk(32, 8), grid((max_w + 31) / 32, (max_h + 7) / 8, n_jobs);
    DVFE_LAUNCH(k_erode_h_jobs, grid, blk, 0, st, d_jobs);
    DVFE_LAUNCH(k_erode_v_jobs, grid, blk, 0, st, d_jobs);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
