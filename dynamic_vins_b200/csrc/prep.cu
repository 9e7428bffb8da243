// Upstream frame preparation — the two byte-wise passes that bracket the hot path today:
//   SemanticImage::SetGrayImage[Gpu]   cv::cvtColor(color, gray, CV_BGR2GRAY)      basic/semantic_image.cpp:69-93
//   SemanticImage::SetMaskAndRoi / SetBackgroundMask: merge_mask = any(instance mask) * 255,
//                                       inv_merge_mask = bitwise_not(merge_mask)   basic/semantic_image.cpp:20-63,103-117
// Pure HBM-bandwidth kernels, 4 pixels per thread.  cvtColor arithmetic: the 15-bit fixed-point coefficients of the
// cv2 4.13 oracle, gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15 (OpenCV 3.4's scalar path uses the 14-bit set
// 1868/9617/4899, which differs by one grey level on ~0.2 % of pixels).
#include "kernels.cuh"

__global__ void __launch_bounds__(256) k_bgr_to_gray(const uint8_t* __restrict__ bgr, int spitch, uint8_t* __restrict__ gray,
                                                     int dpitch, int w, int h) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= w || y >= h) return;
    const uint8_t* s = bgr + (size_t)y * spitch + 3 * x0;
    uint8_t* d = gray + (size_t)y * dpitch + x0;
    unsigned out = 0;
    const int n = min(4, w - x0);
    if (n == 4 && ((uintptr_t)s & 3) == 0) {
        const unsigned* p = reinterpret_cast<const unsigned*>(s);        // 12 bytes: B G R B | G R B G | R B G R
        const unsigned w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const unsigned px[4][3] = {{w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255},
                                   {w0 >> 24, w1 & 255, (w1 >> 8) & 255},
                                   {(w1 >> 16) & 255, w1 >> 24, w2 & 255},
                                   {(w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24}};
#pragma unroll
        for (int i = 0; i < 4; i++) out |= ((px[i][0] * 3735u + px[i][1] * 19235u + px[i][2] * 9798u + 16384u) >> 15) << (8 * i);
    } else {
        for (int i = 0; i < n; i++)
            out |= (((unsigned)s[3 * i] * 3735u + (unsigned)s[3 * i + 1] * 19235u + (unsigned)s[3 * i + 2] * 9798u + 16384u) >> 15) << (8 * i);
    }
    if (n == 4 && ((uintptr_t)d & 3) == 0) *reinterpret_cast<unsigned*>(d) = out;
    else for (int i = 0; i < n; i++) d[i] = (uint8_t)(out >> (8 * i));
}

__global__ void __launch_bounds__(256) k_merge_masks(const uint8_t* __restrict__ masks, int n_masks, size_t mask_stride,
                                                     int spitch, uint8_t* __restrict__ merge, uint8_t* __restrict__ inv,
                                                     int dpitch, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    unsigned any = 0;
    for (int i = 0; i < n_masks; i++) any |= masks[i * mask_stride + (size_t)y * spitch + x];
    const uint8_t m = any ? 255 : 0;
    merge[(size_t)y * dpitch + x] = m;
    inv[(size_t)y * dpitch + x] = (uint8_t)~m;
}

extern "C" int dvfe_op_bgr_to_gray(const uint8_t* bgr, int w, int h, int pitch, uint8_t* gray_out) {
    if (!bgr || !gray_out || w < 1 || h < 1 || pitch < 3 * w) { dvfe_set_error("op_bgr_to_gray: bad argument"); return DVFE_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { dvfe_set_error("no CUDA device available: libdvfe has no CPU fallback"); return DVFE_ERR_NO_DEVICE; }
    uint8_t *d_bgr = nullptr, *d_gray = nullptr;
    DVFE_CUDA(cudaMalloc((void**)&d_bgr, (size_t)3 * w * h));
    DVFE_CUDA(cudaMalloc((void**)&d_gray, (size_t)w * h));
    DVFE_CUDA(cudaMemcpy2D(d_bgr, (size_t)3 * w, bgr, pitch, (size_t)3 * w, h, cudaMemcpyHostToDevice));
    dim3 blk(32, 8), grid(((w + 3) / 4 + 31) / 32, (h + 7) / 8);
    DVFE_LAUNCH(k_bgr_to_gray, grid, blk, 0, 0, d_bgr, 3 * w, d_gray, w, w, h);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(gray_out, d_gray, (size_t)w * h, cudaMemcpyDeviceToHost);
    cudaFree(d_bgr); cudaFree(d_gray);
    if (e != cudaSuccess) { dvfe_set_error("op_bgr_to_gray: %s", cudaGetErrorString(e)); return DVFE_ERR_CUDA; }
    return DVFE_OK;
}

extern "C" int dvfe_op_merge_masks(const uint8_t* masks, int n_masks, int w, int h, uint8_t* merge_out, uint8_t* inv_out) {
    if (n_masks < 0 || (n_masks > 0 && !masks) || !merge_out || !inv_out || w < 1 || h < 1) {
        dvfe_set_error("op_merge_masks: bad argument");
        return DVFE_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { dvfe_set_error("no CUDA device available: libdvfe has no CPU fallback"); return DVFE_ERR_NO_DEVICE; }
    const size_t P = (size_t)w * h;
    uint8_t *d_m = nullptr, *d_merge = nullptr, *d_inv = nullptr;
    DVFE_CUDA(cudaMalloc((void**)&d_m, P * (n_masks > 0 ? n_masks : 1)));
    DVFE_CUDA(cudaMalloc((void**)&d_merge, P));
    DVFE_CUDA(cudaMalloc((void**)&d_inv, P));
    if (n_masks > 0) DVFE_CUDA(cudaMemcpy(d_m, masks, P * n_masks, cudaMemcpyHostToDevice));
    dim3 blk(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    DVFE_LAUNCH(k_merge_masks, grid, blk, 0, 0, d_m, n_masks, P, w, d_merge, d_inv, w, w, h);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(merge_out, d_merge, P, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(inv_out, d_inv, P, cudaMemcpyDeviceToHost);
    cudaFree(d_m); cudaFree(d_merge); cudaFree(d_inv);
    if (e != cudaSuccess) { dvfe_set_error("op_merge_masks: %s", cudaGetErrorString(e)); return DVFE_ERR_CUDA; }
    return DVFE_OK;
}
