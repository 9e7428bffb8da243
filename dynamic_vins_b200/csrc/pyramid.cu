// Image pyramid build — replaces cv::buildOpticalFlowPyramid as called inside
// cv::calcOpticalFlowPyrLK (reference call sites: dynamic_vins/src/front_end/feature_utils.cpp:43,50).
//   level 0 = the image; level l = pyrDown(level l-1): separable [1 4 6 4 1], BORDER_REFLECT_101,
//   (sum + 128) >> 8, size ((w+1)/2, (h+1)/2).  Integer arithmetic, bit-exact.
// Every level is written once, with a REFLECT_101 border (common.cuh), and is then reused by the
// temporal LK of this frame, the stereo LK of this frame and the temporal LK of the next frame
// (the reference rebuilds both pyramids inside each of its 4 LK calls per stereo frame).
#include "kernels.cuh"

// ---- level 0: copy the u8 image into the padded level and fill the border --------------------
__global__ void __launch_bounds__(256) k_pyr_level0(PyrImgSet set, PyrLevel L, int spitch) {
    const int img = blockIdx.z;
    const int which = img >= set.per_set;
    const int idx = which ? img - set.per_set : img;
    const uint8_t* __restrict__ src = set.src[which] + (size_t)idx * set.src_stride;
    uint8_t* __restrict__ dst = set.dst[which] + (size_t)idx * set.dst_stride + L.offset;

    const int wx = blockIdx.x * blockDim.x + threadIdx.x;      // word index in the padded row
    const int py = blockIdx.y * blockDim.y + threadIdx.y;      // padded row
    if (wx * 4 >= L.pitch || py >= L.h + 2 * DVFE_PADY) return;
    const int y = reflect101(py - DVFE_PADY, L.h);
    const int x0 = wx * 4 - DVFE_PADX;
    const uint8_t* row = src + (size_t)y * spitch;
    uint32_t v;
    if (x0 >= 0 && x0 + 3 < L.w && ((((uintptr_t)row) + x0) & 3) == 0) {
        v = __ldg(reinterpret_cast<const uint32_t*>(row + x0));
    } else {
        v = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) v |= (uint32_t)__ldg(row + reflect101(x0 + i, L.w)) << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)py * L.pitch + wx * 4) = v;
}

// ---- level l from level l-1 (both padded) ---------------------------------------------------
__device__ __forceinline__ int pyr_tap5(const uint8_t* __restrict__ p) {
    return (int)p[-2] + 4 * (int)p[-1] + 6 * (int)p[0] + 4 * (int)p[1] + (int)p[2];
}

__global__ void __launch_bounds__(256) k_pyr_down(PyrImgSet set, PyrLevel S, PyrLevel D) {
    const int img = blockIdx.z;
    const int which = img >= set.per_set;
    const int idx = which ? img - set.per_set : img;
    uint8_t* base = set.dst[which] + (size_t)idx * set.dst_stride;
    const uint8_t* __restrict__ src = base + S.offset + (size_t)DVFE_PADY * S.pitch + DVFE_PADX;   // pixel (0,0)
    uint8_t* __restrict__ dst = base + D.offset;

    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    if (wx * 4 >= D.pitch || py >= D.h + 2 * DVFE_PADY) return;
    const int y = reflect101(py - DVFE_PADY, D.h);
    const int x0 = wx * 4 - DVFE_PADX;
    uint32_t out = 0;
    if (x0 >= 0 && x0 + 3 < D.w) {
        // interior word: 4 outputs share their taps; aligned word loads (2*x0 is a multiple of 8)
        int acc[4] = {0, 0, 0, 0};
        const int kw[5] = {1, 4, 6, 4, 1};
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t* rp = reinterpret_cast<const uint32_t*>(src + (size_t)(2 * y + r - 2) * S.pitch + 2 * x0 - 4);
            const uint32_t w0 = __ldg(rp), w1 = __ldg(rp + 1), w2 = __ldg(rp + 2), w3 = __ldg(rp + 3);
            // bytes b[-4..11] ; need b[-2..8]
            int b[11];
            b[0] = (w0 >> 16) & 255; b[1] = w0 >> 24;
            b[2] = w1 & 255; b[3] = (w1 >> 8) & 255; b[4] = (w1 >> 16) & 255; b[5] = w1 >> 24;
            b[6] = w2 & 255; b[7] = (w2 >> 8) & 255; b[8] = (w2 >> 16) & 255; b[9] = w2 >> 24;
            b[10] = w3 & 255;
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc[j] += kw[r] * (b[2 * j] + 4 * b[2 * j + 1] + 6 * b[2 * j + 2] + 4 * b[2 * j + 3] + b[2 * j + 4]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) out |= (uint32_t)((acc[j] + 128) >> 8) << (8 * j);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = reflect101(x0 + i, D.w);
            const uint8_t* c = src + (size_t)(2 * y) * S.pitch + 2 * x;
            const int s = pyr_tap5(c - 2 * S.pitch) + 4 * pyr_tap5(c - S.pitch) + 6 * pyr_tap5(c) +
                          4 * pyr_tap5(c + S.pitch) + pyr_tap5(c + 2 * S.pitch);
            out |= (uint32_t)((s + 128) >> 8) << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)py * D.pitch + wx * 4) = out;
}

int launch_build_pyramids(const PyrImgSet& set, int n_img, const PyrDesc& desc, int spitch, cudaStream_t st) {
    const dim3 blk(32, 8);
    {
        const PyrLevel& L = desc.lv[0];
        dim3 grid((L.pitch / 4 + 31) / 32, (L.h + 2 * DVFE_PADY + 7) / 8, n_img);
        DVFE_LAUNCH(k_pyr_level0, grid, blk, 0, st, set, L, spitch);
    }
    for (int l = 1; l < desc.n_levels; l++) {
        const PyrLevel& D = desc.lv[l];
        dim3 grid((D.pitch / 4 + 31) / 32, (D.h + 2 * DVFE_PADY + 7) / 8, n_img);
        DVFE_LAUNCH(k_pyr_down, grid, blk, 0, st, set, desc.lv[l - 1], D);
    }
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}

// unpadded copy of one level back out (seam op dvfe_op_build_pyramid)
__global__ void k_pyr_extract(const uint8_t* base, PyrLevel L, uint8_t* out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= L.w || y >= L.h) return;
    out[(size_t)y * L.w + x] = base[L.offset + (size_t)(y + DVFE_PADY) * L.pitch + DVFE_PADX + x];
}

int launch_pyr_extract(const uint8_t* pyr, const PyrLevel& L, uint8_t* out, cudaStream_t st) {
    dim3 blk(32, 8), grid((L.w + 31) / 32, (L.h + 7) / 8);
    DVFE_LAUNCH(k_pyr_extract, grid, blk, 0, st, pyr, L, out);
    DVFE_CUDA(cudaGetLastError());
    return DVFE_OK;
}
