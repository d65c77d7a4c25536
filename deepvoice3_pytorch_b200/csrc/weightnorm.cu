// Weight normalisation  w = g * v / ||v||  (reference modules.py:85,100,109 -- old-style
// torch.nn.utils.weight_norm, dim=0, re-evaluated by a forward-pre-hook on every call).
// The forward writes w directly in the two packed layouts the conv kernels consume
// (forward operand and data-gradient operand) so no separate transpose/pack pass exists;
// the backward folds the split-K reduction of the weight-gradient partials into the g/v gradient.
#include "common.cuh"
#include "wn_device.cuh"

namespace dv3 {

// one warp per row r: inv_norm[r] = 1/||v[r,:]||, scale[r] = g[r]*inv_norm[r]
__global__ void wn_norm_kernel(const float* __restrict__ v, const float* __restrict__ g,
                               float* __restrict__ inv_norm, float* __restrict__ scale, int R, int L) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* row = v + (size_t)r * L;
    float s = 0.f;
    for (int e = lane; e < L; e += 32) { const float x = row[e]; s = fmaf(x, x, s); }
    s = warp_sum(s);
    if (lane == 0) {
        const float inv = 1.f / sqrtf(s);
        inv_norm[r] = inv;
        scale[r] = g[r] * inv;
    }
}

// 32 rows x 32 elements per CTA (block 32x8); element e of a row = (x, j) with e = x*k + j.
// out1 is written with lanes along r, out2 with lanes along e; either may be null.
__global__ void wn_pack_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                               float* __restrict__ out1, float* __restrict__ out2, int R, int L, int k,
                               long long s1r, long long s1x, long long s1j, long long s2r, long long s2x,
                               long long s2j) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, e0 = blockIdx.x * 32;
    const int lx = threadIdx.x, ly = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ly + 8 * i, e = e0 + lx;
        float w = 0.f;
        if (r < R && e < L) {
            w = v[(size_t)r * L + e] * scale[r];
            if (out2) { const int x = e / k, j = e - x * k; out2[r * s2r + x * s2x + j * s2j] = w; }
        }
        tile[ly + 8 * i][lx] = w;
    }
    __syncthreads();
    if (out1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = e0 + ly + 8 * i, r = r0 + lx;
            if (r < R && e < L) {
                const int x = e / k, j = e - x * k;
                out1[r * s1r + x * s1x + j * s1j] = tile[lx][ly + 8 * i];
            }
        }
    }
}

// backward, one CTA (256 threads) per row (body in wn_device.cuh)
__global__ void __launch_bounds__(256) wn_bwd_kernel(float* __restrict__ dw_partials, long long split_stride,
                                                     int nsplit, int jmajor_X, const float* __restrict__ v,
                                                     const float* __restrict__ g,
                                                     const float* __restrict__ inv_norm, float* __restrict__ dv,
                                                     float* __restrict__ dg, int R, int L, int accumulate) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    wn_bwd_row(dw_partials, split_stride, nsplit, jmajor_X, v, g, inv_norm, dv, dg, R, L, accumulate, blockIdx.x);
}

}  // namespace dv3

using namespace dv3;

extern "C" {

// v: [R][X][k] (k fastest).  out1[r*s1r + x*s1x + j*s1j], out2[...]; inv_norm/scale: [R] workspaces.
int dv3_weightnorm_fwd(const float* v, const float* g, float* inv_norm, float* scale, float* out1,
                       float* out2, int R, int X, int k, long long s1r, long long s1x, long long s1j,
                       long long s2r, long long s2x, long long s2j, void* stream) {
    DV3_REQUIRE(R > 0 && X > 0 && k > 0, "weightnorm_fwd: empty weight");
    const int L = X * k;
    cudaStream_t st = (cudaStream_t)stream;
    launch_k(wn_norm_kernel, ceil_div(R * 32, 256), 256, 0, st, v, g, inv_norm, scale, R, L);
    if (int e = check_launch("weightnorm_fwd(norm)")) return e;
    launch_k(wn_pack_kernel, dim3(ceil_div(L, 32), ceil_div(R, 32)), dim3(32, 8), 0, st, 
        v, scale, out1, out2, R, L, k, s1r, s1x, s1j, s2r, s2x, s2j);
    return check_launch("weightnorm_fwd(pack)");
}

// dw_partials: [nsplit][R*X*k] (slot 0 is overwritten with the reduced dW); tap_major = 0: v's own layout
// (r, x, j); 1: [j][r][x].  accumulate = 1 adds into dv / dg instead of overwriting them.
int dv3_weightnorm_bwd(float* dw_partials, long long split_stride, int nsplit, int tap_major, const float* v,
                       const float* g, const float* inv_norm, float* dv, float* dg, int R, int X, int k,
                       int accumulate, void* stream) {
    launch_k(wn_bwd_kernel, R, 256, 0, (cudaStream_t)stream, 
        dw_partials, split_stride, nsplit, tap_major ? X : 0, v, g, inv_norm, dv, dg, R, X * k, accumulate);
    return check_launch("weightnorm_bwd");
}

}  // extern "C"
