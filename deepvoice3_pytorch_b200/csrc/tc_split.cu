// Operand preparation for the tensor-core path (tc_gemm.cu): every fp32 operand x becomes bf16 planes
// p0 = bf16(x), p1 = bf16(x - p0) [, p2 = bf16(x - p0 - p1)], written in the K-major layout each implicit GEMM
// consumes.  These passes are pure HBM streaming (read 4 B, write 4-12 B per element) and absorb work the fp32 path
// does inside its loaders: the conv-input dropout mask, the ReLU mask of the incoming gradient, the bias-gradient
// reduction and the (B,C,T) <-> (B,T,C) layout change.  Channel pitches are padded to a multiple of 8 (16 bytes,
// the TMA stride granularity); pad columns are never read (the tensor maps carry the true extent).
#include <cuda_bf16.h>
#include "common.cuh"
#include "wn_device.cuh"

namespace dv3 {

// x (B,C,T) fp32 -> conv-input dropout -> (hi, lo) fp16 planes in (B,T,Cp): the forward GEMM's K-major operand and
// the weight gradient's MN-major operand.  32(c) x 32(t) tile per CTA, block (32, 8).
__global__ void split_input_kernel(const float* __restrict__ x, bf16* __restrict__ btc, int Bn, int C, int Cp, int T,
                                   float p, const unsigned long long* __restrict__ seed_ptr, unsigned salt) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float tile[32][33];
    const DropCfg drop = make_drop(p, seed_ptr, salt);
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + threadIdx.y + 8 * i, t = t0 + threadIdx.x;
        float v = 0.f;
        if (c < C && t < T) {
            const size_t row = ((size_t)b * C + c) * T;
            v = x[row + t] * drop_scale(drop, (uint32_t)(row + t));
        }
        tile[threadIdx.y + 8 * i][threadIdx.x] = v;
    }
    __syncthreads();
    const size_t btc_plane = (size_t)Bn * T * Cp;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + threadIdx.y + 8 * i, c = c0 + threadIdx.x;
        if (c < C && t < T)
            split_store<FMT_F16>(tile[threadIdx.x][threadIdx.y + 8 * i], btc, ((size_t)b * T + t) * Cp + c, btc_plane);
    }
}

// gate backward (see conv.cu gate_bwd_kernel) producing dAB = [da ; db] * GRAD_SCALE directly as 2 fp16 planes in
// (B,T,2C) [data-gradient operand] and (B,2C,T) [weight-gradient operand]; dbias[2C] += sums over (b,t).
__global__ void gate_bwd_split_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                      const float* __restrict__ s, const float* __restrict__ x,
                                      bf16* __restrict__ btc, float* __restrict__ dbias, int Bn, int C, int T, int mode,
                                      int residual) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float ta[32][33], tb[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const float gs = (mode == 0 && residual) ? 0.70710678118654752f : 1.f;
    const size_t plane = (size_t)Bn * 2 * C * T;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + threadIdx.y + 8 * i, t = t0 + threadIdx.x;
        float da = 0.f, db = 0.f;
        if (c < C && t < T) {
            const size_t in = ((size_t)b * C + c) * T + t;
            const float g = dy[in] * gs, av = a[in], sv = s[in];
            da = g * sv;
            db = g * ((mode == 0) ? av : (av - x[in])) * sv * (1.f - sv);
        }
        ta[threadIdx.y + 8 * i][threadIdx.x] = da;
        tb[threadIdx.y + 8 * i][threadIdx.x] = db;
        // bias gradient: reduce this row's 32 time steps across the warp (threadIdx.x = lane)
        const float sa = warp_sum(da), sb = warp_sum(db);
        if (threadIdx.x == 0 && c < C && dbias) { atomicAdd(&dbias[c], sa); atomicAdd(&dbias[C + c], sb); }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + threadIdx.y + 8 * i, c = c0 + threadIdx.x;
        if (c < C && t < T) {
            const size_t o = ((size_t)b * T + t) * 2 * C + c;
            split_store<FMT_F16>(ta[threadIdx.x][threadIdx.y + 8 * i] * GRAD_SCALE, btc, o, plane);
            split_store<FMT_F16>(tb[threadIdx.x][threadIdx.y + 8 * i] * GRAD_SCALE, btc, o + C, plane);
        }
    }
}

// plain conv backward prologue: g = dy * (relu ? y > 0 : 1) -> 2 planes in (B,T,Cp) and (B,C,T); dbias[C] += sums.
__global__ void grad_split_kernel(const float* __restrict__ dy, const float* __restrict__ y, bf16* __restrict__ btc,
                                  float* __restrict__ dbias, int Bn, int C, int Cp, int T, int relu) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + threadIdx.y + 8 * i, t = t0 + threadIdx.x;
        float g = 0.f;
        if (c < C && t < T) {
            const size_t in = ((size_t)b * C + c) * T + t;
            g = dy[in];
            if (relu && !(y[in] > 0.f)) g = 0.f;
        }
        tile[threadIdx.y + 8 * i][threadIdx.x] = g;
        const float sg = warp_sum(g);
        if (threadIdx.x == 0 && c < C && dbias) atomicAdd(&dbias[c], sg);
    }
    __syncthreads();
    if (btc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 + threadIdx.y + 8 * i, c = c0 + threadIdx.x;
            if (c < C && t < T)
                split_store<FMT_F16>(tile[threadIdx.x][threadIdx.y + 8 * i] * GRAD_SCALE, btc,
                                     ((size_t)b * T + t) * Cp + c, (size_t)Bn * T * Cp);
        }
    }
}

// weight-norm pack: v [R][X][k] fp32, scale[R] = g/||v|| -> two plane sets with element (r,x,j) at
// r*s_r + x*s_x + j*s_j: outA is written with lanes along (x,j) (choose the set whose unit stride is s_x),
// outB with lanes along r (unit stride s_r).
template <int FMTA, int FMTB>
__global__ void wn_pack_split_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                     bf16* __restrict__ outA, long long a_r, long long a_x, long long a_j,
                                     long long a_plane, bf16* __restrict__ outB, long long b_r, long long b_x,
                                     long long b_j, long long b_plane, int R, int X, int k) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float tile[32][33];
    wn_pack_split_tile<FMTA, FMTB>(v, scale, outA, a_r, a_x, a_j, a_plane, outB, b_r, b_x, b_j, b_plane, R, X, k,
                                   blockIdx.x, blockIdx.y, tile);
}

__global__ void wn_norm_kernel2(const float* __restrict__ v, const float* __restrict__ g,
                                float* __restrict__ inv_norm, float* __restrict__ scale, int R, int L) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    wn_norm_row(v, g, inv_norm, scale, R, L, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31);
}

}  // namespace dv3

using namespace dv3;

extern "C" {

int dv3_tc_split_input(const float* x, void* btc, int npl, void* bct, int B, int C, int T, int k, int dilation,
                       int causal, float p_drop, const unsigned long long* seed_ptr, unsigned salt, void* stream) {
    DV3_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "tc_split_input: grid too large");
    DV3_REQUIRE(npl == 2 && bct == nullptr, "tc_split_input: npl must be 2 and bct NULL (the weight gradient reads btc)");
    (void)k; (void)dilation; (void)causal;
    const int Cp = (C + 7) / 8 * 8;
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    launch_k(split_input_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, x, (bf16*)btc, B, C, Cp, T, p_drop, seed_ptr,
             salt);
    return check_launch("tc_split_input");
}

int dv3_tc_gate_bwd_split(const float* dy, const float* a, const float* s, const float* x, void* btc, void* bct,
                          float* dbias, int B, int C, int T, int mode, int residual, void* stream) {
    DV3_REQUIRE(bct == nullptr, "tc_gate_bwd_split: bct must be NULL");
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    launch_k(gate_bwd_split_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, dy, a, s, x, (bf16*)btc, dbias, B, C, T,
             mode, residual);
    return check_launch("tc_gate_bwd_split");
}

int dv3_tc_grad_split(const float* dy, const float* y, void* btc, void* bct, float* dbias, int B, int C, int T,
                      int relu, void* stream) {
    DV3_REQUIRE(bct == nullptr, "tc_grad_split: bct must be NULL");
    const int Cp = (C + 7) / 8 * 8;
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    launch_k(grad_split_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, dy, y, (bf16*)btc, dbias, B, C, Cp, T, relu);
    return check_launch("tc_grad_split");
}

// Weight norm + split for a conv weight v (Cout, Cin, k), g [Cout]:
//   wfwd: [2][k][Cout][Cinp] fp16 planes (forward operand: rows co, K = ci)
//   wbwd: [2][k][Cin][Coutp] fp16 planes (data-gradient operand)
int dv3_tc_weightnorm_fwd(const float* v, const float* g, float* inv_norm, float* scale, void* wfwd, int npl,
                          void* wbwd, int Cout, int Cin, int k, void* stream) {
    DV3_REQUIRE(npl == 2, "tc_weightnorm_fwd: npl must be 2");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = Cin * k;
    const long long Cinp = (Cin + 7) / 8 * 8, Coutp = (Cout + 7) / 8 * 8;
    launch_k(wn_norm_kernel2, (Cout * 32 + 255) / 256, 256, 0, st, v, g, inv_norm, scale, Cout, L);
    if (int e = check_launch("tc_weightnorm_fwd(norm)")) return e;
    dim3 grid((L + 31) / 32, (Cout + 31) / 32);
    launch_k(wn_pack_split_kernel<FMT_F16, FMT_F16>, grid, dim3(32, 8), 0, st, v, scale, (bf16*)wfwd, Cinp, 1,
             (long long)Cout * Cinp, (long long)k * Cout * Cinp, (bf16*)wbwd, 1, Coutp, (long long)Cin * Coutp,
             (long long)k * Cin * Coutp, Cout, Cin, k);
    return check_launch("tc_weightnorm_fwd(pack)");
}

// ConvTranspose1d(k=2,s=2) weight v (Cin, Cout, 2), g [Cin] (norm over dim 0 = Cin), run as a 1x1 conv with
// 2*Cout output rows ordered (j, co):
//   wfwd: [2][2*Cout][Cinp] fp16, rows (j,co), K = ci        wbwd: [2][Cin][K2p] fp16, rows ci, K = (j,co)
int dv3_tc_weightnorm_convt_fwd(const float* v, const float* g, float* inv_norm, float* scale, void* wfwd, int npl,
                                void* wbwd, int Cin, int Cout, void* stream) {
    DV3_REQUIRE(npl == 2, "tc_weightnorm_convt_fwd: npl must be 2");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = Cout * 2;
    const long long Cinp = (Cin + 7) / 8 * 8, K2p = (2 * Cout + 7) / 8 * 8;
    launch_k(wn_norm_kernel2, (Cin * 32 + 255) / 256, 256, 0, st, v, g, inv_norm, scale, Cin, L);
    if (int e = check_launch("tc_weightnorm_convt_fwd(norm)")) return e;
    dim3 grid((L + 31) / 32, (Cin + 31) / 32);
    // r = ci, x = co, j: outA (lanes along (x,j)) = wbwd [ci][j*Cout+co] ; outB (lanes along r) = wfwd [(j*Cout+co)][ci]
    launch_k(wn_pack_split_kernel<FMT_F16, FMT_F16>, grid, dim3(32, 8), 0, st, v, scale, (bf16*)wbwd, K2p, 1,
             (long long)Cout, (long long)Cin * K2p, (bf16*)wfwd, 1, Cinp, (long long)Cout * Cinp,
             (long long)2 * Cout * Cinp, Cin, Cout, 2);
    return check_launch("tc_weightnorm_convt_fwd(pack)");
}

}  // extern "C"
