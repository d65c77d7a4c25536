// Operand preparation for the tensor-core path (tc_gemm.cu): every fp32 operand x becomes a 16-bit (hi, lo * 2^11) pair
// (common.cuh: fp16 pairs for forward operands, bf16 pairs for gradients), written in the (B,T,C) layout each implicit
// GEMM consumes (K-major for the forward / data gradient, MN-major for the weight gradient).  These passes are pure HBM
// streaming and absorb work the fp32 path
// does inside its loaders: the conv-input dropout mask, the ReLU mask of the incoming gradient, the bias-gradient
// reduction and the (B,C,T) <-> (B,T,C) layout change.  Channel pitches are padded to a multiple of 8 (16 bytes,
// the TMA stride granularity); pad columns are never read (the tensor maps carry the true extent).
#include <cuda_bf16.h>
#include "common.cuh"
#include "wn_device.cuh"

namespace dv3 {

// ---- activation-side operand preparation: ONE kernel template, three uses ------------------------------------
//   SPLIT_INPUT  x (B,C,T) -> conv-input dropout -> fp16 planes (B,T,Cp) [forward operand] and, when wg != NULL, the same
//                values as a bf16 pair [operand of the weight gradient, which multiplies them with bf16 gradient
//                planes: tcgen05 kind::f16 cannot mix fp16 and bf16 operands in one MMA]
//   SPLIT_GATE   gate backward (conv.cu gate_bwd_kernel) of dy with the saved a, s (, x) -> dAB = [da | db] as bf16
//                planes (B,T,2C); dbias[2C] += sums over (b,t)                           [data-/weight-gradient operand]
//   SPLIT_GRAD   g = dy * (relu ? y > 0 : 1) -> bf16 planes (B,T,Cp); dbias[C] += sums
// A CTA (256 threads) owns 64 channels x 64 time steps of one utterance: float4 loads along T (16 threads per channel
// row, 3-4 tensors in flight per thread), bias-gradient sums reduced across those 16 lanes (one atomic per row and
// tile), transpose through shared memory, and 16-byte stores of 8 channels per thread -- 8 threads cover the 128
// contiguous bytes of one (b,t) row of a plane.  (The round-1 kernels used 32x32 tiles with 2-byte stores, 64 B per
// warp-store row: 25 % of the HBM roof over a training step.)
enum { SPLIT_INPUT = 0, SPLIT_GATE = 1, SPLIT_GRAD = 2 };

struct SplitParams {
    const float* in0;          // x | dy | dy
    const float* in1;          // - | a  | y (relu) or null
    const float* in2;          // - | s  | -
    const float* in3;          // - | x (highway) | -
    bf16* planes;              // [2][B][T][pitch]
    bf16* wg;                  // SPLIT_INPUT: bf16 copy [2][B][T][pitch] or null
    float* dbias;              // null | [2C] | [C]
    int B, C, T, pitch;
    int mode, residual, relu;  // gate mode (0 GLU, 1 highway), GLU residual flag; ReLU flag
    float p; const unsigned long long* seed_ptr; unsigned salt;     // SPLIT_INPUT dropout
};

template <int KIND>
__global__ void __launch_bounds__(256) plane_split_kernel(const __grid_constant__ SplitParams p) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float sa[64][65];
    __shared__ float sb[KIND == SPLIT_GATE ? 64 : 1][65];
    const int tid = threadIdx.x;
    const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
    const int C = p.C, T = p.T;
    const bool vec = (T & 3) == 0;
    const DropCfg drop = make_drop(KIND == SPLIT_INPUT ? p.p : 0.f, p.seed_ptr, p.salt);
    const float gs = (KIND == SPLIT_GATE && p.mode == 0 && p.residual) ? 0.70710678118654752f : 1.f;
    const int q = tid & 15;                                  // float4 slot along T
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = (tid >> 4) + 16 * j, c = c0 + r, t = t0 + 4 * q;
        float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < C && t < T) {
            const size_t row = ((size_t)b * C + c) * T + t;
            float x0[4], x1[4], x2[4], x3[4];
            const bool full = vec && t + 3 < T;
            auto ld = [&](const float* src, float* dst) {
                if (full) {
                    const float4 u = __ldg(reinterpret_cast<const float4*>(src + row));
                    dst[0] = u.x; dst[1] = u.y; dst[2] = u.z; dst[3] = u.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) dst[e] = (t + e < T) ? __ldg(src + row + e) : 0.f;
                }
            };
            ld(p.in0, x0);
            if (KIND == SPLIT_GATE) { ld(p.in1, x1); ld(p.in2, x2); if (p.mode != 0) ld(p.in3, x3); }
            if (KIND == SPLIT_GRAD && p.relu) ld(p.in1, x1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (KIND == SPLIT_INPUT) {
                    v0[e] = x0[e] * drop_scale(drop, (uint32_t)(row + e));
                } else if (KIND == SPLIT_GATE) {
                    const float g = x0[e] * gs, av = x1[e], sv = x2[e];
                    v0[e] = g * sv;
                    v1[e] = g * (p.mode == 0 ? av : (av - x3[e])) * sv * (1.f - sv);
                } else {
                    v0[e] = (p.relu && !(x1[e] > 0.f)) ? 0.f : x0[e];
                }
                if (t + e >= T) { v0[e] = 0.f; v1[e] = 0.f; }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sa[r][4 * q + e] = v0[e];
            if (KIND == SPLIT_GATE) sb[r][4 * q + e] = v1[e];
        }
        if (KIND != SPLIT_INPUT && p.dbias) {                // all 32 lanes take part: 16 lanes share a channel row
            float s0 = v0[0] + v0[1] + v0[2] + v0[3], s1 = v1[0] + v1[1] + v1[2] + v1[3];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                if (KIND == SPLIT_GATE) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            if (q == 0 && c < C) {
                atomicAdd(&p.dbias[c], s0);
                if (KIND == SPLIT_GATE) atomicAdd(&p.dbias[C + c], s1);
            }
        }
    }
    __syncthreads();
    constexpr int FMT = KIND == SPLIT_INPUT ? FMT_F16 : FMT_BF16;
    const size_t plane = (size_t)p.B * T * p.pitch;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int slot = tid + 256 * j, tt = slot >> 3, cg = slot & 7, t = t0 + tt, c = c0 + cg * 8;
        if (t >= T || c >= (KIND == SPLIT_GATE ? C : p.pitch)) continue;
        const size_t off = ((size_t)b * T + t) * p.pitch + c;
#pragma unroll
        for (int half = 0; half < (KIND == SPLIT_GATE ? 2 : 1); ++half) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = half ? sb[(cg * 8 + i) % 64][tt] : sa[cg * 8 + i][tt];
            uint32_t h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint16_t h0, l0, h1, l1;
                split_pair<FMT>(e[2 * i], h0, l0);
                split_pair<FMT>(e[2 * i + 1], h1, l1);
                h[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                l[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
            uint16_t* d = reinterpret_cast<uint16_t*>(p.planes) + off + (half ? C : 0);
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(d + plane) = make_uint4(l[0], l[1], l[2], l[3]);
            if (KIND == SPLIT_INPUT && p.wg) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint16_t h0, l0, h1, l1;
                    split_pair<FMT_BF16>(e[2 * i], h0, l0);
                    split_pair<FMT_BF16>(e[2 * i + 1], h1, l1);
                    h[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    l[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
                uint16_t* w = reinterpret_cast<uint16_t*>(p.wg) + off;
                *reinterpret_cast<uint4*>(w) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(w + plane) = make_uint4(l[0], l[1], l[2], l[3]);
            }
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

static dim3 split_grid(int B, int C, int T) { return dim3((T + 63) / 64, (C + 63) / 64, B); }

int dv3_tc_split_input(const float* x, void* btc, int npl, void* bct, int B, int C, int T, int k, int dilation,
                       int causal, float p_drop, const unsigned long long* seed_ptr, unsigned salt, void* stream) {
    DV3_REQUIRE(B <= 65535 && (C + 63) / 64 <= 65535, "tc_split_input: grid too large");
    DV3_REQUIRE(npl == 2, "tc_split_input: npl must be 2");
    (void)k; (void)dilation; (void)causal;
    SplitParams p = {};
    p.in0 = x; p.planes = (bf16*)btc; p.wg = (bf16*)bct; p.B = B; p.C = C; p.T = T; p.pitch = (C + 7) / 8 * 8;
    p.p = p_drop; p.seed_ptr = seed_ptr; p.salt = salt;
    launch_k(plane_split_kernel<SPLIT_INPUT>, split_grid(B, C, T), dim3(256), 0, (cudaStream_t)stream, p);
    return check_launch("tc_split_input");
}

int dv3_tc_gate_bwd_split(const float* dy, const float* a, const float* s, const float* x, void* btc, void* bct,
                          float* dbias, int B, int C, int T, int mode, int residual, void* stream) {
    DV3_REQUIRE(bct == nullptr, "tc_gate_bwd_split: bct must be NULL");
    DV3_REQUIRE(C % 8 == 0 && (mode == 0 || x != nullptr), "tc_gate_bwd_split: C %% 8 != 0 or highway without x");
    SplitParams p = {};
    p.in0 = dy; p.in1 = a; p.in2 = s; p.in3 = x; p.planes = (bf16*)btc; p.dbias = dbias;
    p.B = B; p.C = C; p.T = T; p.pitch = 2 * C; p.mode = mode; p.residual = residual;
    launch_k(plane_split_kernel<SPLIT_GATE>, split_grid(B, C, T), dim3(256), 0, (cudaStream_t)stream, p);
    return check_launch("tc_gate_bwd_split");
}

int dv3_tc_grad_split(const float* dy, const float* y, void* btc, void* bct, float* dbias, int B, int C, int T,
                      int relu, void* stream) {
    DV3_REQUIRE(bct == nullptr, "tc_grad_split: bct must be NULL");
    DV3_REQUIRE(!relu || y != nullptr, "tc_grad_split: ReLU backward needs the forward output");
    SplitParams p = {};
    p.in0 = dy; p.in1 = y; p.planes = (bf16*)btc; p.dbias = dbias;
    p.B = B; p.C = C; p.T = T; p.pitch = (C + 7) / 8 * 8; p.relu = relu;
    launch_k(plane_split_kernel<SPLIT_GRAD>, split_grid(B, C, T), dim3(256), 0, (cudaStream_t)stream, p);
    return check_launch("tc_grad_split");
}

// Weight norm + split for a conv weight v (Cout, Cin, k), g [Cout]:
//   wfwd: [2][k][Cout][Cinp] fp16 planes (forward operand: rows co, K = ci)
//   wbwd: [2][k][Cin][Coutp] bf16 planes (data-gradient operand, multiplied with bf16 gradient planes)
int dv3_tc_weightnorm_fwd(const float* v, const float* g, float* inv_norm, float* scale, void* wfwd, int npl,
                          void* wbwd, int Cout, int Cin, int k, void* stream) {
    DV3_REQUIRE(npl == 2, "tc_weightnorm_fwd: npl must be 2");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = Cin * k;
    const long long Cinp = (Cin + 7) / 8 * 8, Coutp = (Cout + 7) / 8 * 8;
    launch_k(wn_norm_kernel2, (Cout * 32 + 255) / 256, 256, 0, st, v, g, inv_norm, scale, Cout, L);
    if (int e = check_launch("tc_weightnorm_fwd(norm)")) return e;
    dim3 grid((L + 31) / 32, (Cout + 31) / 32);
    launch_k(wn_pack_split_kernel<FMT_F16, FMT_BF16>, grid, dim3(32, 8), 0, st, v, scale, (bf16*)wfwd, Cinp, 1,
             (long long)Cout * Cinp, (long long)k * Cout * Cinp, (bf16*)wbwd, 1, Coutp, (long long)Cin * Coutp,
             (long long)k * Cin * Coutp, Cout, Cin, k);
    return check_launch("tc_weightnorm_fwd(pack)");
}

// ConvTranspose1d(k=2,s=2) weight v (Cin, Cout, 2), g [Cin] (norm over dim 0 = Cin), run as a 1x1 conv with
// 2*Cout output rows ordered (j, co):
//   wfwd: [2][2*Cout][Cinp] fp16, rows (j,co), K = ci        wbwd: [2][Cin][K2p] bf16, rows ci, K = (j,co)
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
    launch_k(wn_pack_split_kernel<FMT_BF16, FMT_F16>, grid, dim3(32, 8), 0, st, v, scale, (bf16*)wbwd, K2p, 1,
             (long long)Cout, (long long)Cin * K2p, (bf16*)wfwd, 1, Cinp, (long long)Cout * Cinp,
             (long long)2 * Cout * Cinp, Cin, Cout, 2);
    return check_launch("tc_weightnorm_convt_fwd(pack)");
}

}  // extern "C"
