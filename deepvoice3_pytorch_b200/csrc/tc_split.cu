// Operand preparation for the tensor-core ConvBlock (tc_gemm.cu): every fp32 operand x becomes two bf16 planes
// hi = bf16(x), lo = bf16(x - hi), written in the K-major layout each implicit GEMM consumes.  These passes are
// pure HBM streaming (read 4 B, write 4-8 B per element) and absorb work the fp32 path does inside its loaders:
// the conv-input dropout mask and the (B,C,T) <-> (B,T,C) layout change.
#include <cuda_bf16.h>
#include "common.cuh"

namespace dv3 {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

struct TapList { int k; int off[8]; };

// x (B,C,T) fp32 -> dropout -> planes in (B,T,C) [forward operand] and, for the weight gradient, k time-shifted
// copies in (k,B,C,T): bct[j][b][c][t] = xd[b][c][t + off_j] (zero outside [0,T)).  The shift is baked into the
// copy because the weight-gradient GEMM contracts over t (its contiguous axis) and a TMA box cannot start at an
// element offset that is not 16-byte aligned.  32(c) x 32(t) tile per CTA, block (32, 8).
__global__ void split_input_kernel(const float* __restrict__ x, bf16* __restrict__ btc_hi,
                                   bf16* __restrict__ btc_lo, bf16* __restrict__ bct_hi,
                                   bf16* __restrict__ bct_lo, int Bn, int C, int T, float p,
                                   const unsigned long long* __restrict__ seed_ptr, unsigned salt,
                                   const TapList taps) {
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
            if (bct_hi) {
                for (int j = 0; j < taps.k; ++j) {
                    const int ts = t + taps.off[j];
                    float vs = 0.f;
                    if (ts >= 0 && ts < T) vs = (ts == t) ? v : x[row + ts] * drop_scale(drop, (uint32_t)(row + ts));
                    bf16 h, l; split_bf16(vs, h, l);
                    const size_t o = (((size_t)j * Bn + b) * C + c) * T + t;
                    bct_hi[o] = h; bct_lo[o] = l;
                }
            }
        }
        tile[threadIdx.y + 8 * i][threadIdx.x] = v;
    }
    __syncthreads();
    if (btc_hi) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 + threadIdx.y + 8 * i, c = c0 + threadIdx.x;
            if (c < C && t < T) {
                bf16 h, l;
                split_bf16(tile[threadIdx.x][threadIdx.y + 8 * i], h, l);
                const size_t o = ((size_t)b * T + t) * C + c;
                btc_hi[o] = h; btc_lo[o] = l;
            }
        }
    }
}

// gate backward (see conv.cu gate_bwd_kernel) producing dAB = [da ; db] directly as bf16 planes in
// (B,T,2C) [data-gradient operand] and (B,2C,T) [weight-gradient operand]; dbias[2C] += sums over (b,t).
__global__ void gate_bwd_split_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                      const float* __restrict__ s, const float* __restrict__ x,
                                      bf16* __restrict__ btc_hi, bf16* __restrict__ btc_lo,
                                      bf16* __restrict__ bct_hi, bf16* __restrict__ bct_lo,
                                      float* __restrict__ dbias, int C, int T, int mode, int residual) {
    __shared__ float ta[32][33], tb[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const float gs = (mode == 0 && residual) ? 0.70710678118654752f : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + threadIdx.y + 8 * i, t = t0 + threadIdx.x;
        float da = 0.f, db = 0.f;
        if (c < C && t < T) {
            const size_t in = ((size_t)b * C + c) * T + t;
            const float g = dy[in] * gs, av = a[in], sv = s[in];
            da = g * sv;
            db = g * ((mode == 0) ? av : (av - x[in])) * sv * (1.f - sv);
            if (bct_hi) {
                const size_t oa = ((size_t)b * 2 * C + c) * T + t, ob = oa + (size_t)C * T;
                bf16 h, l;
                split_bf16(da, h, l); bct_hi[oa] = h; bct_lo[oa] = l;
                split_bf16(db, h, l); bct_hi[ob] = h; bct_lo[ob] = l;
            }
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
            bf16 h, l;
            split_bf16(ta[threadIdx.x][threadIdx.y + 8 * i], h, l); btc_hi[o] = h; btc_lo[o] = l;
            split_bf16(tb[threadIdx.x][threadIdx.y + 8 * i], h, l); btc_hi[o + C] = h; btc_lo[o + C] = l;
        }
    }
}

// weight-norm pack for the tensor-core path: v (R=Cout, X=Cin, k) fp32, scale[R] = g/||v|| ->
//   wb planes [k][Cout][Cin] (forward operand: rows co, K = ci) and wf planes [k][Cin][Cout] (dgrad: rows ci, K = co)
__global__ void wn_pack_split_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                     bf16* __restrict__ wb_hi, bf16* __restrict__ wb_lo,
                                     bf16* __restrict__ wf_hi, bf16* __restrict__ wf_lo, int R, int X, int k) {
    __shared__ float tile[32][33];
    const int L = X * k;
    const int r0 = blockIdx.y * 32, e0 = blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + threadIdx.y + 8 * i, e = e0 + threadIdx.x;
        float w = 0.f;
        if (r < R && e < L) {
            w = v[(size_t)r * L + e] * scale[r];
            const int xx = e / k, j = e - xx * k;
            bf16 h, l; split_bf16(w, h, l);
            const size_t o = ((size_t)j * R + r) * X + xx;
            wb_hi[o] = h; wb_lo[o] = l;
        }
        tile[threadIdx.y + 8 * i][threadIdx.x] = w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = e0 + threadIdx.y + 8 * i, r = r0 + threadIdx.x;
        if (r < R && e < L) {
            const int xx = e / k, j = e - xx * k;
            bf16 h, l; split_bf16(tile[threadIdx.x][threadIdx.y + 8 * i], h, l);
            const size_t o = ((size_t)j * X + xx) * R + r;
            wf_hi[o] = h; wf_lo[o] = l;
        }
    }
}

__global__ void wn_norm_kernel2(const float* __restrict__ v, const float* __restrict__ g,
                                float* __restrict__ inv_norm, float* __restrict__ scale, int R, int L) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* row = v + (size_t)r * L;
    float s = 0.f;
    for (int e = lane; e < L; e += 32) { const float xx = row[e]; s = fmaf(xx, xx, s); }
    s = warp_sum(s);
    if (lane == 0) { const float inv = 1.f / sqrtf(s); inv_norm[r] = inv; scale[r] = g[r] * inv; }
}

}  // namespace dv3

using namespace dv3;

extern "C" {

int dv3_tc_split_input(const float* x, void* btc_hi, void* btc_lo, void* bct_hi, void* bct_lo, int B, int C,
                       int T, int k, int dilation, int causal, float p_drop, const unsigned long long* seed_ptr,
                       unsigned salt, void* stream) {
    DV3_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "tc_split_input: grid too large");
    DV3_REQUIRE(k >= 1 && k <= 8, "tc_split_input: kernel size %d not in [1,8]", k);
    TapList taps;
    taps.k = k;
    const int padl = causal ? (k - 1) * dilation : (k - 1) / 2 * dilation;
    for (int j = 0; j < 8; ++j) taps.off[j] = j < k ? j * dilation - padl : 0;
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    split_input_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, (bf16*)btc_hi, (bf16*)btc_lo,
                                                                      (bf16*)bct_hi, (bf16*)bct_lo, B, C, T, p_drop,
                                                                      seed_ptr, salt, taps);
    return check_launch("tc_split_input");
}

int dv3_tc_gate_bwd_split(const float* dy, const float* a, const float* s, const float* x, void* btc_hi,
                          void* btc_lo, void* bct_hi, void* bct_lo, float* dbias, int B, int C, int T, int mode,
                          int residual, void* stream) {
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    gate_bwd_split_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        dy, a, s, x, (bf16*)btc_hi, (bf16*)btc_lo, (bf16*)bct_hi, (bf16*)bct_lo, dbias, C, T, mode, residual);
    return check_launch("tc_gate_bwd_split");
}

// v (Cout, Cin, k), g (Cout) -> bf16 planes wb [k][Cout][Cin], wf [k][Cin][Cout]; inv_norm, scale: [Cout] fp32.
int dv3_tc_weightnorm_fwd(const float* v, const float* g, float* inv_norm, float* scale, void* wb_hi, void* wb_lo,
                          void* wf_hi, void* wf_lo, int Cout, int Cin, int k, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int L = Cin * k;
    wn_norm_kernel2<<<(Cout * 32 + 255) / 256, 256, 0, st>>>(v, g, inv_norm, scale, Cout, L);
    if (int e = check_launch("tc_weightnorm_fwd(norm)")) return e;
    wn_pack_split_kernel<<<dim3((L + 31) / 32, (Cout + 31) / 32), dim3(32, 8), 0, st>>>(
        v, scale, (bf16*)wb_hi, (bf16*)wb_lo, (bf16*)wf_hi, (bf16*)wf_lo, Cout, Cin, k);
    return check_launch("tc_weightnorm_fwd(pack)");
}

}  // extern "C"
