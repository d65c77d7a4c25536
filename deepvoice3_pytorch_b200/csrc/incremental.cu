// Autoregressive (incremental) decoding, one time step per launch sequence -- reference conv.py:17-46
// (Conv1d.incremental_forward: ring buffer of the last (k-1)*dilation+1 inputs x linearised weight),
// modules.py:145-167 / 200-226 (gate epilogues on a (B,1,C) slice), deepvoice3.py:132-176 (attention with the
// monotonic window) and the decoder loops deepvoice3.py:367-485 / nyanko.py:250-338.
//
// B200 design: a decoder step is ~40 dependent matrix-VECTOR products (B is 1..16, M = 1), i.e. pure weight
// streaming out of L2 (the 10-25 MB of decoder weights stay resident in the 126 MB L2) and launch latency.  So the
// step is a fixed sequence of small kernels with ALL loop state in device memory -- the step counter t, the ring
// buffers, the monotonic-attention cursor, the output arrays indexed by t -- which makes the sequence identical
// from step to step: the host captures it once in a CUDA graph and replays it, checking the done flags only every
// few steps.  Every pointer of a step record can advance by a per-step stride (x + b*ld + t*t_stride), so frames /
// decoder states / alignments are written in place and teacher-forced inputs are read in place.
// Arithmetic is exact fp32 (fmaf accumulation, one warp per output channel).
#include "common.cuh"
#include "../../include/dv3b200.h"

namespace dv3 {

constexpr float kSqrtHalf = 0.70710678118654752f;

// one warp per output channel c (GLU / highway: rows c and C + c); BT batch rows per pass
template <int BT>
__global__ void __launch_bounds__(256) inc_conv_step_kernel(const __grid_constant__ Dv3IncStep p) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int gated = p.mode != 0;
    const int C = gated ? p.Cout / 2 : p.Cout;
    const long long t = p.t_ptr ? (long long)*p.t_ptr : 0;
    const int k = p.k, Cin = p.Cin, L = (k - 1) * p.dilation + 1;
    const int slot_now = (int)(t % L);
    if (warp < C) {
        const int c = warp;
        const float* __restrict__ wa = p.w + (size_t)c * k * Cin;
        const float* __restrict__ wb = p.w + (size_t)(C + c) * k * Cin;
        for (int b0 = 0; b0 < p.B; b0 += BT) {
            float acc_a[BT], acc_b[BT];
#pragma unroll
            for (int i = 0; i < BT; ++i) { acc_a[i] = 0.f; acc_b[i] = 0.f; }
            for (int j = 0; j < k; ++j) {
                const bool cur = (j == k - 1);
                // tap j sees the input of time t - (k-1-j)*dilation; older than the sequence start = zero (ring is
                // zero-initialised), the current input comes straight from x (+ add)
                int slot = 0;
                if (!cur) { const long long tj = t - (long long)(k - 1 - j) * p.dilation; slot = (int)(((tj % L) + L) % L); }
                const float* wja = wa + (size_t)j * Cin;
                const float* wjb = wb + (size_t)j * Cin;
                if (p.vec4) {
                    for (int ci = lane * 4; ci < Cin; ci += 128) {
                        const float4 a4 = *reinterpret_cast<const float4*>(wja + ci);
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (gated) b4 = *reinterpret_cast<const float4*>(wjb + ci);
#pragma unroll
                        for (int i = 0; i < BT; ++i) {
                            const int b = b0 + i;
                            if (b >= p.B) break;
                            float4 x4;
                            if (cur) {
                                x4 = *reinterpret_cast<const float4*>(p.x + b * p.x_ld + t * p.x_t + ci);
                                if (p.add) {
                                    const float4 e4 = *reinterpret_cast<const float4*>(p.add + b * p.add_ld + t * p.add_t + ci);
                                    x4.x += e4.x; x4.y += e4.y; x4.z += e4.z; x4.w += e4.w;
                                }
                            } else {
                                x4 = *reinterpret_cast<const float4*>(p.ring + ((size_t)b * L + slot) * Cin + ci);
                            }
                            acc_a[i] = fmaf(a4.x, x4.x, acc_a[i]); acc_a[i] = fmaf(a4.y, x4.y, acc_a[i]);
                            acc_a[i] = fmaf(a4.z, x4.z, acc_a[i]); acc_a[i] = fmaf(a4.w, x4.w, acc_a[i]);
                            if (gated) {
                                acc_b[i] = fmaf(b4.x, x4.x, acc_b[i]); acc_b[i] = fmaf(b4.y, x4.y, acc_b[i]);
                                acc_b[i] = fmaf(b4.z, x4.z, acc_b[i]); acc_b[i] = fmaf(b4.w, x4.w, acc_b[i]);
                            }
                        }
                    }
                } else {
                    for (int ci = lane; ci < Cin; ci += 32) {
                        const float a1 = wja[ci];
                        const float b1 = gated ? wjb[ci] : 0.f;
#pragma unroll
                        for (int i = 0; i < BT; ++i) {
                            const int b = b0 + i;
                            if (b >= p.B) break;
                            float xv;
                            if (cur) {
                                xv = p.x[b * p.x_ld + t * p.x_t + ci];
                                if (p.add) xv += p.add[b * p.add_ld + t * p.add_t + ci];
                            } else {
                                xv = p.ring[((size_t)b * L + slot) * Cin + ci];
                            }
                            acc_a[i] = fmaf(a1, xv, acc_a[i]);
                            if (gated) acc_b[i] = fmaf(b1, xv, acc_b[i]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < BT; ++i) {
                const int b = b0 + i;
                if (b >= p.B) break;                       // uniform across the warp
                float a = warp_sum(acc_a[i]);
                float g = gated ? warp_sum(acc_b[i]) : 0.f;
                if (lane != 0) continue;
                a += p.bias[c];
                float y;
                if (p.mode == 0) {
                    y = a;
                    if (p.act == 1) y = fmaxf(y, 0.f);
                    else if (p.act == 2) y = sigmoidf_(y);
                } else {
                    g += p.bias[C + c];
                    const float s = sigmoidf_(g);
                    const float xin = p.x[b * p.x_ld + t * p.x_t + c];      // gated blocks: Cin == C
                    if (p.mode == 1) {
                        if (p.spk) a += p.spk[b * p.spk_ld + c];
                        y = a * s;
                    } else {
                        y = s * a + (1.f - s) * xin;
                    }
                }
                if (p.res1) y = (y + p.res1[b * p.res1_ld + t * p.res1_t + c]) * kSqrtHalf;
                if (p.res2) y = (y + p.res2[b * p.res2_ld + t * p.res2_t + c]) * kSqrtHalf;
                p.y[b * p.y_ld + t * p.y_t + c] = y;
                if (p.y2) {
                    float y2 = y;
                    if (p.y2_mode == 1) y2 = sigmoidf_(y);
                    else if (p.y2_mode == 2) y2 = y + p.yadd[b * p.yadd_ld + t * p.yadd_t + c];
                    p.y2[b * p.y2_ld + t * p.y2_t + c] = y2;
                }
            }
        }
    }
    // the last CTA also files the current input into the ring (slot t mod L is not read by anybody this step)
    if (p.ring && blockIdx.x == gridDim.x - 1) {
        for (int i = threadIdx.x; i < p.B * Cin; i += blockDim.x) {
            const int b = i / Cin, ci = i - b * Cin;
            float xv = p.x[b * p.x_ld + t * p.x_t + ci];
            if (p.add) xv += p.add[b * p.add_ld + t * p.add_t + ci];
            p.ring[((size_t)b * L + slot_now) * Cin + ci] = xv;
        }
    }
}

// one CTA per batch row: scores = q . keys, monotonic window, softmax, context = probs . values * Ts*sqrt(1/Ts)
__global__ void __launch_bounds__(256) inc_attn_step_kernel(const __grid_constant__ Dv3IncAttn p) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    extern __shared__ float sm[];
    float* q = sm;                 // [E]
    float* sc = sm + p.E;          // [Ts]
    __shared__ float red[8];
    __shared__ float bcast;
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long t = p.t_ptr ? (long long)*p.t_ptr : 0;
    for (int e = tid; e < p.E; e += 256) q[e] = p.q[b * p.q_ld + e];
    __syncthreads();
    int lo = 0, hi = p.Ts;                                  // unmasked key range
    if (p.last_attended) {
        const int la = p.last_attended[t & 1];
        const int backward = la - p.window_backward;
        if (backward > 0) lo = backward;
        const int ahead = la + p.window_ahead;
        if (ahead < p.Ts) hi = ahead;
    }
    const float* __restrict__ K = p.keys + (size_t)b * p.E * p.Ts;
    float mx = -INFINITY;
    for (int s = tid; s < p.Ts; s += 256) {
        float acc = 0.f;
        for (int e = 0; e < p.E; ++e) acc = fmaf(q[e], K[(size_t)e * p.Ts + s], acc);
        if (s < lo || s >= hi) acc = -INFINITY;
        sc[s] = acc;
        mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) { float m = red[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]); bcast = m; }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int s = tid; s < p.Ts; s += 256) { const float e = expf(sc[s] - mx); sc[s] = e; sum += e; }
    sum = warp_sum(sum);
    __syncthreads();                                        // everybody has read bcast
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) { float s = 0.f; for (int i = 0; i < 8; ++i) s += red[i]; bcast = s; }
    __syncthreads();
    const float inv = 1.f / bcast;
    for (int s = tid; s < p.Ts; s += 256) {
        const float pr = sc[s] * inv;
        sc[s] = pr;
        if (p.align) p.align[b * p.align_ld + t * p.align_t + s] = pr * p.align_scale;
    }
    __syncthreads();
    if (p.last_attended && b == 0 && tid == 0) {            // reference: alignment.max(-1)[1] of batch row 0
        int best = 0; float bv = sc[0];
        for (int s = 1; s < p.Ts; ++s) if (sc[s] > bv) { bv = sc[s]; best = s; }
        p.last_attended[(t + 1) & 1] = best;
    }
    const float* __restrict__ V = p.values + (size_t)b * p.Ts * p.E;
    const float scale = (float)p.Ts * sqrtf(1.0f / (float)p.Ts);
    for (int e = tid; e < p.E; e += 256) {
        float acc = 0.f;
        for (int s = 0; s < p.Ts; ++s) acc = fmaf(sc[s], V[(size_t)s * p.E + e], acc);
        p.ctx[b * p.ctx_ld + e] = acc * scale;
    }
}

__global__ void inc_advance_kernel(int* t) {
    pdl_trigger(); pdl_wait(); *t += 1; }

}  // namespace dv3

using namespace dv3;

extern "C" {

int dv3_inc_conv_step(const Dv3IncStep* p, void* stream) {
    DV3_REQUIRE(p && p->B > 0 && p->Cin > 0 && p->Cout > 0 && p->k >= 1 && p->dilation >= 1,
                "inc_conv_step: bad shape");
    DV3_REQUIRE(p->mode == 0 || (p->Cout == 2 * p->Cin), "inc_conv_step: gated blocks need Cout == 2*Cin");
    DV3_REQUIRE(p->k == 1 || p->ring != nullptr, "inc_conv_step: k > 1 needs a ring buffer");
    const int C = p->mode != 0 ? p->Cout / 2 : p->Cout;
    const int blocks = (C * 32 + 255) / 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->B == 1) launch_k(inc_conv_step_kernel<1>, blocks, 256, 0, st, *p);
    else if (p->B == 2) launch_k(inc_conv_step_kernel<2>, blocks, 256, 0, st, *p);
    else launch_k(inc_conv_step_kernel<4>, blocks, 256, 0, st, *p);
    return check_launch("inc_conv_step");
}

int dv3_inc_attn_step(const Dv3IncAttn* p, void* stream) {
    DV3_REQUIRE(p && p->B > 0 && p->E > 0 && p->Ts > 0, "inc_attn_step: bad shape");
    const size_t smem = (size_t)(p->E + p->Ts) * sizeof(float);
    DV3_REQUIRE(smem <= 48 * 1024, "inc_attn_step: E + Ts = %d floats exceed 48 KB of shared memory", p->E + p->Ts);
    launch_k(inc_attn_step_kernel, p->B, 256, smem, (cudaStream_t)stream, *p);
    return check_launch("inc_attn_step");
}

int dv3_inc_advance(int* t_ptr, void* stream) {
    launch_k(inc_advance_kernel, 1, 1, 0, (cudaStream_t)stream, t_ptr);
    return check_launch("inc_advance");
}

}  // extern "C"
