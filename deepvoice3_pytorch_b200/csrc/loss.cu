// Fused training losses + their gradients in one pass each (reference train.py:537-601, 665-740): masked-L1 +
// binary-divergence spectrogram loss (mel and linear), BCE on the done flag, guided-attention loss with the soft
// mask W[b,t,n] = 1 - exp(-(n/N_b - t/T_b)^2 / 2g^2) generated on the fly (the reference builds W with numba on the
// host and uploads it every step).  Each kernel adds its already-normalised contribution to ONE device scalar and
// writes dLoss/dPrediction, so the autograd graph of ~60 elementwise ATen kernels over (16,800,513) tensors
// collapses to 3 launches.  HBM-bound: 8 B read + 4 B written per spectrogram element.
#include "common.cuh"

namespace dv3 {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;            // valid in thread 0 (and lanes of warp 0)
}

// y_hat (B, T, D) predictions, y (B, T, D) targets; pairs (y_hat[b,t], y[b,t+r]) for t < T-r; lengths int64 [B]
// (valid target frames per utterance, in this tensor's time units); grad (B, T, D) fully written.
//   loss += sum (1-bw)*coef1*|d| + bw*coef*z,  coef = w*m/Sm + (1-w)/N,  m = (t+r < len_b)
// priority bins (train.py:559-567): the L1 term becomes (1-pw)*L1(all bins) + pw*L1(bins < pbin), i.e.
//   coef1 = (1-pw)*coef + [d < pbin]*pw*coef*(D/pbin)     (the same masked/plain means over the pbin-wide slice)
__global__ void spec_loss_kernel(const float* __restrict__ y_hat, const float* __restrict__ y,
                                 const long long* __restrict__ lengths, float* __restrict__ grad,
                                 float* __restrict__ loss, int B, int T, int D, int r, float w, float bw,
                                 float eps, int pbin, float pw) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float red[8];
    __shared__ float s_inv_sm;
    if (threadIdx.x == 0) {
        double sm = 0.0;
        for (int b = 0; b < B; ++b) {
            long long v = lengths[b] - r;
            if (v < 0) v = 0;
            if (v > T - r) v = T - r;
            sm += (double)v;
        }
        s_inv_sm = sm > 0 ? (float)(1.0 / (sm * D)) : 0.f;
    }
    __syncthreads();
    const float inv_sm = s_inv_sm;
    const float inv_n = 1.f / ((float)B * (float)(T - r) * (float)D);
    const long long total = (long long)B * T * D;
    const long long shift = (long long)r * D;
    const bool prio = pbin > 0 && pw > 0.f;
    const float prio_gain = prio ? pw * ((float)D / (float)pbin) : 0.f;
    float acc = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long bt = i / D;
        const int t = (int)(bt % T), b = (int)(bt / T);
        float g = 0.f;
        if (t < T - r) {
            const float p = y_hat[i], tg = y[i + shift];
            const float m = (t + r < lengths[b]) ? 1.f : 0.f;
            const float coef = w * m * inv_sm + (1.f - w) * inv_n;
            const float d = p - tg;
            float c1 = 1.f;                                   // L1 weight relative to coef
            if (prio) c1 = (1.f - pw) + ((int)(i - bt * D) < pbin ? prio_gain : 0.f);
            float e = c1 * (1.f - bw) * fabsf(d);
            float de = c1 * (1.f - bw) * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
            if (bw > 0.f) {
                const float L = logf(p + eps) - logf(1.f - p + eps);
                const float u = expf(L);
                e += bw * (-tg * L + log1pf(u));
                de += bw * (u / (1.f + u) - tg) * (1.f / (p + eps) + 1.f / (1.f - p + eps));
            }
            acc += coef * e;
            g = coef * de;
        }
        grad[i] = g;
    }
    const float s = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, s);
}

// done BCE (mean) + guided attention (mean of attn*W):
//   done_hat, done (n_done);  attn (A, B, Td, Ts), in_len / dec_len int64 [B];  grads written in full.
__global__ void aux_loss_kernel(const float* __restrict__ done_hat, const float* __restrict__ done,
                                float* __restrict__ d_done, long long n_done, const float* __restrict__ attn,
                                float* __restrict__ d_attn, const long long* __restrict__ in_len,
                                const long long* __restrict__ dec_len, int A, int B, int Td, int Ts, float sigma,
                                int use_attn, float* __restrict__ loss) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float red[8];
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x, start = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const float inv_nd = 1.f / (float)n_done;
    for (long long i = start; i < n_done; i += stride) {
        const float p = done_hat[i], t = done[i];
        acc -= inv_nd * (t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f));
        d_done[i] = inv_nd * (p - t) / fmaxf((1.f - p) * p, 1e-12f);
    }
    if (use_attn) {
        const long long n_attn = (long long)A * B * Td * Ts;
        const float inv_na = 1.f / (float)n_attn;
        const double inv2g2 = 1.0 / (2.0 * (double)sigma * (double)sigma);
        for (long long i = start; i < n_attn; i += stride) {
            const int n = (int)(i % Ts);
            const int t = (int)((i / Ts) % Td);
            const int b = (int)((i / ((long long)Ts * Td)) % B);
            const long long N = in_len[b], Tl = dec_len[b];
            float wv = 0.f;
            if (n < N && t < Tl) {
                const double q = (double)n / (double)N - (double)t / (double)Tl;
                wv = (float)(1.0 - exp(-q * q * inv2g2));
            }
            acc += inv_na * attn[i] * wv;
            d_attn[i] = inv_na * wv;
        }
    } else if (d_attn) {
        const long long n_attn = (long long)A * B * Td * Ts;
        for (long long i = start; i < n_attn; i += stride) d_attn[i] = 0.f;
    }
    const float s = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, s);
}

}  // namespace dv3

using namespace dv3;

extern "C" {

int dv3_spec_loss(const float* y_hat, const float* y, const long long* lengths, float* grad, float* loss, int B,
                  int T, int D, int r, float masked_loss_weight, float binary_divergence_weight, int priority_bin,
                  float priority_weight, void* stream) {
    DV3_REQUIRE(T > r && r >= 0, "spec_loss: need T > r");
    DV3_REQUIRE(priority_bin >= 0 && priority_bin <= D && priority_weight >= 0.f && priority_weight <= 1.f,
                "spec_loss: priority_bin=%d (D=%d) priority_weight=%g out of range", priority_bin, D, priority_weight);
    long long blocks = ((long long)B * T * D + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(spec_loss_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, y_hat, y, lengths, grad, loss, B, T, D, r,
                                                                   masked_loss_weight, binary_divergence_weight, 1e-8f,
                                                                   priority_bin, priority_weight);
    return check_launch("spec_loss");
}

int dv3_aux_loss(const float* done_hat, const float* done, float* d_done, long long n_done, const float* attn,
                 float* d_attn, const long long* in_len, const long long* dec_len, int A, int B, int Td, int Ts,
                 float sigma, int use_attn, float* loss, void* stream) {
    launch_k(aux_loss_kernel, 148 * 2, 256, 0, (cudaStream_t)stream, done_hat, done, d_done, n_done, attn, d_attn, in_len,
                                                              dec_len, A, B, Td, Ts, sigma, use_attn, loss);
    return check_launch("aux_loss");
}

}  // extern "C"
