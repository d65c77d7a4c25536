// Fused audio front-end: preemphasis -> sqrt-Hann STFT (1024 / hop 256, 768-sample zero padding on both sides)
// -> |.| -> { linear: dB, normalise } and { mel filterbank -> dB, normalise } in ONE pass over the waveform.
// Replaces reference audio.py:31-34 (spectrogram) and :46-51 (melspectrogram), which run TWO independent lws
// STFTs per clip on the CPU (ljspeech.py:63-67).  HBM-bound by construction (about 32 FLOP/B): every sample is read
// once, and only the (513 + n_mels) floats the trainer stores per frame are written back.
//
// Work decomposition: one CTA (8 warps) walks 32 consecutive frames of one clip, 8 at a time -- ONE WARP PER FRAME.
//   * the 11*256 raw samples the 8 frames overlap on are staged once in shared memory by cp.async (zero-filled
//     outside the clip), the copy for the next 8 frames in flight while the mel rows of the current ones are formed;
//   * each warp runs the register-resident radix-8 transform of stft_core.cuh (pre-emphasis and window applied as
//     the points are read; 16 complex points per lane, two exchanges through its private work area, __syncwarp only),
//     splits it into the 513-bin half spectrum two bins at a time, writes the normalised dB row straight from the
//     split and leaves the magnitudes in shared memory;
//   * the mel rows are sparse (mel_start / mel_len): the non-zero weights are packed into shared memory once per CTA
//     and every warp forms its share of the filters for all 8 frames at once (lane = frame x 4 bin phases, bank-
//     conflict free because consecutive frames' planes are 4 banks apart);
//   * window, twiddles and split factors are built once per CTA (tables in shared memory), dB through lg2.approx.
// Frames beyond a clip's own count (ragged batches) are zero-filled by the kernel, so callers pass uninitialised
// output buffers.
#include "common.cuh"
#include "stft_core.cuh"

namespace dv3 {

using namespace stftc;

constexpr int FFT_N = 1024, HOP = 256, NH = 512, NBINS = 513, PAD = FFT_N - HOP;
constexpr int STFT_WARPS = 8, STFT_GROUPS = 4, STFT_FRAMES = STFT_WARPS * STFT_GROUPS;     // frames per CTA
constexpr int STAGE_N = (STFT_WARPS + 3) * HOP;                                            // samples 8 frames span
constexpr int MAX_MELS = 128;

struct StftParams {
    const float* wav;          // (nclips, max_len)
    const int* lengths;        // (nclips) valid samples per clip
    const float* mel_basis;    // (n_mels, 513) dense
    const int* mel_start;      // (n_mels) first non-zero bin
    const int* mel_len;        // (n_mels) number of non-zero bins
    float* linear;             // (nclips, max_frames, 513) or null
    float* mel;                // (nclips, max_frames, n_mels) or null
    int max_len, max_frames, n_mels;
    float preemph, min_level_db, ref_level_db;
};

constexpr int MEL_NNZ = 2048;           // packed non-zero mel weights kept in shared memory (680 for the presets)

struct StftSmem {
    float raw[STAGE_N + 4];               // raw[2 + i] = x[s0 + i], raw[1] = x[s0 - 1]: raw samples of the 8-frame group
    f2 win[NH];                           // (w[2n], w[2n+1])
    f2 tw1[7 * 64];                       // W512^(t*k0)
    f2 tw2[7 * 8];                        // W64^(n0*k1)
    f2 wsp[NH / 2 + 2];                   // W1024^k, k <= 256
    int2 melseg[MAX_MELS];                // (start, len) of every mel row
    int meloff[MAX_MELS + 1];             // offset of the row's weights in wt[]
    float wt[MEL_NNZ];
    float work[STFT_WARPS][2][WORK];      // per-warp re / im planes; the magnitudes end up in plane 0
};
static_assert((2 * WORK) % 32 == 4, "frame planes must sit 4 banks apart for the mel stage");

__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// raw samples x[s0-1 .. s0+STAGE_N) of the clip -> sm.raw[1 ..], zero outside [0, len)
__device__ __forceinline__ void stage_async(float* raw, const float* x, int s0, int len, int tid) {
    for (int i = tid; i < STAGE_N + 1; i += STFT_WARPS * 32) {
        const int s = s0 - 1 + i;
        const bool ok = s >= 0 && s < len;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(raw + 1 + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(x + (ok ? s : 0)), "r"(ok ? 4 : 0)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(STFT_WARPS * 32, 3) stft_mel_kernel(const __grid_constant__ StftParams p) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_raw);
    const int clip = blockIdx.y, f_begin = blockIdx.x * STFT_FRAMES, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int len = p.lengths[clip];
    const int nframes = min((len + 2 * PAD - FFT_N + HOP - 1) / HOP + 1, p.max_frames);   // ceil((len+2*768-1024)/256)+1
    const int f_end = min(f_begin + STFT_FRAMES, p.max_frames);

    // frames of this chunk past the clip's own end: zero-fill (contiguous rows)
    {
        const int z0 = max(f_begin, nframes);
        if (z0 < f_end) {
            const size_t row0 = (size_t)clip * p.max_frames + z0;
            if (p.linear) {
                float* o = p.linear + row0 * NBINS;
                for (int i = tid; i < (f_end - z0) * NBINS; i += blockDim.x) o[i] = 0.f;
            }
            if (p.mel) {
                float* o = p.mel + row0 * p.n_mels;
                for (int i = tid; i < (f_end - z0) * p.n_mels; i += blockDim.x) o[i] = 0.f;
            }
        }
    }
    if (f_begin >= nframes) return;

    const float* x = p.wav + (size_t)clip * p.max_len;
    stage_async(sm.raw, x, f_begin * HOP - PAD, len, tid);          // in flight while the tables are built

    // ---- tables, once per CTA ----
    for (int n = tid; n < NH; n += blockDim.x) {
        float w[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * n + e;
            const float hann = 0.5f - 0.5f * cospif((2 * i + 1) / (float)FFT_N);   // 0.5*(1-cos(2*pi*(i+.5)/N))
            w[e] = sqrtf(hann * (2.f * HOP / FFT_N));
        }
        sm.win[n] = {w[0], w[1]};
    }
    for (int i = tid; i < 7 * 64; i += blockDim.x) {
        const int k0 = i / 64 + 1, t = i & 63;
        float s, c; sincospif(-(float)(t * k0) / 256.f, &s, &c);
        sm.tw1[i] = {c, s};
    }
    if (tid < 7 * 8) {
        const int k1 = tid / 8 + 1, n0 = tid & 7;
        float s, c; sincospif(-(float)(n0 * k1) / 32.f, &s, &c);
        sm.tw2[tid] = {c, s};
    }
    for (int k = tid; k <= NH / 2; k += blockDim.x) {
        float s, c; sincospif(-(float)k / 512.f, &s, &c);
        sm.wsp[k] = {c, s};
    }
    if (p.mel) {
        for (int m = tid; m < p.n_mels; m += blockDim.x) sm.melseg[m] = make_int2(p.mel_start[m], p.mel_len[m]);
        if (tid == 0) {
            int off = 0;
            for (int m = 0; m < p.n_mels; ++m) { sm.meloff[m] = off; off += p.mel_len[m]; }
            sm.meloff[p.n_mels] = off;
        }
    }
    __syncthreads();
    const bool packed = p.mel && sm.meloff[p.n_mels] <= MEL_NNZ;    // else the weights stay in global memory
    if (packed)
        for (int m = warp; m < p.n_mels; m += STFT_WARPS) {
            const int2 seg = sm.melseg[m];
            const float* row = p.mel_basis + (size_t)m * NBINS + seg.x;
            for (int j = lane; j < seg.y; j += 32) sm.wt[sm.meloff[m] + j] = row[j];
        }

    // normalised dB: clip((20*log10(max(min_level, v)) - ref - min_db) / -min_db, 0, 1)   (audio.py:79-81, :88-89)
    //   = sat(c2 * log2(max(min_level, v)) + c0);  on p4 = |2X|^2: log2(v) = log2(p4)/2 - 1
    const float inv = 1.f / -p.min_level_db;
    const float c2 = 6.020599913279624f * inv, c0 = 1.f - p.ref_level_db * inv;
    const float min_level = exp2f(p.min_level_db * 0.16609640474436813f);        // 10^(min_db/20)
    const float c2h = 0.5f * c2, c0l = c0 - c2, min_p4 = 4.f * min_level * min_level;

    float* re = sm.work[warp][0];
    float* im = sm.work[warp][1];

    for (int g = 0; g < STFT_GROUPS; ++g) {
        const int f0 = f_begin + g * STFT_WARPS;
        if (f0 >= nframes) break;                                   // uniform over the CTA
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                            // raw[] landed; the previous group's mel stage is done
        const int frame = f0 + warp;
        const size_t fidx = (size_t)clip * p.max_frames + frame;
        if (frame < nframes) {                                      // warp-uniform
            cpx v[2][8];
            pass1(lane, sm.raw + 2 + warp * HOP, p.preemph, len - (frame * HOP - PAD), sm.win, sm.tw1, v);
            store1(lane, v, re, im);
            __syncwarp();
            pass2(lane, re, im, sm.tw2, v);
            __syncwarp();
            store2(lane, v, re, im);
            __syncwarp();
            pass3(lane, re, im, v);
            __syncwarp();
            store3(lane, v, re, im);
            __syncwarp();

            // split into the half spectrum: bins (k, 512-k), k = lane + 32*j; dB row out, magnitudes back into re[]
            float* lin = p.linear ? p.linear + fidx * NBINS : nullptr;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int k = lane + 32 * j;
                if (k <= NH / 2) {
                    float plo, phi;
                    split_pair(k, re, im, sm.wsp[k], plo, phi);
                    if (lin) {
                        lin[k] = __saturatef(fmaf(c2h, lg2_approx(fmaxf(plo, min_p4)), c0l));
                        if (k != NH / 2) lin[NH - k] = __saturatef(fmaf(c2h, lg2_approx(fmaxf(phi, min_p4)), c0l));
                    }
                    // every index is read by exactly one lane (its own pair), so overwriting in place is safe
                    re[k] = 0.5f * sqrt_approx(plo);
                    if (k != NH / 2) re[NH - k] = 0.5f * sqrt_approx(phi);
                }
            }
        }
        __syncthreads();                                            // all 8 frames' magnitudes are in place; raw[] is free
        if (g + 1 < STFT_GROUPS && f0 + STFT_WARPS < nframes)
            stage_async(sm.raw, x, (f0 + STFT_WARPS) * HOP - PAD, len, tid);

        if (p.mel) {
            // lane = 4*frame + phase: the 4 phases of a frame stride a filter's bins, all 8 frames in one go
            const int fl = lane >> 2, q = lane & 3;
            const float* magf = sm.work[fl][0];
            const bool fvalid = f0 + fl < nframes;
            float* out = p.mel + ((size_t)clip * p.max_frames + f0 + fl) * p.n_mels;
            for (int m = warp; m < p.n_mels; m += STFT_WARPS) {
                const int2 seg = sm.melseg[m];
                const float* w = packed ? sm.wt + sm.meloff[m] : p.mel_basis + (size_t)m * NBINS + seg.x;
                const float* mg = magf + seg.x;
                float acc = 0.f;
                for (int jj = q; jj < seg.y; jj += 4) acc = fmaf(w[jj], mg[jj], acc);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (q == 0 && fvalid) out[m] = __saturatef(fmaf(c2, lg2_approx(fmaxf(acc, min_level)), c0));
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace dv3

using namespace dv3;

extern "C" {

// frames produced for a clip of n samples: ceil((n + 2*768 - 1024)/256) + 1   (lws "perfectrec" padding)
int dv3_stft_num_frames(int n_samples) { return (n_samples + 2 * PAD - FFT_N + HOP - 1) / HOP + 1; }

int dv3_stft_mel(const float* wav, const int* lengths, const float* mel_basis, const int* mel_start,
                 const int* mel_len, float* linear, float* mel, int nclips, int max_len, int max_frames,
                 int n_mels, float preemph, float min_level_db, float ref_level_db, void* stream) {
    DV3_REQUIRE(nclips >= 1 && nclips <= 65535, "stft_mel: nclips %d out of range", nclips);
    DV3_REQUIRE(max_frames >= 1, "stft_mel: bad max_frames %d", max_frames);
    DV3_REQUIRE(n_mels >= 0 && n_mels <= MAX_MELS, "stft_mel: n_mels %d > %d", n_mels, MAX_MELS);
    StftParams p = {wav, lengths, mel_basis, mel_start, mel_len, linear, mel, max_len, max_frames, n_mels,
                    preemph, min_level_db, ref_level_db};
    static const cudaError_t attr = cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)sizeof(StftSmem));
    DV3_REQUIRE(attr == cudaSuccess, "stft_mel: cannot reserve %zu bytes of shared memory", sizeof(StftSmem));
    launch_k(stft_mel_kernel, dim3((max_frames + STFT_FRAMES - 1) / STFT_FRAMES, nclips), STFT_WARPS * 32,
             sizeof(StftSmem), (cudaStream_t)stream, p);
    return check_launch("stft_mel");
}

}  // extern "C"
