// Fused audio front-end: preemphasis -> sqrt-Hann STFT (1024 / hop 256, 768-sample zero padding on both sides)
// -> |.| -> { linear: dB, normalise } and { mel filterbank -> dB, normalise } in ONE pass over the waveform.
// Replaces reference audio.py:31-34 (spectrogram) and :46-51 (melspectrogram), which run TWO independent lws
// STFTs per clip on the CPU (ljspeech.py:63-67).  HBM/PCIe-bound by construction (32 FLOP/B): each 1024-sample
// frame is read once from L2-resident waveform, transformed in shared memory, and only the (513 + n_mels) floats the
// trainer stores per frame are written back.
//
// One CTA (256 threads) per frame.  Real-input trick: 1024 real samples -> 512-point complex radix-2 FFT in shared
// memory (one butterfly per thread per stage) -> split into the 513-bin half spectrum.
#include "common.cuh"

namespace dv3 {

constexpr int FFT_N = 1024, HOP = 256, NH = 512, NBINS = 513, PAD = FFT_N - HOP;

__device__ __forceinline__ int bitrev9(int x) { return (int)(__brev((unsigned)x) >> 23); }

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

__device__ __forceinline__ float amp_to_norm_db(float v, float min_level, float min_db, float ref_db) {
    const float s = 20.f * log10f(fmaxf(min_level, v)) - ref_db;     // audio.py:79-81, :33/:49
    return fminf(fmaxf((s - min_db) / -min_db, 0.f), 1.f);            // audio.py:88-89
}

__global__ void __launch_bounds__(256) stft_mel_kernel(const __grid_constant__ StftParams p) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    __shared__ float zr[NH], zi[NH];
    __shared__ float twr[NH / 2], twi[NH / 2];
    __shared__ float mag[NBINS + 3];
    const int clip = blockIdx.y, frame = blockIdx.x, tid = threadIdx.x;
    const int len = p.lengths[clip];
    const int nframes = (len + 2 * PAD - FFT_N + HOP - 1) / HOP + 1;       // ceil((len+2*768-1024)/256)+1
    if (frame >= nframes) return;
    const float* x = p.wav + (size_t)clip * p.max_len;

    // twiddles W512^j = exp(-2*pi*i*j/512), j < 256
    {
        float s, c;
        sincospif(-(float)tid / 256.f, &s, &c);
        twr[tid] = c; twi[tid] = s;
    }
    // load: z[n] = w[2n]*xe[2n] + i*w[2n+1]*xe[2n+1] into bit-reversed position; thread handles n = tid, tid+256
    const int base = frame * HOP - PAD;
    const float wscale = 2.f * HOP / FFT_N;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int n = tid + h * 256;
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * n + e, s = base + i;
            float xe = 0.f;
            if (s >= 0 && s < len) xe = x[s] - (s > 0 ? p.preemph * x[s - 1] : 0.f);
            const float hann = 0.5f - 0.5f * cospif((2 * i + 1) / (float)FFT_N);   // 0.5*(1-cos(2*pi*(i+.5)/N))
            v[e] = xe * sqrtf(hann * wscale);
        }
        const int r = bitrev9(n);
        zr[r] = v[0]; zi[r] = v[1];
    }
    __syncthreads();
    // 9 radix-2 DIT stages, one butterfly per thread
#pragma unroll
    for (int s = 0; s < 9; ++s) {
        const int half = 1 << s;
        const int pos = tid & (half - 1);
        const int i0 = ((tid >> s) << (s + 1)) + pos, i1 = i0 + half;
        const int tw = pos << (8 - s);
        const float wr = twr[tw], wi = twi[tw];
        const float ar = zr[i0], ai = zi[i0], br0 = zr[i1], bi0 = zi[i1];
        const float br = br0 * wr - bi0 * wi, bi = br0 * wi + bi0 * wr;
        zr[i0] = ar + br; zi[i0] = ai + bi;
        zr[i1] = ar - br; zi[i1] = ai - bi;
        __syncthreads();
    }
    // split: X[k] = E + W1024^k * O,  E = (Z[k]+conj(Z[N-k]))/2,  O = -i*(Z[k]-conj(Z[N-k]))/2
    for (int k = tid; k <= NH; k += 256) {
        const int ka = k & (NH - 1), kb = (NH - k) & (NH - 1);
        const float ar = zr[ka], ai = zi[ka], br = zr[kb], bi = -zi[kb];
        const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
        const float dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);
        const float orr = di, oi = -dr;                                  // -i * (dr + i di)
        float s, c;
        sincospif(-(float)k / 512.f, &s, &c);
        const float xr = er + c * orr - s * oi, xi = ei + c * oi + s * orr;
        mag[k] = sqrtf(xr * xr + xi * xi);
    }
    __syncthreads();
    const float min_level = expf(p.min_level_db / 20.f * 2.302585092994046f);
    const size_t fidx = (size_t)clip * p.max_frames + frame;
    if (p.linear) {
        float* out = p.linear + fidx * NBINS;
        for (int k = tid; k < NBINS; k += 256)
            out[k] = amp_to_norm_db(mag[k], min_level, p.min_level_db, p.ref_level_db);
    }
    if (p.mel) {
        const int warp = tid >> 5, lane = tid & 31;
        float* out = p.mel + fidx * p.n_mels;
        for (int m = warp; m < p.n_mels; m += 8) {
            const int st = p.mel_start[m], ln = p.mel_len[m];
            const float* row = p.mel_basis + (size_t)m * NBINS + st;
            float acc = 0.f;
            for (int j = lane; j < ln; j += 32) acc = fmaf(row[j], mag[st + j], acc);
            acc = warp_sum(acc);
            if (lane == 0) out[m] = amp_to_norm_db(acc, min_level, p.min_level_db, p.ref_level_db);
        }
    }
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
    DV3_REQUIRE(max_frames >= dv3_stft_num_frames(max_len) || max_frames > 0, "stft_mel: bad max_frames");
    StftParams p = {wav, lengths, mel_basis, mel_start, mel_len, linear, mel, max_len, max_frames, n_mels,
                    preemph, min_level_db, ref_level_db};
    launch_k(stft_mel_kernel, dim3(max_frames, nclips), 256, 0, (cudaStream_t)stream, p);
    return check_launch("stft_mel");
}

}  // extern "C"
