// fp32 CUDA-core tiled GEMM mainloop shared by every contraction on the path (exact-fp32 math mode):
//   conv forward / data-gradient (implicit GEMM, im2col gathered straight into shared memory),
//   conv weight-gradient, and the batched attention contractions.
//
// CTA tile 128 (M) x BN (N, 128 or 64) x BK=16, 256 threads, thread tile 8 x (BN/16),
// double-buffered shared memory with register-staged prefetch (the loaders transform while they
// load -- dropout mask, tap shift with zero padding -- so cp.async/TMA cannot be used here; the
// tensor-core path in conv_tc.cu is the TMA one).
//
// Shared tile layout: [BK][W] floats, float4 column index XOR-swizzled with (kk & 7) so that both the
// "direct" (consecutive lanes -> consecutive rows) and the "transposed" (4 rows x 8 k per warp
// instruction, for operands that are contiguous along K in HBM) store patterns are bank-conflict
// free, and the compute loop reads operands with LDS.128.
#pragma once
#include "common.cuh"

namespace dv3 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;

__device__ __forceinline__ int swz(int kk, int row) {
    return ((((row >> 2) ^ (kk & 7)) << 2) | (row & 3));
}

// thread -> (kk, row) maps for filling a [BK][W] tile -------------------------------------------
template <int W>
struct DirectMap {                       // consecutive lanes own consecutive rows of one k-slice
    static constexpr int N = GEMM_BK * W / GEMM_THREADS;
    int row_, kk0_;
    __device__ explicit DirectMap(int tid) : row_(tid % W), kk0_(tid / W) {}
    __device__ int row(int) const { return row_; }
    __device__ int kk(int i) const { return kk0_ + i * (GEMM_THREADS / W); }
};
template <int W>
struct TransMap {                        // a warp instruction covers 4 rows x 8 consecutive k
    static constexpr int N = GEMM_BK * W / GEMM_THREADS;
    int rr_, kq_, wi_;
    __device__ explicit TransMap(int tid) : rr_(tid & 3), kq_((tid & 31) >> 2), wi_(tid >> 5) {}
    __device__ int row(int i) const { return ((wi_ + 8 * i) % (W / 4)) * 4 + rr_; }
    __device__ int kk(int i) const { return ((wi_ + 8 * i) / (W / 4)) * 8 + kq_; }
};
template <int W, class Map>
__device__ __forceinline__ void tile_store(float* S, const Map& m, const float* r) {
#pragma unroll
    for (int i = 0; i < Map::N; ++i) S[m.kk(i) * W + swz(m.kk(i), m.row(i))] = r[i];
}

// Accumulator fragment of one thread: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+q (, 64+tx*4+q)}.
template <int BN>
struct Acc {
    static constexpr int NC = BN / 16;   // 8 or 4 columns per thread
    float v[8][NC];
};

template <int BN>
__device__ __forceinline__ void tile_mma(const float* __restrict__ As, const float* __restrict__ Bs,
                                         Acc<BN>& acc, int tx, int ty) {
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
        const int x = kk & 7;
        float a[8], b[Acc<BN>::NC];
        *reinterpret_cast<float4*>(&a[0]) =
            *reinterpret_cast<const float4*>(&As[kk * GEMM_BM + ((ty ^ x) << 2)]);
        *reinterpret_cast<float4*>(&a[4]) =
            *reinterpret_cast<const float4*>(&As[kk * GEMM_BM + (((16 + ty) ^ x) << 2)]);
        *reinterpret_cast<float4*>(&b[0]) =
            *reinterpret_cast<const float4*>(&Bs[kk * BN + ((tx ^ x) << 2)]);
        if (BN == 128)
            *reinterpret_cast<float4*>(&b[Acc<BN>::NC - 4]) =
                *reinterpret_cast<const float4*>(&Bs[kk * BN + (((16 + tx) ^ x) << 2)]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < Acc<BN>::NC; ++q) acc.v[i][q] = fmaf(a[i], b[q], acc.v[i][q]);
    }
}

// Problem policy P provides:
//   struct Params; struct ALoad{ALoad(p,m_tile,z,tid); fetch(chunk,float*); store(float*,const float*)};
//   struct BLoad (same, n_tile);  static int num_chunks(p,z);  static void epilogue(p,acc,m_tile,n_tile,z,tx,ty)
template <class P, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_simt_kernel(const __grid_constant__ typename P::Params p) {
    pdl_trigger(); pdl_wait();     // programmatic dependent launch: see common.cuh
    extern __shared__ __align__(16) float smem[];
    float* As[2] = {smem, smem + GEMM_BK * GEMM_BM};
    float* Bs[2] = {smem + 2 * GEMM_BK * GEMM_BM, smem + 2 * GEMM_BK * GEMM_BM + GEMM_BK * BN};
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y, z = blockIdx.z;

    typename P::ALoad la(p, m_tile, z, tid);
    typename P::template BLoad<BN> lb(p, n_tile, z, tid);
    float ra[P::ALoad::N], rb[P::template BLoad<BN>::N];
    Acc<BN> acc;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < Acc<BN>::NC; ++q) acc.v[i][q] = 0.f;

    const int nchunks = P::num_chunks(p, z);
    if (nchunks > 0) {
        la.fetch(0, ra); lb.fetch(0, rb);
        la.store(As[0], ra); lb.store(Bs[0], rb);
    }
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const bool more = (c + 1 < nchunks);
        if (more) { la.fetch(c + 1, ra); lb.fetch(c + 1, rb); }
        tile_mma<BN>(As[c & 1], Bs[c & 1], acc, tx, ty);
        if (more) { la.store(As[(c + 1) & 1], ra); lb.store(Bs[(c + 1) & 1], rb); }
        __syncthreads();
    }
    P::template epilogue<BN>(p, acc, m_tile, n_tile, z, tx, ty);
}

template <int BN>
constexpr int gemm_smem_bytes() { return 2 * GEMM_BK * (GEMM_BM + BN) * (int)sizeof(float); }

// Pick BN so that the grid has the fewest (waves x tile cost): small problems prefer BN=64.
static inline int pick_bn(long long n_cols, int m_tiles, int z, int num_sms) {
    long long c128 = ((n_cols + 127) / 128) * m_tiles * (long long)z;
    long long c64 = ((n_cols + 63) / 64) * m_tiles * (long long)z;
    long long slots = 2LL * num_sms;
    long long w128 = (c128 + slots - 1) / slots * 128;
    long long w64 = (c64 + slots - 1) / slots * 68;      // BN=64 tiles are slightly less efficient
    return w64 < w128 ? 64 : 128;
}

}  // namespace dv3
