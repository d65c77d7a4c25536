// tcgen05 tensor-core path of the ConvBlock and of the plain (1x1 / k-tap) weight-normed convolutions: forward,
// data-gradient and weight-gradient as implicit GEMMs with fp32-equivalent accuracy from split-bf16 operands.
// Every fp32 operand is pre-split into bf16 planes p0 = bf16(x), p1 = bf16(x - p0) (tc_split.cu, which also applies
// the input dropout and emits the K-major layout each GEMM wants) and each K-step issues p0*p0 + p0*p1 + p1*p0
// (operand error ~2^-17; single-pass TF32 would miss the rtol=1e-3/atol=1e-4 parity bar after ~30 blocks).
// Accuracy note (measured, tools/precision_report.py): the tensor core adds each MMA into the fp32 accumulator with
// truncation, a bias of ~N_mma x 2^-25 relative -- with one accumulator that dominated the operand-split error and
// made a 3-plane / 6-product variant WORSE than this one.  So the main term p0*p0 and the 2^-8-smaller cross terms
// accumulate in two separate TMEM accumulators (the cross terms then truncate at 2^-8 of the scale and the main one
// sees a third of the events) which the epilogue adds in fp32.
//
//   GATED : D[t, (a|b) c] = sum_{j,ci} Xd[b, t+off_j, ci] * W[j, (a|b) c, ci]    M = 128 time steps, N = 128 a | 128 b
//   CONV  : D[t, n]       = sum_{j,kc} A[b, t+off_j, kc]  * W[j, n, kc]           M = 128 time steps, N = NBOX x 128
//           (plain conv forward with bias/ReLU, and every data gradient: A = dY or dAB, W = transposed weight)
//   WGRAD : D[m, n] (j)   = sum_{b,t}  dY[b, m, t] * Xs_j[b, n, t]                 M = 128 rows,       N = NBOX x 128
//
// All operands are K-major bf16 tiles of 128 rows x BK (BK = 64: 128-byte rows, SWIZZLE_128B -- the default wherever
// the channel count is a multiple of 64; BK = 32: 64-byte rows, SWIZZLE_64B) fetched by TMA; the conv's zero padding,
// the causal shift, ragged T / channel tails are TMA out-of-bounds zero fill.  Warp roles (192 threads): warp 0 = TMA
// producer, warp 1 = TMEM owner + single-thread MMA issuer, warps 2-5 = epilogue (TMEM -> registers -> fused gate /
// bias / mask / residual math -> stores coalesced along T).
//
// Kernels in this file (selection logic + measurements at dv3_tc_convblock_fwd / dv3_tc_conv below):
//   tc_conv_kernel          one output tile per CTA (small layers; 64- or 128-column tiles; optional weight multicast)
//   tc_conv_persist_kernel  one CTA per SM walks the tile list, two accumulator sets in TMEM (epilogue of tile n
//                           overlaps the MMAs of tile n+1)                                   <- default for big layers
//   tc_conv_taps_kernel     persistent + activation rows fetched once per channel slice for all taps (opt-in)
//   tc_conv_pair_kernel     persistent + CTA pairs, tcgen05 cta_group::2, M = 256 (opt-in); tc_conv_pair64_kernel
//                           = its BK = 64 variant (round-2 candidate, not yet run)
//   tc_wgrad_mn_kernel      weight gradient from MN-major operands
// In every kernel the hi and lo weight planes sit back to back in the stage, so p0(A) x [p0(W) ; p1(W)] is ONE
// N = 2*NCOLS MMA filling the main | cross accumulators, followed by p1(A) x p0(W) into the cross accumulator.
#include "tc_common.cuh"

namespace dv3 {

using namespace tc;

constexpr int TC_THREADS = 192;
constexpr int MAX_TAPS_TC = 8;
constexpr int SMEM_LIMIT = 232448;          // 227 KB opt-in dynamic shared memory per CTA

enum { TC_GATED = 0, TC_CONV = 1, TC_WGRAD = 2 };

struct TcMaps { CUtensorMap a[2]; CUtensorMap b[2]; CUtensorMap bs[2]; };   // bs: B boxes of 128/CL rows (multicast slices)

struct TcParams {
    int T, B;
    int Kc;                    // contraction channels per tap (GATED: C, CONV: A channels); WGRAD: unused
    int Nc;                    // output channels (GATED: C per half, CONV: out channels, WGRAD: N = Cin)
    int Mw;                    // WGRAD: rows of dY
    int rows_per_tap;          // rows of the weight matrix per tap (GATED: 2C, CONV: Nc)
    int k, kb_n;               // taps; K blocks per tap (GATED/CONV) or time chunks per batch (WGRAD)
    int tap_off[MAX_TAPS_TC];
    // gated epilogue
    const float* bias; const float* spk; const float* res;
    float* y; float* save_a; float* save_s;
    int gate_mode, residual;
    // conv epilogue: out = acc*dropmask + bias + addend ; relu
    float* out; const float* e1; const float* e2; float alpha; int addmode, relu;
    float p_drop; const unsigned long long* seed_ptr; uint32_t salt;
    // wgrad
    float* dw; long long split_stride; int nsplit, batches_per_split;
    int msplit; long long s_m, s_mh, s_n, s_j;
    int debug;                 // DV3_TC_DEBUG bit 0: epilogue skipped, bit 1: MMAs skipped, bit 2: TMA loads skipped (timing experiments only)
};

template <int BK> struct SwizzleOf;
template <> struct SwizzleOf<64> { static constexpr uint32_t layout = 2, sbo = 1024; };   // SWIZZLE_128B
template <> struct SwizzleOf<32> { static constexpr uint32_t layout = 4, sbo = 512; };    // SWIZZLE_64B

template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address / 16
    d |= (uint64_t)(SwizzleOf<BK>::sbo >> 4) << 32;          // stride between 8-row swizzle atoms
    d |= (uint64_t)1 << 46;                                  // sm_100 descriptor version
    d |= (uint64_t)SwizzleOf<BK>::layout << 61;
    return d;
}

// BR = rows of one B-operand box (128, or 64 for problems too small to fill the machine with 128-wide tiles)
template <int NBOX, int BK, int NPL, int BR = 128>
struct TcCfg {
    static constexpr int TILE = 128 * BK * 2;            // A tile (128 rows)
    static constexpr int TILE_B = BR * BK * 2;           // one B box
    static constexpr int STAGE = NPL * (TILE + NBOX * TILE_B);
    static constexpr int STAGES_RAW = (SMEM_LIMIT - 2048) / STAGE;
    // (measured: 3 stages + 2 CTAs/SM for the 64-row variants -- epilogue of one CTA overlapping the K loop of the
    // other -- is 8-13 % faster on the big shapes but 7-9 % slower on the small ones they exist for; not kept)
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int SMEM = STAGES * STAGE + 1024 + 512;
    static constexpr int NCOLS = BR * NBOX;          // columns per accumulator; two accumulators (main, cross)
    static constexpr int TMEM_COLS = 2 * NCOLS;
};

// main + cross accumulator -> registers, summed in fp32 (round-to-nearest)
__device__ __forceinline__ void tmem_ld_add(uint32_t taddr, int cross_off, float* v) {
    float c[32];
    tmem_ld_32x32(taddr, v);
    tmem_ld_32x32(taddr + cross_off, c);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += c[i];
}

// ---- epilogues (shared by the one-tile-per-CTA and the persistent kernels) ---------------------------------
// NOTE on the epilogue loads: residual / addend / bias reads go through __ldg (ld.global.nc) and are issued as a
// batch of 32 independent loads BEFORE the dependent math and stores of the chunk.  With plain loads the compiler
// must order every load after the previous iteration's stores (possible aliasing), which serialised 128
// global-memory round trips per thread and made the epilogue as long as the whole K loop (ncu: 40 % of the stall
// samples sat on the first use of these loads).
template <int BR, int NCOLS>
__device__ __forceinline__ void epilogue_gated(const TcParams& p, uint32_t taddr, int a_row0, int a_z, int b_row0,
                                               int row) {
        const int t = a_row0 + row, b = a_z, C = p.Nc;
        const bool tv = t < p.T;
        const float* __restrict__ bias = p.bias;
        const float* __restrict__ res = p.res;
        const float* __restrict__ spk = p.spk;
        float* __restrict__ yo = p.y;
        float* __restrict__ ao = p.save_a;
        float* __restrict__ so = p.save_s;
        const bool need_res = (p.gate_mode != 0) || p.residual;
        const size_t base = ((size_t)b * C + b_row0) * p.T + (tv ? t : 0);
        for (int c32 = 0; c32 < BR; c32 += 32) {
            float va[32], vb[32], rr[32];
            tmem_ld_add(taddr + c32, NCOLS, va);
            tmem_ld_add(taddr + BR + c32, NCOLS, vb);
            if (!tv) continue;
            const size_t cb = base + (size_t)c32 * p.T;
#pragma unroll
            for (int i = 0; i < 32; ++i) rr[i] = need_res ? __ldg(&res[cb + (size_t)i * p.T]) : 0.f;
            if (spk) {
#pragma unroll
                for (int i = 0; i < 32; ++i) va[i] += __ldg(&spk[cb + (size_t)i * p.T]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int c = b_row0 + c32 + i;
                const size_t idx = cb + (size_t)i * p.T;
                const float a = va[i] + __ldg(&bias[c]);
                const float s = sigmoidf_(vb[i] + __ldg(&bias[C + c]));
                float y;
                if (p.gate_mode == 0) {
                    y = a * s;
                    if (p.residual) y = (y + rr[i]) * 0.70710678118654752f;
                } else {
                    y = s * a + (1.f - s) * rr[i];
                }
                yo[idx] = y;
                if (ao) ao[idx] = a;
                if (so) so[idx] = s;
            }
        }
}

template <int NCOLS>
__device__ __forceinline__ void epilogue_conv(const TcParams& p, uint32_t taddr, int a_row0, int a_z, int b_row0,
                                              int row) {
        const int t = a_row0 + row, b = a_z;
        const bool tv = t < p.T;
        const DropCfg drop = make_drop(p.p_drop, p.seed_ptr, p.salt);
        const float* __restrict__ bias = p.bias;
        const float* __restrict__ e1 = p.e1;
        const float* __restrict__ e2 = p.e2;
        float* __restrict__ out = p.out;
        for (int c32 = 0; c32 < NCOLS; c32 += 32) {
            float v[32], x1[32], x2[32];
            tmem_ld_add(taddr + c32, NCOLS, v);
            if (!tv) continue;
            const int n0 = b_row0 + c32;
            const size_t cb = ((size_t)b * p.Nc + n0) * p.T + t;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const bool ok = n0 + i < p.Nc;
                x1[i] = (p.addmode != 0 && ok) ? __ldg(&e1[cb + (size_t)i * p.T]) : 0.f;
                x2[i] = (p.addmode == 2 && ok) ? __ldg(&e2[cb + (size_t)i * p.T]) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = n0 + i;
                if (n >= p.Nc) continue;
                const size_t idx = cb + (size_t)i * p.T;
                float g = v[i] * drop_scale(drop, (uint32_t)idx);
                if (bias) g += __ldg(&bias[n]);
                if (p.addmode == 1) g += p.alpha * x1[i];
                else if (p.addmode == 2) g += x1[i] * (1.f - x2[i]);
                if (p.relu) g = fmaxf(g, 0.f);
                out[idx] = g;
            }
        }
}

// CL > 1: thread-block cluster of CL CTAs along the batch axis.  They need the same weight tiles, so CTA r fetches
// rows [r*128/CL, (r+1)*128/CL) of every weight box and TMA-multicasts them into all CL shared memories: weight
// bytes read from L2 per CTA drop by CL (the kernels are L2->SMEM bandwidth bound, ~3.8 TB/s measured).
template <int MODE, int NBOX, int BK, int NPL, int CL, int BR>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    using Cfg = TcCfg<NBOX, BK, NPL, BR>;
    constexpr int TILE = Cfg::TILE, TILE_B = Cfg::TILE_B, STAGE = Cfg::STAGE, STAGES = Cfg::STAGES, NCOLS = Cfg::NCOLS;
    constexpr int B_OFF = NPL * TILE;                    // B planes start after the A planes
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates ----------------------------------------------------------------------
    int a_row0, a_z = 0, b_row0, b_row1, n_iters, wg_j = 0, wg_split = 0, b_beg = 0;
    if (MODE == TC_GATED) {
        a_row0 = blockIdx.x * 128; a_z = blockIdx.z;                 // t0, batch
        b_row0 = blockIdx.y * BR; b_row1 = p.Nc + blockIdx.y * BR;   // a half, b half
        n_iters = p.k * p.kb_n;
    } else if (MODE == TC_CONV) {
        a_row0 = blockIdx.x * 128; a_z = blockIdx.z;
        b_row0 = blockIdx.y * BR * NBOX; b_row1 = b_row0 + BR;       // n0
        n_iters = p.k * p.kb_n;
    } else {
        wg_j = blockIdx.z % p.k; wg_split = blockIdx.z / p.k;
        a_row0 = blockIdx.y * 128;                                   // m0
        b_row0 = blockIdx.x * BR * NBOX; b_row1 = b_row0 + BR;       // n0
        b_beg = wg_split * p.batches_per_split;
        int b_end = b_beg + p.batches_per_split; if (b_end > p.B) b_end = p.B;
        n_iters = (b_end > b_beg ? b_end - b_beg : 0) * p.kb_n;
    }

    static_assert(CL == 1 || MODE != TC_WGRAD, "weight-gradient tiles share no operand across the batch");
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1);
    constexpr int SLICE_ROWS = BR / CL, SLICE_BYTES = SLICE_ROWS * BK * 2;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) { prefetch_tmap(&maps.a[i]); prefetch_tmap(CL > 1 ? &maps.bs[i] : &maps.b[i]); }
        // empty[s] collects one tcgen05.commit arrival from every CTA of the cluster (all of them read the slices
        // this CTA multicasts); full[s] gets this CTA's expect_tx arrival + bytes from all CL producers
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // every CTA's barriers are initialised before any remote arrival
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * STAGE;
            int ax, ay, az, bx, by0, by1, bz;
            if (MODE == TC_WGRAD) {
                const int bi = it / p.kb_n, tc_ = it - bi * p.kb_n;
                ax = tc_ * BK; ay = a_row0; az = b_beg + bi;
                // the tap shift is baked into the wg_j-th shifted copy of the input (tc_split.cu): a TMA box
                // cannot start at a K (time) coordinate that is not 16-byte aligned
                bx = tc_ * BK; by0 = b_row0; by1 = b_row1; bz = wg_j * p.B + b_beg + bi;
            } else {
                const int j = it / p.kb_n, kb = it - j * p.kb_n;
                ax = kb * BK; ay = a_row0 + p.tap_off[j]; az = a_z;
                bx = kb * BK; by0 = j * p.rows_per_tap + b_row0; by1 = j * p.rows_per_tap + b_row1; bz = 0;
            }
            mbar_arrive_expect_tx(&full[s], STAGE);
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
                tma_load_3d(st + pl * TILE, &maps.a[pl], &full[s], ax, ay, az);
                uint8_t* bdst = st + B_OFF + pl * NBOX * TILE_B;
                if (CL == 1) {
                    tma_load_3d(bdst, &maps.b[pl], &full[s], bx, by0, bz);
                    if (NBOX == 2) tma_load_3d(bdst + TILE_B, &maps.b[pl], &full[s], bx, by1, bz);
                } else {
                    const int ro = crank * SLICE_ROWS, so = crank * SLICE_BYTES;
                    tma_load_3d_multicast(bdst + so, &maps.bs[pl], &full[s], bx, by0 + ro, bz, CL_MASK);
                    if (NBOX == 2)
                        tma_load_3d_multicast(bdst + TILE_B + so, &maps.bs[pl], &full[s], bx, by1 + ro, bz, CL_MASK);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOLS);
        constexpr bool MERGE_HI = NPL == 2 && 2 * NCOLS <= 256;
        constexpr uint32_t idesc2 = make_idesc_bf16(128, MERGE_HI ? 2 * NCOLS : NCOLS);
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE);
            uint64_t da[NPL], db[NPL];
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
                da[pl] = make_desc<BK>(sa + pl * TILE);
                db[pl] = make_desc<BK>(sa + B_OFF + pl * NBOX * TILE_B);
            }
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 2);             // 16 bf16 = 32 bytes, in 16-byte units
                if (MERGE_HI) {
                    // p0 x [p0 ; p1]: the two weight planes sit back to back in the stage, so ONE N = 2*NCOLS MMA
                    // fills the main | cross accumulators and reads the activation tile from shared memory once
                    umma_bf16(tmem_base, da[0] + adv, db[0] + adv, idesc2, (it | kk) != 0);
                } else {
                    umma_bf16(tmem_base, da[0] + adv, db[0] + adv, idesc, (it | kk) != 0);          // main accumulator
                    umma_bf16(tmem_base + NCOLS, da[0] + adv, db[1] + adv, idesc, (it | kk) != 0);  // cross accumulator
                }
                umma_bf16(tmem_base + NCOLS, da[1] + adv, db[0] + adv, idesc, 1);
            }
            if (CL == 1) umma_commit(&empty[s]);                     // frees the stage once these MMAs retire
            else umma_commit_multicast(&empty[s], CL_MASK);          // ... in every CTA that multicasts into it
        }
        umma_commit(tmem_full);
    } else if (warp >= 2) {
        // ================= epilogue =================
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;                                      // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;                               // accumulator row (M index)
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (MODE == TC_GATED) {
            epilogue_gated<BR, NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
        } else if (MODE == TC_CONV) {
            epilogue_conv<NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
        } else {
            const int m = a_row0 + row;
            float* __restrict__ out = p.dw + (size_t)wg_split * p.split_stride + (size_t)wg_j * p.s_j;
            const size_t ma = (size_t)(m % p.msplit) * p.s_m + (size_t)(m / p.msplit) * p.s_mh;
            // partials with unit stride along n ([split][j][m][n]): each thread owns a contiguous run -> float4
            // stores, every 32-byte sector written whole (the v-layout (m, n, j) scatters 4-byte stores k apart)
            const bool vec = (p.s_n == 1) && ((p.Nc & 3) == 0) && (((ma + (size_t)wg_j * p.s_j +
                                                                     (size_t)wg_split * p.split_stride) & 3) == 0);
            for (int c32 = 0; c32 < NCOLS; c32 += 32) {
                float v[32];
                tmem_ld_add(taddr + c32, NCOLS, v);
                if (m >= p.Mw) continue;
                const int n0 = b_row0 + c32;
                if (n_iters == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                if (vec && n0 + 32 <= p.Nc) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(&out[ma + n0 + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + i < p.Nc) out[ma + (size_t)(n0 + i) * p.s_n] = v[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // no CTA may exit while peers still multicast into / signal it
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Persistent variant of the GATED / CONV kernels: one CTA per SM walks a static round-robin list of output tiles.
// TMEM holds TWO accumulator sets (main + cross each), so the epilogue warps drain tile n (TMEM -> gate math ->
// stores) while the MMA thread already accumulates tile n+1, and the shared-memory ring keeps streaming across tile
// boundaries; barrier / TMEM / tensor-map set-up is paid once per SM instead of once per tile.
// N per tile is limited to 128 columns (4 x 128 = 512 TMEM columns).
// ------------------------------------------------------------------------------------------------
template <int MODE, int NBOX, int BR, int BK = 32>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_persist_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p, int tiles_x,
                       int tiles_y, int num_tiles) {
    constexpr int NPL = 2;
    using Cfg = TcCfg<NBOX, BK, NPL, BR>;
    constexpr int TILE = Cfg::TILE, TILE_B = Cfg::TILE_B, STAGE = Cfg::STAGE, STAGES = Cfg::STAGES, NCOLS = Cfg::NCOLS;
    constexpr int B_OFF = NPL * TILE;
    static_assert(4 * NCOLS <= 512, "two accumulator sets of (main + cross) must fit in TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;          // [2] accumulator set ready for the epilogue
    uint64_t* tempty = tfull + 2;              // [2] accumulator set drained
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iters = p.k * p.kb_n;

    if (threadIdx.x == 0) {
        prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.a[1]); prefetch_tmap(&maps.b[0]); prefetch_tmap(&maps.b[1]);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<4 * NCOLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // tile id -> (time tile, channel tile, batch); channel tiles vary fastest so that concurrently running CTAs share
    // the activation tile in L2
    auto decode = [&](int tile, int& a_row0, int& a_z, int& b_row0, int& b_row1) {
        const int ty = tile % tiles_y, r = tile / tiles_y;
        const int tx = r % tiles_x;
        a_z = r / tiles_x;
        a_row0 = tx * 128;
        if (MODE == TC_GATED) { b_row0 = ty * BR; b_row1 = p.Nc + ty * BR; }
        else { b_row0 = ty * BR * NBOX; b_row1 = b_row0 + BR; }
    };

    if (warp == 0 && lane == 0) {
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            for (int kit = 0; kit < n_iters; ++kit, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE;
                const int j = kit / p.kb_n, kb = kit - j * p.kb_n;
                const int ax = kb * BK, ay = a_row0 + p.tap_off[j];
                const int by0 = j * p.rows_per_tap + b_row0, by1 = j * p.rows_per_tap + b_row1;
                if (p.debug & 4) { mbar_arrive(&full[s]); continue; }
                mbar_arrive_expect_tx(&full[s], STAGE);
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    tma_load_3d(st + pl * TILE, &maps.a[pl], &full[s], ax, ay, a_z);
                    uint8_t* bdst = st + B_OFF + pl * NBOX * TILE_B;
                    tma_load_3d(bdst, &maps.b[pl], &full[s], ax, by0, 0);
                    if (NBOX == 2) tma_load_3d(bdst + TILE_B, &maps.b[pl], &full[s], ax, by1, 0);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOLS), idesc2 = make_idesc_bf16(128, 2 * NCOLS);
        int it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tempty[a], aph ^ 1);                         // the epilogue has drained this accumulator set
            tc_fence_after();
            const uint32_t acc = tmem_base + a * 2 * NCOLS;
            for (int kit = 0; kit < n_iters; ++kit, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE);
                const uint64_t da0 = make_desc<BK>(sa), da1 = make_desc<BK>(sa + TILE);
                const uint64_t db0 = make_desc<BK>(sa + B_OFF);         // plane 1 follows plane 0: rows [NCOLS, 2 NCOLS)
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    const uint64_t adv = (uint64_t)(kk * 2);
                    if (p.debug & 2) continue;
                    umma_bf16(acc, da0 + adv, db0 + adv, idesc2, (kit | kk) != 0);     // p0 x [p0 ; p1] -> main | cross
                    umma_bf16(acc + NCOLS, da1 + adv, db0 + adv, idesc, 1);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(&tfull[a]);
        }
    } else if (warp >= 2) {
        const int q = warp & 3, row = q * 32 + lane;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tfull[a], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + a * 2 * NCOLS + ((uint32_t)(q * 32) << 16);
            if (!(p.debug & 1)) {
                if (MODE == TC_GATED) epilogue_gated<BR, NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
                else epilogue_conv<NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
            }
            tc_fence_before();
            mbar_arrive(&tempty[a]);                                // 128 arrivals release the set to the MMA thread
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<4 * NCOLS>(tmem_base);
}

template <int MODE, int NBOX, int BR, int BK = 32>
static int launch_tc_persist(const TcMaps& maps, const TcParams& p, int tiles_x, int tiles_y, int batch,
                             cudaStream_t st, const char* what) {
    using Cfg = TcCfg<NBOX, BK, 2, BR>;
    constexpr int SMEM = Cfg::STAGES * Cfg::STAGE + 1024 + 512;
    static bool configured = false;
    auto kern = tc_conv_persist_kernel<MODE, NBOX, BR, BK>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("%s: cannot set %d B dynamic smem: %s", what, SMEM, cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int num_tiles = tiles_x * tiles_y * batch;
    const int grid = num_tiles < sms ? num_tiles : sms;
    kern<<<grid, TC_THREADS, SMEM, st>>>(maps, p, tiles_x, tiles_y, num_tiles);
    return check_launch(what);
}

static int g_persist = -1;
static int tc_persist() {                  // DV3_TC_PERSIST=0 disables the persistent kernels
    if (g_persist < 0) { const char* e = getenv("DV3_TC_PERSIST"); g_persist = (e && atoi(e) == 0) ? 0 : 1; }
    return g_persist;
}

// ------------------------------------------------------------------------------------------------
// Tap-reuse variant of the persistent kernel (k > 1).  ncu on the kernels above: both the gated forward and the MN
// weight gradient move ~10 TB/s from L2 into shared memory (l1tex__m_xbar2l1tex_read_bytes / duration) -- the
// fabric limit -- while the tensor pipe idles ~45 % of the time: the kernels are L2->SM bound, and a third to a half
// of that traffic is the SAME activation rows fetched once per tap.  Here one pipeline stage holds, for one
// 32-channel slice, the activation rows [t0 + off_min, t0 + off_max + 128) ONCE (a single TMA box of
// 128 + (k-1)*dilation rows) plus the k weight boxes; tap j's MMAs read the A operand through a descriptor whose
// start address is advanced by (off_j - off_min) rows of 64 bytes.  (SWIZZLE_64B is a function of the absolute
// shared-memory address bits, so a row-shifted start inside a 512-byte-aligned tile addresses exactly what TMA
// wrote.)  L2->SM bytes per output tile drop from k*(A + B) to A*(1 + span/128) + k*B.
// The stage geometry depends on k and the dilation, so stage size / count are run-time values.
// ------------------------------------------------------------------------------------------------
struct TapGeom { int a_rows, a_plane, b_box, stage, stages, off_min, base_off; };

template <int NBOX, int BR>
static TapGeom tap_geom(int k, const int* tap_off) {
    int lo = tap_off[0], hi = tap_off[0];
    for (int j = 1; j < k; ++j) { lo = tap_off[j] < lo ? tap_off[j] : lo; hi = tap_off[j] > hi ? tap_off[j] : hi; }
    TapGeom g;
    g.off_min = lo;
    g.a_rows = (128 + (hi - lo) + 7) / 8 * 8;
    g.a_plane = (g.a_rows * 64 + 1023) / 1024 * 1024;
    g.b_box = BR * 64;
    g.stage = 2 * g.a_plane + 2 * k * NBOX * g.b_box;
    g.stages = (SMEM_LIMIT - 2048) / g.stage;
    if (g.stages > 6) g.stages = 6;
    { static int cap = -1; if (cap < 0) { const char* e = getenv("DV3_TC_MAXSTAGES"); cap = e ? atoi(e) : 0; }
      if (cap >= 2 && g.stages > cap) g.stages = cap; }
    // debugging aid: DV3_TC_TAPS_BASEOFF=1 also writes (start >> 7) & 7 into the descriptor's base-offset field
    static int bo = -1;
    if (bo < 0) { const char* e = getenv("DV3_TC_TAPS_BASEOFF"); bo = (e && atoi(e) == 1) ? 1 : 0; }
    g.base_off = bo;
    return g;
}

// CL = 2: the two CTAs of a cluster work on two BATCHES of the same (time tile, channel tile): they need the same
// weight boxes, so CTA r fetches only weight plane r of every tap and TMA-multicasts it into both shared memories
// (the kernel is L2-feed bound and weights are 3/4 of its bytes: 64.6 -> 40.6 KB per CTA and channel slice).
// empty[s] then collects one commit from each CTA (the peer's producer writes into this CTA's stage too).
template <int MODE, int NBOX, int BR, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_taps_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p,
                    const __grid_constant__ TapGeom g, int tiles_x, int tiles_y, int num_tiles) {
    constexpr int BK = 32, NPL = 2;
    constexpr int NCOLS = BR * NBOX;
    static_assert(4 * NCOLS <= 512, "two accumulator sets of (main + cross) must fit in TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int STAGES = g.stages, STAGE = g.stage;
    const int B_OFF = NPL * g.a_plane;                 // weight boxes: [tap][plane][box]
    const int B_PLANE = p.k * NBOX * g.b_box;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iters = p.kb_n;                         // one stage per 32-channel slice, all taps inside
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    const int first_tile = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int tile_step = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;

    if (threadIdx.x == 0) {
        prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.a[1]); prefetch_tmap(&maps.b[0]); prefetch_tmap(&maps.b[1]);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<4 * NCOLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                    // the peer's barriers exist before anything is multicast into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto decode = [&](int tile, int& a_row0, int& a_z, int& b_row0, int& b_row1) {
        const int ty = tile % tiles_y, r = tile / tiles_y;
        const int tx = r % tiles_x;
        a_z = CL * (r / tiles_x) + (int)rank;
        a_row0 = tx * 128;
        if (MODE == TC_GATED) { b_row0 = ty * BR; b_row1 = p.Nc + ty * BR; }
        else { b_row0 = ty * BR * NBOX; b_row1 = b_row0 + BR; }
    };

    if (warp == 0 && lane == 0) {
        int it = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            for (int kb = 0; kb < n_iters; ++kb, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE;
                const int ax = kb * BK;
                if (CL == 1 && (p.debug & 4)) { mbar_arrive(&full[s]); continue; }   // timing experiment: no loads at all
                mbar_arrive_expect_tx(&full[s], NPL * (g.a_rows * 64 + B_PLANE));
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    tma_load_3d(st + pl * g.a_plane, &maps.a[pl], &full[s], ax, a_row0 + g.off_min, a_z);
                    if (CL > 1 && pl != (int)rank) continue;       // the peer multicasts the other weight plane
                    for (int j = 0; j < p.k; ++j) {
                        uint8_t* bdst = st + B_OFF + (j * NPL + pl) * NBOX * g.b_box;
                        if (CL > 1) {
                            tma_load_3d_multicast(bdst, &maps.b[pl], &full[s], ax, j * p.rows_per_tap + b_row0, 0, 3);
                            if (NBOX == 2)
                                tma_load_3d_multicast(bdst + g.b_box, &maps.b[pl], &full[s], ax,
                                                      j * p.rows_per_tap + b_row1, 0, 3);
                        } else {
                            tma_load_3d(bdst, &maps.b[pl], &full[s], ax, j * p.rows_per_tap + b_row0, 0);
                            if (NBOX == 2)
                                tma_load_3d(bdst + g.b_box, &maps.b[pl], &full[s], ax, j * p.rows_per_tap + b_row1, 0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOLS), idesc2 = make_idesc_bf16(128, 2 * NCOLS);
        int it = 0, tcount = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++tcount) {
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + a * 2 * NCOLS;
            for (int kb = 0; kb < n_iters; ++kb, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE);
                for (int j = 0; j < p.k; ++j) {
                    const uint32_t ashift = (uint32_t)(p.tap_off[j] - g.off_min) * 64u;
                    uint64_t da0 = make_desc<BK>(sa + ashift), da1 = make_desc<BK>(sa + g.a_plane + ashift);
                    if (g.base_off) {
                        da0 |= (uint64_t)(((sa + ashift) >> 7) & 7) << 49;
                        da1 |= (uint64_t)(((sa + g.a_plane + ashift) >> 7) & 7) << 49;
                    }
                    const uint64_t db0 = make_desc<BK>(sa + B_OFF + j * NPL * NBOX * g.b_box);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2);
                        if (p.debug & 2) continue;
                        umma_bf16(acc, da0 + adv, db0 + adv, idesc2, (kb | j | kk) != 0);   // p0 x [p0 ; p1]
                        umma_bf16(acc + NCOLS, da1 + adv, db0 + adv, idesc, 1);
                    }
                }
                if (CL > 1) umma_commit_multicast(&empty[s], 3);     // the stage is free once BOTH CTAs consumed it
                else umma_commit(&empty[s]);
            }
            umma_commit(&tfull[a]);
        }
    } else if (warp >= 2) {
        const int q = warp & 3, row = q * 32 + lane;
        int tcount = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++tcount) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tfull[a], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + a * 2 * NCOLS + ((uint32_t)(q * 32) << 16);
            if (!(p.debug & 1)) {
                if (MODE == TC_GATED) epilogue_gated<BR, NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
                else epilogue_conv<NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
            }
            tc_fence_before();
            mbar_arrive(&tempty[a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                    // no CTA exits while its peer still multicasts into / signals it
    if (warp == 1) tmem_dealloc<4 * NCOLS>(tmem_base);
}

template <int MODE, int NBOX, int BR, int CL = 1>
static int launch_tc_taps(const TcMaps& maps, const TcParams& p, const TapGeom& g, int tiles_x, int tiles_y, int batch,
                          cudaStream_t st, const char* what) {
    static bool configured = false;
    auto kern = tc_conv_taps_kernel<MODE, NBOX, BR, CL>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("%s: cannot set %d B dynamic smem: %s", what, SMEM_LIMIT, cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int smem = g.stages * g.stage + 1024 + 512;
    if (CL == 1) {
        const int num_tiles = tiles_x * tiles_y * batch;
        const int grid = num_tiles < sms ? num_tiles : sms;
        kern<<<grid, TC_THREADS, smem, st>>>(maps, p, g, tiles_x, tiles_y, num_tiles);
    } else {
        const int num_tiles = tiles_x * tiles_y * (batch / CL);      // cluster work units
        int clusters = sms / CL;
        if (num_tiles < clusters) clusters = num_tiles;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, p, g, tiles_x, tiles_y, num_tiles);
        if (e != cudaSuccess) { set_error("%s: cluster launch failed: %s", what, cudaGetErrorString(e)); return 1; }
    }
    return check_launch(what);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the tap-reuse kernel (tcgen05 cta_group::2).  The shared-memory port is what bounds the
// kernels above (operand reads by the MMAs + TMA writes <= 128 B/clk/SM); in a pair the two SMs of a TPC execute one
// M = 256 MMA: each reads its OWN 128 activation rows (CTA r = batch 2*bp + r, same time tile, same output
// channels) but only HALF of the weight rows, so weight bytes fetched, written and read per SM halve.
// With the split-bf16 scheme the halves fall out naturally:
//     MMA 1:  p0(A) x [p0(W) ; p1(W)]   N = 256   -> CTA 0 holds p0(W) (128 rows), CTA 1 holds p1(W)
//     MMA 2:  p1(A) x  p0(W)            N = 128   -> CTA 0 holds rows 0..63 of p0(W), CTA 1 rows 64..127
// Per CTA and tap: 12 KB of weights instead of 16 KB, 14 KB of operand reads per K=16 step instead of 20 KB.
// Protocol: both CTAs run a TMA producer; all loads of a stage credit the LEADER's full barrier; the leader's MMA
// thread issues for the pair and its commits arrive on the empty / accumulator-full barriers of BOTH CTAs; the
// epilogue warps of both CTAs release an accumulator set by arriving on the leader's barrier (256 arrivals).
// ------------------------------------------------------------------------------------------------
struct PairGeom { int a_rows, a_plane, stage, stages, off_min; };

static PairGeom pair_geom(int k, const int* tap_off) {
    int lo = tap_off[0], hi = tap_off[0];
    for (int j = 1; j < k; ++j) { lo = tap_off[j] < lo ? tap_off[j] : lo; hi = tap_off[j] > hi ? tap_off[j] : hi; }
    PairGeom g;
    g.off_min = lo;
    g.a_rows = (128 + (hi - lo) + 7) / 8 * 8;
    g.a_plane = (g.a_rows * 64 + 1023) / 1024 * 1024;
    g.stage = 2 * g.a_plane + k * 12288;
    g.stages = (SMEM_LIMIT - 2048) / g.stage;
    if (g.stages > 6) g.stages = 6;
    { static int cap = -1; if (cap < 0) { const char* e = getenv("DV3_TC_MAXSTAGES"); cap = e ? atoi(e) : 0; }
      if (cap >= 2 && g.stages > cap) g.stages = cap; }
    return g;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_conv_pair_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p,
                    const __grid_constant__ PairGeom g, int tiles_x, int tiles_y, int num_tiles) {
    constexpr int BK = 32, NCOLS = 128;
    constexpr int B_TAP = 12288;                        // per tap: [plane r: 128 rows][p0 rows 64r..64r+63]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int STAGES = g.stages, STAGE = g.stage;
    const int B_OFF = 2 * g.a_plane;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_iters = p.kb_n;

    if (threadIdx.x == 0) {
        prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.a[1]); prefetch_tmap(&maps.b[0]); prefetch_tmap(&maps.b[1]);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 256); }
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();                                 // both CTAs' barriers exist before any remote signal
    if (warp == 1) tmem_alloc_pair<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // pair tile -> (time tile, channel tile, batch pair); this CTA's batch = 2*pair + rank
    auto decode = [&](int tile, int& a_row0, int& a_z, int& b_row0, int& b_row1) {
        const int ty = tile % tiles_y, r = tile / tiles_y;
        const int tx = r % tiles_x;
        a_z = 2 * (r / tiles_x) + (int)rank;
        a_row0 = tx * 128;
        if (MODE == TC_GATED) { b_row0 = ty * 64; b_row1 = p.Nc + ty * 64; }
        else { b_row0 = ty * 128; b_row1 = b_row0 + 64; }
    };

    if (warp == 0 && lane == 0) {
        int it = 0;
        const uint32_t stage_tx = 2u * (uint32_t)(2 * g.a_rows * 64 + p.k * B_TAP);     // both CTAs' bytes
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            for (int kb = 0; kb < n_iters; ++kb, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait_cluster(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE;
                const int ax = kb * BK;
                if (rank == 0) mbar_arrive_expect_tx(&full[s], stage_tx);
                const uint32_t lead_full = mapa_u32(&full[s], 0);
                tma_load_3d_pair(st, &maps.a[0], lead_full, ax, a_row0 + g.off_min, a_z);
                tma_load_3d_pair(st + g.a_plane, &maps.a[1], lead_full, ax, a_row0 + g.off_min, a_z);
                for (int j = 0; j < p.k; ++j) {
                    uint8_t* bd = st + B_OFF + j * B_TAP;
                    const int r0 = j * p.rows_per_tap + b_row0, r1 = j * p.rows_per_tap + b_row1;
                    // this CTA's half of [p0 ; p1]: plane `rank`, both 64-row boxes
                    tma_load_3d_pair(bd, &maps.b[rank], lead_full, ax, r0, 0);
                    tma_load_3d_pair(bd + 4096, &maps.b[rank], lead_full, ax, r1, 0);
                    // this CTA's half of p0: box `rank`
                    tma_load_3d_pair(bd + 8192, &maps.b[0], lead_full, ax, rank == 0 ? r0 : r1, 0);
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        constexpr uint32_t idesc1 = make_idesc_bf16(256, 2 * NCOLS), idesc2 = make_idesc_bf16(256, NCOLS);
        int it = 0, tcount = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tcount) {
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait_cluster(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + a * 2 * NCOLS;
            for (int kb = 0; kb < n_iters; ++kb, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait_cluster(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE);
                for (int j = 0; j < p.k; ++j) {
                    const uint32_t ashift = (uint32_t)(p.tap_off[j] - g.off_min) * 64u;
                    const uint64_t da0 = make_desc<BK>(sa + ashift), da1 = make_desc<BK>(sa + g.a_plane + ashift);
                    const uint64_t dby = make_desc<BK>(sa + B_OFF + j * B_TAP);
                    const uint64_t dbx = make_desc<BK>(sa + B_OFF + j * B_TAP + 8192);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2);
                        if (p.debug & 2) continue;
                        umma_bf16_pair(acc, da0 + adv, dby + adv, idesc1, (kb | j | kk) != 0);   // p0 x [p0 ; p1]
                        umma_bf16_pair(acc + NCOLS, da1 + adv, dbx + adv, idesc2, 1);            // p1 x p0
                    }
                }
                umma_commit_pair(&empty[s]);
            }
            umma_commit_pair(&tfull[a]);
        }
    } else if (warp >= 2) {
        const int q = warp & 3, row = q * 32 + lane;
        int tcount = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tcount) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait_cluster(&tfull[a], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + a * 2 * NCOLS + ((uint32_t)(q * 32) << 16);
            if (!(p.debug & 1)) {
                if (MODE == TC_GATED) epilogue_gated<64, NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
                else epilogue_conv<NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
            }
            tc_fence_before();
            if (rank == 0) mbar_arrive(&tempty[a]);
            else mbar_arrive_remote(&tempty[a], 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                 // the peer's shared memory / barriers stay alive until both are done
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel with 128-byte operand rows and one (tap, 64-channel slice) per stage -- the round-2 candidate
// (opt-in: DV3_TC_PAIR64=1 when there are more tiles than SMs, =2 whenever the batch is even).
// Why: the decomposition experiments put the MMA stream itself at 75 % of nominal -- a fixed
// ~30 clk per tcgen05.mma on top of 128 / 64 clk of math for the N = 256 / N = 128 instructions of the split-bf16
// scheme.  cta_group::2 doubles the work per instruction (M = 256) at the same fixed cost and halves the weight bytes
// per SM; the first pair kernel above showed neither effect because it also carried the BK = 32 rows (twice the TMA
// operations per byte) and the row-shifted tap-reuse descriptors (MMAs 18 % slower).  This variant keeps the pair
// protocol of tc_conv_pair_kernel (validated bit-for-bit) and changes only the stage geometry:
//     stage (56 KB, 4 deep) = A: 2 planes x 128 rows x 128 B   |   B: [plane r: 128 rows][p0 rows 64r..64r+63] x 128 B
// STATUS: compiles; NOT yet run on a GPU (round 1 ran out of GPU minutes) -- default off, first item of round 2.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_conv_pair64_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p, int tiles_x,
                      int tiles_y, int num_tiles) {
    constexpr int BK = 64, NCOLS = 128;
    constexpr int A_PLANE = 128 * BK * 2;               // 16 KB
    constexpr int B_OFF = 2 * A_PLANE;
    // weights per stage: [this CTA's half of (p0 ; p1): two 64-row boxes = B_Y bytes][its half of p0: one box = B_X]
    constexpr int B_Y = 128 * BK * 2, B_X = 64 * BK * 2;
    constexpr int STAGE = B_OFF + B_Y + B_X;            // 56 KB
    constexpr int STAGES = (SMEM_LIMIT - 2048) / STAGE; // 4
    static_assert(STAGES >= 2, "pipeline needs at least two stages");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_iters = p.k * p.kb_n;                   // kb_n = 64-channel slices per tap

    if (threadIdx.x == 0) {
        prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.a[1]); prefetch_tmap(&maps.b[0]); prefetch_tmap(&maps.b[1]);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 256); }
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_alloc_pair<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto decode = [&](int tile, int& a_row0, int& a_z, int& b_row0, int& b_row1) {
        const int ty = tile % tiles_y, r = tile / tiles_y;
        const int tx = r % tiles_x;
        a_z = 2 * (r / tiles_x) + (int)rank;
        a_row0 = tx * 128;
        if (MODE == TC_GATED) { b_row0 = ty * 64; b_row1 = p.Nc + ty * 64; }
        else { b_row0 = ty * 128; b_row1 = b_row0 + 64; }
    };

    if (warp == 0 && lane == 0) {
        int it = 0;
        constexpr uint32_t stage_tx = 2u * (uint32_t)STAGE;               // both CTAs' bytes
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            for (int kit = 0; kit < n_iters; ++kit, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait_cluster(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE;
                const int j = kit / p.kb_n, kb = kit - j * p.kb_n;
                const int ax = kb * BK, ay = a_row0 + p.tap_off[j];
                const int r0 = j * p.rows_per_tap + b_row0, r1 = j * p.rows_per_tap + b_row1;
                if (rank == 0) mbar_arrive_expect_tx(&full[s], stage_tx);
                const uint32_t lead_full = mapa_u32(&full[s], 0);
                tma_load_3d_pair(st, &maps.a[0], lead_full, ax, ay, a_z);
                tma_load_3d_pair(st + A_PLANE, &maps.a[1], lead_full, ax, ay, a_z);
                uint8_t* bd = st + B_OFF;
                tma_load_3d_pair(bd, &maps.b[rank], lead_full, ax, r0, 0);              // plane `rank`, box 0
                tma_load_3d_pair(bd + B_X, &maps.b[rank], lead_full, ax, r1, 0);        // plane `rank`, box 1
                tma_load_3d_pair(bd + B_Y, &maps.b[0], lead_full, ax, rank == 0 ? r0 : r1, 0);   // p0, box `rank`
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        constexpr uint32_t idesc1 = make_idesc_bf16(256, 2 * NCOLS), idesc2 = make_idesc_bf16(256, NCOLS);
        int it = 0, tcount = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tcount) {
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait_cluster(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + a * 2 * NCOLS;
            for (int kit = 0; kit < n_iters; ++kit, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait_cluster(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE);
                const uint64_t da0 = make_desc<BK>(sa), da1 = make_desc<BK>(sa + A_PLANE);
                const uint64_t dby = make_desc<BK>(sa + B_OFF), dbx = make_desc<BK>(sa + B_OFF + B_Y);
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    const uint64_t adv = (uint64_t)(kk * 2);
                    if (p.debug & 2) continue;
                    umma_bf16_pair(acc, da0 + adv, dby + adv, idesc1, (kit | kk) != 0);   // p0 x [p0 ; p1]
                    umma_bf16_pair(acc + NCOLS, da1 + adv, dbx + adv, idesc2, 1);         // p1 x p0
                }
                umma_commit_pair(&empty[s]);
            }
            umma_commit_pair(&tfull[a]);
        }
    } else if (warp >= 2) {
        const int q = warp & 3, row = q * 32 + lane;
        int tcount = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tcount) {
            int a_row0, a_z, b_row0, b_row1;
            decode(tile, a_row0, a_z, b_row0, b_row1);
            const int a = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait_cluster(&tfull[a], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + a * 2 * NCOLS + ((uint32_t)(q * 32) << 16);
            if (!(p.debug & 1)) {
                if (MODE == TC_GATED) epilogue_gated<64, NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
                else epilogue_conv<NCOLS>(p, taddr, a_row0, a_z, b_row0, row);
            }
            tc_fence_before();
            if (rank == 0) mbar_arrive(&tempty[a]);
            else mbar_arrive_remote(&tempty[a], 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

static bool pair64_usable(int B, int Kc, long long num_tiles) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("DV3_TC_PAIR64"); mode = e ? atoi(e) : 0; }
    return mode > 0 && (B & 1) == 0 && Kc % 64 == 0 && (mode > 1 || num_tiles > 148);
}

template <int MODE>
static int launch_tc_pair64(const TcMaps& maps, const TcParams& p, int tiles_x, int tiles_y, int batch,
                            cudaStream_t st, const char* what) {
    constexpr int STAGE = 2 * 128 * 64 * 2 + 128 * 64 * 2 + 64 * 64 * 2;
    constexpr int SMEM = ((SMEM_LIMIT - 2048) / STAGE) * STAGE + 1024 + 512;
    static bool configured = false;
    auto kern = tc_conv_pair64_kernel<MODE>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("%s: cannot set %d B dynamic smem: %s", what, SMEM, cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int num_tiles = tiles_x * tiles_y * (batch / 2);
    int clusters = sms / 2;
    if (num_tiles < clusters) clusters = num_tiles;
    kern<<<2 * clusters, TC_THREADS, SMEM, st>>>(maps, p, tiles_x, tiles_y, num_tiles);
    return check_launch(what);
}

static int g_pair = -1;
static int tc_pair() {                     // DV3_TC_PAIR: 0 = off, 1 = when there are more tiles than SMs, 2 = whenever possible
    if (g_pair < 0) { const char* e = getenv("DV3_TC_PAIR"); g_pair = e ? atoi(e) : 0; }
    return g_pair;
}
static bool pair_usable(int B, int k, const int* tap_off, long long num_tiles64) {
    const int mode = tc_pair();
    if (mode <= 0 || (B & 1) || (mode == 1 && num_tiles64 <= 148)) return false;
    const PairGeom g = pair_geom(k, tap_off);
    return g.a_rows <= 256 && g.stages >= 2;
}

template <int MODE>
static int launch_tc_pair(const TcMaps& maps, const TcParams& p, const PairGeom& g, int tiles_x, int tiles_y,
                          int batch, cudaStream_t st, const char* what) {
    static bool configured = false;
    auto kern = tc_conv_pair_kernel<MODE>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("%s: cannot set %d B dynamic smem: %s", what, SMEM_LIMIT, cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int num_tiles = tiles_x * tiles_y * (batch / 2);
    int clusters = sms / 2;
    if (num_tiles < clusters) clusters = num_tiles;
    const int smem = g.stages * g.stage + 1024 + 512;
    kern<<<2 * clusters, TC_THREADS, smem, st>>>(maps, p, g, tiles_x, tiles_y, num_tiles);
    return check_launch(what);
}

static int persist_bk() {                 // 128-byte rows / SWIZZLE_128B in the persistent kernels; DV3_TC_PERSIST_BK=32: 64-byte rows
    static int v = -1;
    if (v < 0) { const char* e = getenv("DV3_TC_PERSIST_BK"); v = (e && atoi(e) == 32) ? 32 : 64; }
    return v;
}
static int small_bk() {                   // DV3_TC_SMALL_BK=32: the same for the one-tile-per-CTA kernels
    static int v = -1;
    if (v < 0) { const char* e = getenv("DV3_TC_SMALL_BK"); v = (e && atoi(e) == 32) ? 32 : 64; }
    return v;
}

static int g_taps = -1;
static int tc_taps() {                     // DV3_TC_TAPS: 0 = off (default), 1 = when there are more tiles than SMs, 2 = always
    if (g_taps < 0) { const char* e = getenv("DV3_TC_TAPS"); g_taps = e ? atoi(e) : 0; }
    return g_taps;
}

// usable when every tap fits one TMA box (<= 256 rows) and at least two stages fit in shared memory.  Measured
// (tools/tc_time.py): 7-10 % faster than the per-tap persistent kernel on the (16,512,800) blocks, on par at
// (16,256,800), 2-4 % slower on the <= 148-tile shapes (coarser stages, longer pipeline fill); but its row-shifted
// descriptors make the MMAs themselves ~18 % slower (MMA-only time 105 vs 89 us) and the BK = 64 persistent kernel beats
// it (107 vs 119 us) -> opt-in only.
static int g_mcast = -1;
static bool taps_mcast(int B) {            // DV3_TC_MCAST=1: weight multicast across 2-CTA clusters (even batch sizes)
    if (g_mcast < 0) { const char* e = getenv("DV3_TC_MCAST"); g_mcast = e ? atoi(e) : 0; }
    return g_mcast > 0 && (B & 1) == 0;
}

template <int NBOX, int BR>
static bool taps_usable(int k, const int* tap_off, long long num_tiles) {
    const int mode = tc_taps();
    if (mode <= 0 || k < 2 || (mode == 1 && num_tiles <= 148)) return false;
    const TapGeom g = tap_geom<NBOX, BR>(k, tap_off);
    return g.a_rows <= 256 && g.stages >= 2;
}

// ------------------------------------------------------------------------------------------------
// Weight gradient straight from the (B,T,C) planes the forward / data-gradient GEMMs already use:
//     D[m, n] (tap j) = sum_{b,t} dY[b, t, m] * Xd[b, t + off_j, n]
// Both operands are "MN-major" here (channels contiguous, the contraction index t is the row): 64-channel x 32-row
// TMA boxes (128-byte rows, SWIZZLE_128B), UMMA descriptors with the MN-major canonical layout
// ((64 channels contiguous, chunk stride LBO), (8 rows x 128 B, group stride SBO)) and a_major = b_major = 1 in the
// instruction descriptor.  The tap shift is a ROW coordinate of the TMA box (any alignment, out-of-bounds rows are
// zero = the conv padding), so no time-shifted copies of the input are needed.
// ------------------------------------------------------------------------------------------------
struct TcMnParams {
    int T, B, Mw, Nw, k;
    int tap_off[MAX_TAPS_TC];
    int nsplit, batches_per_split, kb_n;      // kb_n = 32-row time chunks per utterance
    uint32_t lbo, sbo;                        // descriptor strides in bytes (chunk stride, 8-row group stride)
    float* dw; long long split_stride;
    int msplit; long long s_m, s_mh, s_n, s_j;
    int debug;                 // DV3_TC_DEBUG bit 0: epilogue skipped, bit 1: MMAs skipped, bit 2: TMA loads skipped (timing experiments only)
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

template <int NBOX>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_wgrad_mn_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcMnParams p) {
    constexpr int BOX = 64 * 32 * 2;                     // 64 channels x 32 time steps of bf16 = 4 KB
    constexpr int A_PL = 2 * BOX, B_PL = 2 * NBOX * BOX; // per plane: 128 rows of M, 128*NBOX columns of N
    constexpr int STAGE = 2 * (A_PL + B_PL);
    constexpr int STAGES = ((SMEM_LIMIT - 2048) / STAGE) > 6 ? 6 : ((SMEM_LIMIT - 2048) / STAGE);
    constexpr int NCOLS = 128 * NBOX;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int wg_j = blockIdx.z % p.k, wg_split = blockIdx.z / p.k;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128 * NBOX;
    const int b_beg = wg_split * p.batches_per_split;
    int b_end = b_beg + p.batches_per_split; if (b_end > p.B) b_end = p.B;
    const int n_iters = (b_end > b_beg ? b_end - b_beg : 0) * p.kb_n;

    if (threadIdx.x == 0) {
        prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.a[1]); prefetch_tmap(&maps.b[0]); prefetch_tmap(&maps.b[1]);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<2 * NCOLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * STAGE;
            const int bi = it / p.kb_n, tc_ = it - bi * p.kb_n;
            const int b = b_beg + bi, t0 = tc_ * 32;
            mbar_arrive_expect_tx(&full[s], STAGE);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_3d(st + pl * A_PL + h * BOX, &maps.a[pl], &full[s], m0 + h * 64, t0, b);
#pragma unroll
                for (int q = 0; q < 2 * NBOX; ++q)
                    tma_load_3d(st + 2 * A_PL + pl * B_PL + q * BOX, &maps.b[pl], &full[s], n0 + q * 64,
                                t0 + p.tap_off[wg_j], b);
            }
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOLS) | (1u << 15) | (1u << 16);   // A and B MN-major
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {                        // 2 x UMMA_K(16 rows of 128 B)
                const uint32_t ko = kk * 16 * 128;
                const uint64_t a0 = make_desc_mn(sa + ko, p.lbo, p.sbo);
                const uint64_t a1 = make_desc_mn(sa + A_PL + ko, p.lbo, p.sbo);
                const uint64_t b0 = make_desc_mn(sa + 2 * A_PL + ko, p.lbo, p.sbo);
                const uint64_t b1 = make_desc_mn(sa + 2 * A_PL + B_PL + ko, p.lbo, p.sbo);
                umma_bf16(tmem_base, a0, b0, idesc, (it | kk) != 0);
                umma_bf16(tmem_base + NCOLS, a0, b1, idesc, (it | kk) != 0);
                umma_bf16(tmem_base + NCOLS, a1, b0, idesc, 1);
            }
            umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
    } else if (warp >= 2) {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int m = m0 + row;
        float* __restrict__ out = p.dw + (size_t)wg_split * p.split_stride + (size_t)wg_j * p.s_j;
        const size_t ma = (size_t)(m % p.msplit) * p.s_m + (size_t)(m / p.msplit) * p.s_mh;
        const bool vec = (p.s_n == 1) && ((p.Nw & 3) == 0) &&
                         (((ma + (size_t)wg_j * p.s_j + (size_t)wg_split * p.split_stride) & 3) == 0);
        for (int c32 = 0; c32 < NCOLS; c32 += 32) {
            float v[32];
            tmem_ld_add(taddr + c32, NCOLS, v);
            if (m >= p.Mw) continue;
            const int nn = n0 + c32;
            if (n_iters == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (vec && nn + 32 <= p.Nw) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(&out[ma + nn + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (nn + i < p.Nw) out[ma + (size_t)(nn + i) * p.s_n] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<2 * NCOLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

// bf16 3-D tensor map; box = (bk, 128, 1); swizzle chosen from the box width (64 or 128 bytes)
int encode_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || fn == nullptr) {
            set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
            return 1;
        }
        g_encode = (EncodeTiledFn)fn;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = box0 * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1);
        return 1;
    }
    return 0;
}

template <int MODE, int NBOX, int BK, int NPL, int CL = 1, int BR = 128>
static int launch_tc(const TcMaps& maps, const TcParams& p, dim3 grid, cudaStream_t st, const char* what) {
    using Cfg = TcCfg<NBOX, BK, NPL, BR>;
    static_assert(Cfg::STAGES >= 2, "pipeline needs at least two stages");
    static bool configured = false;
    auto kern = tc_conv_kernel<MODE, NBOX, BK, NPL, CL, BR>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) {
            set_error("%s: cannot set %d B dynamic smem: %s", what, Cfg::SMEM, cudaGetErrorString(e));
            return 1;
        }
        configured = true;
    }
    if (CL == 1) {
        kern<<<grid, TC_THREADS, Cfg::SMEM, st>>>(maps, p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = CL;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, p);
        if (e != cudaSuccess) { set_error("%s: cluster launch failed: %s", what, cudaGetErrorString(e)); return 1; }
    }
    return check_launch(what);
}

static int g_cl = 0;
static int tc_cluster() {                  // weight-multicast cluster size along the batch axis: DV3_TC_CLUSTER=1|2|4
    if (!g_cl) {
        const char* e = getenv("DV3_TC_CLUSTER");
        const int v = e ? atoi(e) : 1;
        g_cl = (v == 2 || v == 4) ? v : 1;
    }
    return g_cl;
}
static int pick_cluster(int B) {
    int cl = tc_cluster();
    while (cl > 1 && B % cl != 0) cl >>= 1;
    return cl;
}

static int g_bk = 0;
static int tc_bk() {                       // K-block width: 32 (SWIZZLE_64B, deeper pipeline) unless DV3_TC_BK=64
    if (!g_bk) {
        const char* e = getenv("DV3_TC_BK");
        g_bk = (e && atoi(e) == 64) ? 64 : 32;
    }
    return g_bk;
}

static void fill_taps_tc(int* tap_off, int k, int dilation, int causal, bool transpose) {
    const int padl = causal ? (k - 1) * dilation : (k - 1) / 2 * dilation;
    for (int j = 0; j < MAX_TAPS_TC; ++j)
        tap_off[j] = j < k ? (transpose ? (padl - j * dilation) : (j * dilation - padl)) : 0;
}

// plane p of a [nplanes][...] bf16 buffer
static inline const void* plane(const void* base, int pl, long long plane_elems) {
    return (const char*)base + (size_t)pl * plane_elems * 2;
}

}  // namespace dv3

using namespace dv3;

extern "C" {

int dv3_tc_k_block(void) { return tc_bk(); }

// 1 if the tensor-core ConvBlock path supports this block shape (else the caller uses the exact-fp32 kernels).
// (Only the legacy K-major weight gradient dv3_tc_wgrad needs T % 8 == 0; the default dv3_tc_wgrad_mn does not.)
int dv3_tc_supported(int B, int C, int T, int k) {
    return (C % 128 == 0) && T >= 1 && k >= 1 && k <= MAX_TAPS_TC && B >= 1 && B <= 65535;
}
// plain convs: any channel counts (planes are padded to a multiple of 8 channels), T % 8 == 0; k-tap convs need
// Cout % 128 == 0 so a weight box never straddles two taps
int dv3_tc_conv_supported(int B, int Cin, int Cout, int T, int k) {
    return T >= 1 && k >= 1 && k <= MAX_TAPS_TC && B >= 1 && B <= 65535 && (k == 1 || Cout % 128 == 0) && Cin >= 8 &&
           Cout >= 1;
}

// Gated forward.  xd: [npl][B][T][C] bf16 planes of the (dropped-out) input; w: [npl][k][2C][C] bf16 planes of the
// normalised weight; npl = 2; the rest as dv3_convblock_fwd.
//
// Kernel selection (measured, tools/tc_time.py / bench.py on the B200):
//   more 64-channel tiles than SMs  -> persistent kernel, 128-byte operand rows (BK = 64, SWIZZLE_128B)
//   128-wide tiles would leave SMs idle -> one 64-channel tile per CTA, BK = 64
//   otherwise                        -> one 128-channel tile per CTA (N = 256), BK = 32
// BK = 64 halves the number of TMA operations and mbarrier round trips per byte and makes every L2 request a full
// 128-byte line: 10-25 % faster than BK = 32 on every shape once the epilogue stopped being the bottleneck.
// Opt-in experiments (all parity-clean, none faster): DV3_TC_TAPS (tap reuse), DV3_TC_MCAST (weight multicast),
// DV3_TC_PAIR (cta_group::2), DV3_TC_CLUSTER (multicast in the one-tile kernels).
int dv3_tc_convblock_fwd(const void* xd, const void* w, int npl, const float* bias, const float* spk,
                         const float* res, float* y, float* save_a, float* save_s, int B, int C, int T, int k,
                         int dilation, int causal, int mode, int residual, void* stream) {
    DV3_REQUIRE(dv3_tc_supported(B, C, T, k), "tc_convblock_fwd: unsupported shape B=%d C=%d T=%d k=%d", B, C, T, k);
    DV3_REQUIRE(npl == 2, "tc_convblock_fwd: npl must be 2");
    TcMaps maps;
    const int cl = pick_cluster(B);
    const int t_tiles = (T + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    // tensor maps for a given K-block width and box heights (activation rows, weight rows)
    auto enc = [&](int bkx, int a_rows, int b_rows) -> int {
        for (int pl = 0; pl < 2; ++pl) {
            if (encode_tmap_bf16_3d(&maps.a[pl], plane(xd, pl, (long long)B * T * C), C, T, B, (uint64_t)C * 2,
                                    (uint64_t)T * C * 2, bkx, a_rows)) return 1;
            if (encode_tmap_bf16_3d(&maps.b[pl], plane(w, pl, (long long)k * 2 * C * C), C, (uint64_t)k * 2 * C, 1,
                                    (uint64_t)C * 2, (uint64_t)k * 2 * C * C * 2, bkx, b_rows)) return 1;
        }
        return 0;
    };
    TcParams p = {};
    p.T = T; p.B = B; p.Kc = C; p.Nc = C; p.rows_per_tap = 2 * C; p.k = k; p.kb_n = (C + 31) / 32;
    fill_taps_tc(p.tap_off, k, dilation, causal, false);
    p.bias = bias; p.spk = spk; p.res = res; p.y = y; p.save_a = save_a; p.save_s = save_s;
    p.gate_mode = mode; p.residual = residual;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DV3_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    const long long tiles64 = (long long)t_tiles * (C / 64) * B;

    if (cl == 1 && pair64_usable(B, C, tiles64)) {                     // opt-in, unvalidated: CTA pairs, BK = 64
        p.kb_n = C / 64;
        if (enc(64, 128, 64)) return 1;
        return launch_tc_pair64<TC_GATED>(maps, p, t_tiles, C / 64, B, st, "tc_convblock_fwd(pair64)");
    }
    if (cl == 1 && pair_usable(B, k, p.tap_off, tiles64)) {            // opt-in: CTA pairs (cta_group::2)
        const PairGeom g = pair_geom(k, p.tap_off);
        if (enc(32, g.a_rows, 64)) return 1;
        return launch_tc_pair<TC_GATED>(maps, p, g, t_tiles, C / 64, B, st, "tc_convblock_fwd(pair)");
    }
    if (cl == 1 && taps_usable<2, 64>(k, p.tap_off, tiles64)) {        // opt-in: tap reuse (+ weight multicast)
        const TapGeom g = tap_geom<2, 64>(k, p.tap_off);
        if (enc(32, g.a_rows, 64)) return 1;
        if (taps_mcast(B)) return launch_tc_taps<TC_GATED, 2, 64, 2>(maps, p, g, t_tiles, C / 64, B, st, "tc_convblock_fwd(taps,mcast)");
        return launch_tc_taps<TC_GATED, 2, 64>(maps, p, g, t_tiles, C / 64, B, st, "tc_convblock_fwd(taps)");
    }
    if (cl == 1 && tc_persist() && tiles64 > 148) {                    // persistent, double-buffered accumulators
        if (persist_bk() == 64) {
            p.kb_n = C / 64;
            if (enc(64, 128, 64)) return 1;
            return launch_tc_persist<TC_GATED, 2, 64, 64>(maps, p, t_tiles, C / 64, B, st, "tc_convblock_fwd(persistent)");
        }
        if (enc(32, 128, 64)) return 1;
        return launch_tc_persist<TC_GATED, 2, 64>(maps, p, t_tiles, C / 64, B, st, "tc_convblock_fwd(persistent,bk32)");
    }
    static int force_half = -1, no_narrow = -1;
    if (force_half < 0) { const char* e = getenv("DV3_TC_FORCE_HALF"); force_half = (e && atoi(e) == 1) ? 1 : 0; }
    if (no_narrow < 0) { const char* e = getenv("DV3_TC_NO_NARROW"); no_narrow = (e && atoi(e) == 1) ? 1 : 0; }
    // 64-channel tiles (64 a | 64 b columns) when 128-channel tiles would leave most of the 148 SMs idle
    const bool half = cl == 1 && !no_narrow && (force_half || (long long)t_tiles * (C / 128) * B < 100);
    if (half) {
        dim3 grid(t_tiles, C / 64, B);
        if (small_bk() == 64) {
            p.kb_n = C / 64;
            if (enc(64, 128, 64)) return 1;
            return launch_tc<TC_GATED, 2, 64, 2, 1, 64>(maps, p, grid, st, "tc_convblock_fwd(64)");
        }
        if (enc(32, 128, 64)) return 1;
        return launch_tc<TC_GATED, 2, 32, 2, 1, 64>(maps, p, grid, st, "tc_convblock_fwd(64,bk32)");
    }
    const int bk = tc_bk();
    p.kb_n = (C + bk - 1) / bk;
    if (enc(bk, 128, 128)) return 1;
    if (cl > 1) {
        for (int pl = 0; pl < 2; ++pl)
            if (encode_tmap_bf16_3d(&maps.bs[pl], plane(w, pl, (long long)k * 2 * C * C), C, (uint64_t)k * 2 * C, 1,
                                    (uint64_t)C * 2, (uint64_t)k * 2 * C * C * 2, bk, 128 / cl)) return 1;
    }
    dim3 grid(t_tiles, C / 128, B);
    if (bk == 64) return launch_tc<TC_GATED, 2, 64, 2>(maps, p, grid, st, "tc_convblock_fwd(128,bk64)");
    if (cl == 4) return launch_tc<TC_GATED, 2, 32, 2, 4>(maps, p, grid, st, "tc_convblock_fwd(cluster4)");
    if (cl == 2) return launch_tc<TC_GATED, 2, 32, 2, 2>(maps, p, grid, st, "tc_convblock_fwd(cluster2)");
    return launch_tc<TC_GATED, 2, 32, 2>(maps, p, grid, st, "tc_convblock_fwd(128)");
}

// Generic conv / data-gradient:  out (B, Nc, T) fp32 = sum_j A[b, t+off_j, :] . W[j, n, :]  (+ epilogue)
//   a: [npl][B][T][Kp] bf16 planes, Kp = Kc rounded up to 8;  w: [npl][k][Nc][Kp] bf16 planes; npl = 2.
//   transpose_taps = 1 for a data gradient (offsets padl - j*d), 0 for a forward conv.
// Same kernel selection as dv3_tc_convblock_fwd; BK = 64 needs Kc % 64 == 0 (else BK = 32: the 80-channel mel input).
int dv3_tc_conv(const void* a, const void* w, int npl, float* out, int B, int Kc, int Nc, int T, int k, int dilation,
                int causal, int transpose_taps, const float* bias, int relu, float p_drop,
                const unsigned long long* seed_ptr, unsigned salt, int addmode, const float* e1, const float* e2,
                float alpha, void* stream) {
    DV3_REQUIRE(k >= 1 && k <= MAX_TAPS_TC && (k == 1 || Nc % 128 == 0) && B <= 65535,
                "tc_conv: unsupported shape B=%d Kc=%d Nc=%d T=%d k=%d", B, Kc, Nc, T, k);
    DV3_REQUIRE(npl == 2, "tc_conv: npl must be 2");
    const int Kp = (Kc + 7) / 8 * 8;
    TcMaps maps;
    const int cl = pick_cluster(B);
    const int t_tiles = (T + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    auto enc = [&](int bkx, int a_rows, int b_rows) -> int {
        for (int pl = 0; pl < 2; ++pl) {
            if (encode_tmap_bf16_3d(&maps.a[pl], plane(a, pl, (long long)B * T * Kp), Kc, T, B, (uint64_t)Kp * 2,
                                    (uint64_t)T * Kp * 2, bkx, a_rows)) return 1;
            if (encode_tmap_bf16_3d(&maps.b[pl], plane(w, pl, (long long)k * Nc * Kp), Kc, (uint64_t)k * Nc, 1,
                                    (uint64_t)Kp * 2, (uint64_t)k * Nc * Kp * 2, bkx, b_rows)) return 1;
        }
        return 0;
    };
    TcParams p = {};
    p.T = T; p.B = B; p.Kc = Kc; p.Nc = Nc; p.rows_per_tap = Nc; p.k = k; p.kb_n = (Kc + 31) / 32;
    fill_taps_tc(p.tap_off, k, dilation, causal, transpose_taps != 0);
    p.out = out; p.bias = bias; p.relu = relu; p.e1 = e1; p.e2 = e2; p.alpha = alpha; p.addmode = addmode;
    p.p_drop = p_drop; p.seed_ptr = seed_ptr; p.salt = salt;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DV3_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    static int no_narrow = -1;
    if (no_narrow < 0) { const char* e = getenv("DV3_TC_NO_NARROW"); no_narrow = (e && atoi(e) == 1) ? 1 : 0; }
    const long long tiles128 = (long long)t_tiles * ((Nc + 127) / 128) * B;
    // tile width: two 128-column boxes only when that still fills the machine; 64 columns for small problems
    const bool wide = Nc > 128 && (long long)t_tiles * ((Nc + 255) / 256) * B >= 120;
    const bool narrow = !wide && !no_narrow && cl == 1 && Nc > 64 && (k == 1 || Nc % 64 == 0) && tiles128 < 100;
    const bool k64 = Kc % 64 == 0;

    if (cl == 1 && Nc % 128 == 0 && pair64_usable(B, Kc, tiles128)) {               // opt-in, unvalidated
        p.kb_n = Kc / 64;
        if (enc(64, 128, 64)) return 1;
        return launch_tc_pair64<TC_CONV>(maps, p, t_tiles, Nc / 128, B, st, "tc_conv(pair64)");
    }
    if (cl == 1 && Nc % 128 == 0 && pair_usable(B, k, p.tap_off, tiles128)) {       // opt-in: CTA pairs
        const PairGeom g = pair_geom(k, p.tap_off);
        if (enc(32, g.a_rows, 64)) return 1;
        return launch_tc_pair<TC_CONV>(maps, p, g, t_tiles, Nc / 128, B, st, "tc_conv(pair)");
    }
    if (cl == 1 && k > 1) {                                                          // opt-in: tap reuse
        const bool n64 = narrow && !(tiles128 > 148);
        const bool ok = n64 ? taps_usable<1, 64>(k, p.tap_off, (long long)t_tiles * (Nc / 64) * B)
                            : taps_usable<1, 128>(k, p.tap_off, tiles128);
        if (ok) {
            const TapGeom g = n64 ? tap_geom<1, 64>(k, p.tap_off) : tap_geom<1, 128>(k, p.tap_off);
            if (enc(32, g.a_rows, n64 ? 64 : 128)) return 1;
            if (n64) return launch_tc_taps<TC_CONV, 1, 64>(maps, p, g, t_tiles, Nc / 64, B, st, "tc_conv(taps64)");
            if (taps_mcast(B)) return launch_tc_taps<TC_CONV, 1, 128, 2>(maps, p, g, t_tiles, Nc / 128, B, st, "tc_conv(taps,mcast)");
            return launch_tc_taps<TC_CONV, 1, 128>(maps, p, g, t_tiles, Nc / 128, B, st, "tc_conv(taps)");
        }
    }
    if (cl == 1 && tc_persist() && tiles128 > 148) {
        // more 128-column tiles than SMs: persistent kernel with double-buffered accumulators
        if (persist_bk() == 64 && k64) {
            p.kb_n = Kc / 64;
            if (enc(64, 128, 128)) return 1;
            return launch_tc_persist<TC_CONV, 1, 128, 64>(maps, p, t_tiles, (Nc + 127) / 128, B, st, "tc_conv(persistent)");
        }
        if (enc(32, 128, 128)) return 1;
        return launch_tc_persist<TC_CONV, 1, 128>(maps, p, t_tiles, (Nc + 127) / 128, B, st, "tc_conv(persistent,bk32)");
    }
    if (narrow) {
        dim3 grid(t_tiles, (Nc + 63) / 64, B);
        if (small_bk() == 64 && k64) {
            p.kb_n = Kc / 64;
            if (enc(64, 128, 64)) return 1;
            return launch_tc<TC_CONV, 1, 64, 2, 1, 64>(maps, p, grid, st, "tc_conv(64)");
        }
        if (enc(32, 128, 64)) return 1;
        return launch_tc<TC_CONV, 1, 32, 2, 1, 64>(maps, p, grid, st, "tc_conv(64,bk32)");
    }
    if (wide) {
        const int bk = tc_bk();
        p.kb_n = (Kc + bk - 1) / bk;
        if (enc(bk, 128, 128)) return 1;
        if (cl > 1) {
            for (int pl = 0; pl < 2; ++pl)
                if (encode_tmap_bf16_3d(&maps.bs[pl], plane(w, pl, (long long)k * Nc * Kp), Kc, (uint64_t)k * Nc, 1,
                                        (uint64_t)Kp * 2, (uint64_t)k * Nc * Kp * 2, bk, 128 / cl)) return 1;
        }
        dim3 grid(t_tiles, (Nc + 255) / 256, B);
        if (bk == 64) return launch_tc<TC_CONV, 2, 64, 2>(maps, p, grid, st, "tc_conv(256,bk64)");
        if (cl == 4) return launch_tc<TC_CONV, 2, 32, 2, 4>(maps, p, grid, st, "tc_conv(cluster4)");
        if (cl == 2) return launch_tc<TC_CONV, 2, 32, 2, 2>(maps, p, grid, st, "tc_conv(cluster2)");
        return launch_tc<TC_CONV, 2, 32, 2>(maps, p, grid, st, "tc_conv(256)");
    }
    // one 128-column tile per CTA
    const int bk = (cl == 1 && small_bk() == 64 && k64) ? 64 : tc_bk();
    p.kb_n = (Kc + bk - 1) / bk;
    if (enc(bk, 128, 128)) return 1;
    if (cl > 1) {
        for (int pl = 0; pl < 2; ++pl)
            if (encode_tmap_bf16_3d(&maps.bs[pl], plane(w, pl, (long long)k * Nc * Kp), Kc, (uint64_t)k * Nc, 1,
                                    (uint64_t)Kp * 2, (uint64_t)k * Nc * Kp * 2, bk, 128 / cl)) return 1;
    }
    dim3 grid(t_tiles, (Nc + 127) / 128, B);
    if (bk == 64) return launch_tc<TC_CONV, 1, 64, 2>(maps, p, grid, st, "tc_conv(128)");
    if (cl == 4) return launch_tc<TC_CONV, 1, 32, 2, 4>(maps, p, grid, st, "tc_conv(cluster4)");
    if (cl == 2) return launch_tc<TC_CONV, 1, 32, 2, 2>(maps, p, grid, st, "tc_conv(cluster2)");
    return launch_tc<TC_CONV, 1, 32, 2>(maps, p, grid, st, "tc_conv(128,bk32)");
}

int dv3_tc_wgrad_nsplit(int B, int Mw, int Nw, int T, int k) {
    const int nt = Nw > 128 ? (Nw + 255) / 256 : 1;
    const int tiles = ((Mw + 127) / 128) * nt * k;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("DV3_TC_WGRAD_CTAS"); forced = e ? atoi(e) : 0; }
    // split (b,t) over enough CTAs for two waves, unless that leaves each CTA fewer than ~64 K-iterations (then the
    // per-CTA prologue/epilogue and the extra partial traffic cost more than the parallelism buys: one wave).
    // Measured on the preset shapes (tools/tc_time.py): (512,800) prefers 296, (256,800)/(512,128)/(256,200) prefer 148.
    const int kb_n = (T + 31) / 32;
    int best = 1;
    for (int target = forced > 0 ? forced : 2 * 148; target >= 148; target -= 148) {
        int want = (target + tiles - 1) / tiles;
        if (want > B) want = B;
        if (want < 1) want = 1;
        const int bps = (B + want - 1) / want;
        best = (B + bps - 1) / bps;
        if (forced > 0 || bps * kb_n >= 64) break;
    }
    return best;
}

// Weight gradient.  dy: [2][B][Mw][T] bf16 planes; xs: [2][k][B][Nw][T] bf16 planes (k time-shifted copies of the
// conv input, dv3_tc_split_input); partial element (m, n, j) at (m%msplit)*s_m + (m/msplit)*s_mh + n*s_n + j*s_j.
int dv3_tc_wgrad(const void* dy, const void* xs, float* dw_partials, long long split_stride, int B, int Mw, int Nw,
                 int T, int k, int msplit, long long s_m, long long s_mh, long long s_n, long long s_j,
                 void* stream) {
    DV3_REQUIRE(T % 8 == 0 && k >= 1 && k <= MAX_TAPS_TC && B <= 65535, "tc_wgrad: unsupported shape T=%d k=%d", T, k);
    const int bk = tc_bk();
    TcMaps maps;
    for (int pl = 0; pl < 2; ++pl) {
        if (encode_tmap_bf16_3d(&maps.a[pl], plane(dy, pl, (long long)B * Mw * T), T, Mw, B, (uint64_t)T * 2,
                                (uint64_t)Mw * T * 2, bk, 128)) return 1;
        if (encode_tmap_bf16_3d(&maps.b[pl], plane(xs, pl, (long long)k * B * Nw * T), T, Nw, (uint64_t)k * B,
                                (uint64_t)T * 2, (uint64_t)Nw * T * 2, bk, 128)) return 1;
    }
    TcParams p = {};
    p.T = T; p.B = B; p.Mw = Mw; p.Nc = Nw; p.k = k; p.kb_n = (T + bk - 1) / bk;
    p.nsplit = dv3_tc_wgrad_nsplit(B, Mw, Nw, T, k);
    p.batches_per_split = (B + p.nsplit - 1) / p.nsplit;
    p.dw = dw_partials; p.split_stride = split_stride;
    p.msplit = msplit; p.s_m = s_m; p.s_mh = s_mh; p.s_n = s_n; p.s_j = s_j;
    cudaStream_t st = (cudaStream_t)stream;
    const int m_tiles = (Mw + 127) / 128;
    if (Nw > 128) {
        dim3 grid((Nw + 255) / 256, m_tiles, p.nsplit * k);
        if (bk == 64) return launch_tc<TC_WGRAD, 2, 64, 2>(maps, p, grid, st, "tc_wgrad");
        return launch_tc<TC_WGRAD, 2, 32, 2>(maps, p, grid, st, "tc_wgrad");
    }
    dim3 grid(1, m_tiles, p.nsplit * k);
    if (bk == 64) return launch_tc<TC_WGRAD, 1, 64, 2>(maps, p, grid, st, "tc_wgrad");
    return launch_tc<TC_WGRAD, 1, 32, 2>(maps, p, grid, st, "tc_wgrad");
}

// Weight gradient from (B,T,C) planes (no shifted copies).  dy: [2][B][T][pad8(Mw)], xd: [2][B][T][pad8(Nw)].
int dv3_tc_wgrad_mn(const void* dy, const void* xd, float* dw_partials, long long split_stride, int B, int Mw,
                    int Nw, int T, int k, int dilation, int causal, int msplit, long long s_m, long long s_mh,
                    long long s_n, long long s_j, void* stream) {
    DV3_REQUIRE(k >= 1 && k <= MAX_TAPS_TC && B <= 65535, "tc_wgrad_mn: unsupported shape k=%d", k);
    const int Mp = (Mw + 7) / 8 * 8, Np = (Nw + 7) / 8 * 8;
    TcMaps maps;
    for (int pl = 0; pl < 2; ++pl) {
        if (encode_tmap_bf16_3d(&maps.a[pl], plane(dy, pl, (long long)B * T * Mp), Mw, T, B, (uint64_t)Mp * 2,
                                (uint64_t)T * Mp * 2, 64, 32)) return 1;
        if (encode_tmap_bf16_3d(&maps.b[pl], plane(xd, pl, (long long)B * T * Np), Nw, T, B, (uint64_t)Np * 2,
                                (uint64_t)T * Np * 2, 64, 32)) return 1;
    }
    TcMnParams p = {};
    p.T = T; p.B = B; p.Mw = Mw; p.Nw = Nw; p.k = k; p.kb_n = (T + 31) / 32;
    fill_taps_tc(p.tap_off, k, dilation, causal, false);
    p.nsplit = dv3_tc_wgrad_nsplit(B, Mw, Nw, T, k);
    p.batches_per_split = (B + p.nsplit - 1) / p.nsplit;
    p.dw = dw_partials; p.split_stride = split_stride;
    p.msplit = msplit; p.s_m = s_m; p.s_mh = s_mh; p.s_n = s_n; p.s_j = s_j;
    // 64-channel chunks are 4 KB apart (one TMA box each), 8-row groups 1 KB apart inside a box
    const char* sw = getenv("DV3_TC_MN_SWAP");
    p.lbo = 4096; p.sbo = 1024;
    if (sw && atoi(sw) == 1) { p.lbo = 1024; p.sbo = 4096; }
    cudaStream_t st = (cudaStream_t)stream;
    const int m_tiles = (Mw + 127) / 128;
    if (Nw > 128) {
        constexpr int SMEM = 4 * 2 * (2 * 4096 + 4 * 4096) + 1024 + 512;
        static bool configured = false;
        if (!configured) { cudaFuncSetAttribute(tc_wgrad_mn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); configured = true; }
        tc_wgrad_mn_kernel<2><<<dim3((Nw + 255) / 256, m_tiles, p.nsplit * k), TC_THREADS, SMEM, st>>>(maps, p);
    } else {
        constexpr int SMEM = 6 * 2 * (2 * 4096 + 2 * 4096) + 1024 + 512;
        static bool configured = false;
        if (!configured) { cudaFuncSetAttribute(tc_wgrad_mn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); configured = true; }
        tc_wgrad_mn_kernel<1><<<dim3(1, m_tiles, p.nsplit * k), TC_THREADS, SMEM, st>>>(maps, p);
    }
    return check_launch("tc_wgrad_mn");
}

}  // extern "C"
