// tcgen05 tensor-core path of the ConvBlock: forward, data-gradient and weight-gradient as implicit GEMMs with
// fp32-equivalent accuracy ("bf16x3"): every fp32 operand is pre-split into bf16 hi + lo planes (by the
// elementwise kernels in tc_split.cu, which also apply the input dropout and emit the K-major layout each GEMM
// wants), and each K-step issues hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator (relative error ~2^-16 per
// product, 50x inside the rtol=1e-3 parity bar; plain TF32 would not be).
//
//   forward : D[t, (a|b) c]  = sum_{j,ci}  Xd[b, t+off_j, ci] * W[j, (a|b) c, ci]     M = 128 time steps, N = 2 x 128
//   dgrad   : D[t, ci]       = sum_{j,co}  dAB[b, t-off_j, co] * W[j, ci, co]          M = 128 time steps, N = NBOX x 128
//   wgrad   : D[co, ci] (j)  = sum_{b,t}   dAB[b, co, t] * Xd[b, ci, t+off_j]          M = 128 rows,       N = NBOX x 128
//
// All operands are K-major bf16 tiles of 128 rows x 64 (128 bytes, SWIZZLE_128B) fetched by TMA; the conv's zero
// padding, the causal shift and ragged tails are TMA out-of-bounds zero fill (negative / >= T coordinates).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer, warps 2-5 =
// epilogue (TMEM -> registers -> fused gate / mask / residual math -> coalesced global stores along T).
#include "tc_common.cuh"

namespace dv3 {

using namespace tc;

constexpr int TC_THREADS = 192;
constexpr int TILE_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr int MAX_TAPS_TC = 8;

enum { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2 };

struct TcParams {
    int T, C, M2;              // time steps, block channels, 2*C
    int k, kb_n;               // taps, 64-wide K blocks per tap (fwd: C/64, dgrad: 2C/64); wgrad: t-chunks per batch
    int tap_off[MAX_TAPS_TC];
    // forward epilogue
    const float* bias; const float* spk; const float* res;
    float* y; float* save_a; float* save_s;
    int gate_mode, residual;
    // dgrad epilogue
    float* dx; const float* e1; const float* e2; float alpha; int addmode;
    float p_drop; const unsigned long long* seed_ptr; uint32_t salt;
    // wgrad
    float* dw; long long split_stride; int nsplit, B, batches_per_split, Cin;
};

template <int MODE, int NBOX>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
               const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
               const __grid_constant__ TcParams p) {
    constexpr int STAGE_BYTES = (2 + 2 * NBOX) * TILE_BYTES;
    constexpr int STAGES = NBOX == 2 ? 2 : 3;
    constexpr int NCOLS = 128 * NBOX;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates ----------------------------------------------------------------------
    int a_row0, a_z, b_row0, b_row1, n_iters, it_base = 0, wg_j = 0, wg_split = 0, b_beg = 0;
    if (MODE == TC_FWD) {
        a_row0 = blockIdx.x * 128;                       // t0
        a_z = blockIdx.z;                                // batch
        b_row0 = blockIdx.y * 128;                       // c0 (a half); b half at + C
        b_row1 = p.C + blockIdx.y * 128;
        n_iters = p.k * p.kb_n;
    } else if (MODE == TC_DGRAD) {
        a_row0 = blockIdx.x * 128;
        a_z = blockIdx.z;
        b_row0 = blockIdx.y * 128 * NBOX;                // ci0
        b_row1 = b_row0 + 128;
        n_iters = p.k * p.kb_n;
    } else {
        wg_j = blockIdx.z % p.k; wg_split = blockIdx.z / p.k;
        a_row0 = blockIdx.y * 128;                       // co0
        b_row0 = blockIdx.x * 128 * NBOX;                // ci0
        b_row1 = b_row0 + 128;
        b_beg = wg_split * p.batches_per_split;
        int b_end = b_beg + p.batches_per_split; if (b_end > p.B) b_end = p.B;
        n_iters = (b_end > b_beg ? b_end - b_beg : 0) * p.kb_n;
        a_z = 0;
    }
    (void)it_base;

    if (threadIdx.x == 0) {
        prefetch_tmap(&mapA_hi); prefetch_tmap(&mapA_lo); prefetch_tmap(&mapB_hi); prefetch_tmap(&mapB_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<NCOLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            int ax, ay, az, bx, by0, by1, bz;
            if (MODE == TC_FWD || MODE == TC_DGRAD) {
                const int j = it / p.kb_n, kb = it - j * p.kb_n;
                ax = kb * 64; ay = a_row0 + p.tap_off[j]; az = a_z;
                const int rows_per_tap = (MODE == TC_FWD) ? p.M2 : p.C;
                bx = kb * 64; by0 = j * rows_per_tap + b_row0; by1 = j * rows_per_tap + b_row1; bz = 0;
            } else {
                const int bi = it / p.kb_n, tc_ = it - bi * p.kb_n;
                ax = tc_ * 64; ay = a_row0; az = b_beg + bi;
                // the tap shift is baked into the wg_j-th shifted copy of the input (tc_split.cu): a TMA box
                // cannot start at a K (time) coordinate that is not 16-byte aligned
                bx = tc_ * 64; by0 = b_row0; by1 = b_row1; bz = wg_j * p.B + b_beg + bi;
            }
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            tma_load_3d(st, &mapA_hi, &full[s], ax, ay, az);
            tma_load_3d(st + TILE_BYTES, &mapA_lo, &full[s], ax, ay, az);
            tma_load_3d(st + 2 * TILE_BYTES, &mapB_hi, &full[s], bx, by0, bz);
            if (NBOX == 2) tma_load_3d(st + 3 * TILE_BYTES, &mapB_hi, &full[s], bx, by1, bz);
            tma_load_3d(st + (2 + NBOX) * TILE_BYTES, &mapB_lo, &full[s], bx, by0, bz);
            if (NBOX == 2) tma_load_3d(st + (3 + NBOX) * TILE_BYTES, &mapB_lo, &full[s], bx, by1, bz);
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOLS);
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint64_t a_hi = make_smem_desc_sw128(sa), a_lo = make_smem_desc_sw128(sa + TILE_BYTES);
            const uint64_t b_hi = make_smem_desc_sw128(sa + 2 * TILE_BYTES);
            const uint64_t b_lo = make_smem_desc_sw128(sa + (2 + NBOX) * TILE_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {               // 4 x UMMA_K(16) = 64
                const uint64_t adv = (uint64_t)(kk * 2);   // 32 bytes >> 4
                umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (it | kk) != 0);
                umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
            }
            umma_commit(&empty[s]);                        // frees the stage once these MMAs retire
        }
        umma_commit(tmem_full);
    } else if (warp >= 2) {
        // ================= epilogue =================
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;                            // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;                     // accumulator row (M index)
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (n_iters == 0) {
            // nothing accumulated (only possible for an empty wgrad split): treat as zeros
        }
        if (MODE == TC_FWD) {
            const int t = a_row0 + row, b = a_z;
            const bool tv = t < p.T;
            for (int c32 = 0; c32 < 128; c32 += 32) {
                float va[32], vb[32];
                tmem_ld_32x32(taddr + c32, va);
                tmem_ld_32x32(taddr + 128 + c32, vb);
                if (!tv) continue;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = b_row0 + c32 + i;
                    const size_t idx = ((size_t)b * p.C + c) * p.T + t;
                    float a = va[i] + p.bias[c];
                    if (p.spk) a += p.spk[idx];
                    const float s = sigmoidf_(vb[i] + p.bias[p.C + c]);
                    float y;
                    if (p.gate_mode == 0) {
                        y = a * s;
                        if (p.residual) y = (y + p.res[idx]) * 0.70710678118654752f;
                    } else {
                        y = s * a + (1.f - s) * p.res[idx];
                    }
                    p.y[idx] = y;
                    if (p.save_a) p.save_a[idx] = a;
                    if (p.save_s) p.save_s[idx] = s;
                }
            }
        } else if (MODE == TC_DGRAD) {
            const int t = a_row0 + row, b = a_z;
            const bool tv = t < p.T;
            const DropCfg drop = make_drop(p.p_drop, p.seed_ptr, p.salt);
            for (int c32 = 0; c32 < NCOLS; c32 += 32) {
                float v[32];
                tmem_ld_32x32(taddr + c32, v);
                if (!tv) continue;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int ci = b_row0 + c32 + i;
                    if (ci >= p.C) continue;
                    const size_t idx = ((size_t)b * p.C + ci) * p.T + t;
                    float g = v[i] * drop_scale(drop, (uint32_t)idx);
                    if (p.addmode == 1) g += p.alpha * p.e1[idx];
                    else if (p.addmode == 2) g += p.e1[idx] * (1.f - p.e2[idx]);
                    p.dx[idx] = g;
                }
            }
        } else {
            const int co = a_row0 + row;
            float* out = p.dw + (size_t)wg_split * p.split_stride + wg_j;
            for (int c32 = 0; c32 < NCOLS; c32 += 32) {
                float v[32];
                tmem_ld_32x32(taddr + c32, v);
                if (co >= p.M2) continue;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int ci = b_row0 + c32 + i;
                    if (ci < p.Cin) out[((size_t)co * p.Cin + ci) * p.k] = (n_iters > 0) ? v[i] : 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<NCOLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

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
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1);
        return 1;
    }
    return 0;
}

template <int MODE, int NBOX>
static int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                     const CUtensorMap& b_lo, const TcParams& p, dim3 grid, cudaStream_t st, const char* what) {
    constexpr int STAGES = NBOX == 2 ? 2 : 3;
    constexpr int SMEM = STAGES * (2 + 2 * NBOX) * TILE_BYTES + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<MODE, NBOX>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("%s: cannot set %d B dynamic smem: %s", what, SMEM, cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    tc_conv_kernel<MODE, NBOX><<<grid, TC_THREADS, SMEM, st>>>(a_hi, a_lo, b_hi, b_lo, p);
    return check_launch(what);
}

static void fill_taps_tc(int* tap_off, int k, int dilation, int causal, bool transpose) {
    const int padl = causal ? (k - 1) * dilation : (k - 1) / 2 * dilation;
    for (int j = 0; j < k; ++j) tap_off[j] = transpose ? (padl - j * dilation) : (j * dilation - padl);
}

}  // namespace dv3

using namespace dv3;
typedef __nv_bfloat16 bf16;

extern "C" {

// 1 if the tensor-core path supports this block shape (else the caller must use the exact-fp32 kernels)
int dv3_tc_supported(int B, int C, int T, int k) {
    return (C % 128 == 0) && (T % 8 == 0) && k >= 1 && k <= MAX_TAPS_TC && B >= 1 && B <= 65535;
}

// forward.  xd_hi/xd_lo: (B, T, C) bf16 planes of the (dropped-out) input; w_hi/w_lo: [k][2C][C] bf16 planes of
// the normalised weight; everything else as dv3_convblock_fwd.
int dv3_tc_convblock_fwd(const void* xd_hi, const void* xd_lo, const void* w_hi, const void* w_lo,
                         const float* bias, const float* spk, const float* res, float* y, float* save_a,
                         float* save_s, int B, int C, int T, int k, int dilation, int causal, int mode,
                         int residual, void* stream) {
    DV3_REQUIRE(dv3_tc_supported(B, C, T, k), "tc_convblock_fwd: unsupported shape B=%d C=%d T=%d k=%d", B, C, T, k);
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    if (encode_tmap_bf16_3d(&a_hi, xd_hi, C, T, B, (uint64_t)C * 2, (uint64_t)T * C * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&a_lo, xd_lo, C, T, B, (uint64_t)C * 2, (uint64_t)T * C * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&b_hi, w_hi, C, (uint64_t)k * 2 * C, 1, (uint64_t)C * 2, (uint64_t)k * 2 * C * C * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&b_lo, w_lo, C, (uint64_t)k * 2 * C, 1, (uint64_t)C * 2, (uint64_t)k * 2 * C * C * 2, 64, 128)) return 1;
    TcParams p = {};
    p.T = T; p.C = C; p.M2 = 2 * C; p.k = k; p.kb_n = C / 64;
    fill_taps_tc(p.tap_off, k, dilation, causal, false);
    p.bias = bias; p.spk = spk; p.res = res; p.y = y; p.save_a = save_a; p.save_s = save_s;
    p.gate_mode = mode; p.residual = residual;
    dim3 grid((T + 127) / 128, C / 128, B);
    return launch_tc<TC_FWD, 2>(a_hi, a_lo, b_hi, b_lo, p, grid, (cudaStream_t)stream, "tc_convblock_fwd");
}

// data gradient.  dab_hi/lo: (B, T, 2C) bf16 planes; w_hi/lo: [k][C][2C] bf16 planes; dx (B, C, T) fp32.
int dv3_tc_conv_dgrad(const void* dab_hi, const void* dab_lo, const void* w_hi, const void* w_lo, float* dx,
                      int B, int C, int T, int k, int dilation, int causal, float p_drop,
                      const unsigned long long* seed_ptr, unsigned salt, int addmode, const float* e1,
                      const float* e2, float alpha, void* stream) {
    DV3_REQUIRE(dv3_tc_supported(B, C, T, k), "tc_conv_dgrad: unsupported shape B=%d C=%d T=%d k=%d", B, C, T, k);
    const int M2 = 2 * C;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    if (encode_tmap_bf16_3d(&a_hi, dab_hi, M2, T, B, (uint64_t)M2 * 2, (uint64_t)T * M2 * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&a_lo, dab_lo, M2, T, B, (uint64_t)M2 * 2, (uint64_t)T * M2 * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&b_hi, w_hi, M2, (uint64_t)k * C, 1, (uint64_t)M2 * 2, (uint64_t)k * C * M2 * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&b_lo, w_lo, M2, (uint64_t)k * C, 1, (uint64_t)M2 * 2, (uint64_t)k * C * M2 * 2, 64, 128)) return 1;
    TcParams p = {};
    p.T = T; p.C = C; p.M2 = M2; p.k = k; p.kb_n = M2 / 64;
    fill_taps_tc(p.tap_off, k, dilation, causal, true);
    p.dx = dx; p.e1 = e1; p.e2 = e2; p.alpha = alpha; p.addmode = addmode;
    p.p_drop = p_drop; p.seed_ptr = seed_ptr; p.salt = salt;
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 256 == 0)
        return launch_tc<TC_DGRAD, 2>(a_hi, a_lo, b_hi, b_lo, p, dim3((T + 127) / 128, C / 256, B), st, "tc_conv_dgrad");
    return launch_tc<TC_DGRAD, 1>(a_hi, a_lo, b_hi, b_lo, p, dim3((T + 127) / 128, C / 128, B), st, "tc_conv_dgrad");
}

int dv3_tc_conv_wgrad_nsplit(int B, int C, int T, int k) {
    const int tiles = (2 * C / 128) * ((C % 256 == 0) ? C / 256 : C / 128) * k;
    int want = (2 * 148 + tiles - 1) / tiles;
    if (want > B) want = B;
    if (want < 1) want = 1;
    const int bps = (B + want - 1) / want;
    return (B + bps - 1) / bps;
}

// weight gradient.  dab_hi/lo: (B, 2C, T) bf16 planes; xd_hi/lo: (k, B, C, T) bf16 planes, the j-th being the
// dropped-out input shifted by tap j's offset (dv3_tc_split_input);
// dw_partials: [nsplit][2C*C*k] fp32 in v's layout (2C, C, k).
int dv3_tc_conv_wgrad(const void* dab_hi, const void* dab_lo, const void* xd_hi, const void* xd_lo,
                      float* dw_partials, long long split_stride, int B, int C, int T, int k, int dilation,
                      int causal, void* stream) {
    DV3_REQUIRE(dv3_tc_supported(B, C, T, k), "tc_conv_wgrad: unsupported shape B=%d C=%d T=%d k=%d", B, C, T, k);
    const int M2 = 2 * C;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    if (encode_tmap_bf16_3d(&a_hi, dab_hi, T, M2, B, (uint64_t)T * 2, (uint64_t)M2 * T * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&a_lo, dab_lo, T, M2, B, (uint64_t)T * 2, (uint64_t)M2 * T * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&b_hi, xd_hi, T, C, (uint64_t)k * B, (uint64_t)T * 2, (uint64_t)C * T * 2, 64, 128)) return 1;
    if (encode_tmap_bf16_3d(&b_lo, xd_lo, T, C, (uint64_t)k * B, (uint64_t)T * 2, (uint64_t)C * T * 2, 64, 128)) return 1;
    TcParams p = {};
    p.T = T; p.C = C; p.M2 = M2; p.Cin = C; p.k = k; p.kb_n = (T + 63) / 64; p.B = B;
    fill_taps_tc(p.tap_off, k, dilation, causal, false);
    p.nsplit = dv3_tc_conv_wgrad_nsplit(B, C, T, k);
    p.batches_per_split = (B + p.nsplit - 1) / p.nsplit;
    p.dw = dw_partials; p.split_stride = split_stride;
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 256 == 0)
        return launch_tc<TC_WGRAD, 2>(a_hi, a_lo, b_hi, b_lo, p, dim3(C / 256, M2 / 128, p.nsplit * k), st, "tc_conv_wgrad");
    return launch_tc<TC_WGRAD, 1>(a_hi, a_lo, b_hi, b_lo, p, dim3(C / 128, M2 / 128, p.nsplit * k), st, "tc_conv_wgrad");
}

}  // extern "C"
