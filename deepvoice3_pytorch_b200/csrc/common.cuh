// Shared device/host helpers for the dv3b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dv3 {

// ---- error plumbing (no exceptions across the C ABI) -------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);     // counts the launch; cudaGetLastError() -> 0 / error code + message
unsigned long long launch_count();

#define DV3_REQUIRE(cond, ...)                                    \
    do {                                                          \
        if (!(cond)) { ::dv3::set_error(__VA_ARGS__); return 1; } \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- counter-based dropout mask ----------------------------------------------------------------
// keep(idx) is a pure function of (step seed in device memory, call-site salt, element index), so the
// backward pass and the weight-gradient pass regenerate exactly the mask the forward used, and a
// replayed CUDA graph gets a fresh mask by bumping the 8-byte seed in device memory.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
struct DropCfg {
    uint32_t s0, s1;      // derived per launch from *seed_ptr and salt
    uint32_t thresh;      // drop iff hash < thresh ; thresh = p * 2^32
    float scale;          // 1/(1-p)
    int on;
};
__device__ __forceinline__ DropCfg make_drop(float p, const unsigned long long* seed_ptr, uint32_t salt) {
    DropCfg d;
    d.on = (p > 0.f) && (seed_ptr != nullptr);
    if (d.on) {
        unsigned long long s = *seed_ptr;
        d.s0 = mix32((uint32_t)s ^ (salt * 0x9E3779B1U));
        d.s1 = mix32((uint32_t)(s >> 32) + salt + 0x85ebca6bU);
        double t = (double)p * 4294967296.0;
        d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFU : (uint32_t)t;
        d.scale = 1.f / (1.f - p);
    } else { d.s0 = d.s1 = d.thresh = 0; d.scale = 1.f; }
    return d;
}
__device__ __forceinline__ float drop_scale(const DropCfg& d, uint32_t idx) {
    // returns 0 (dropped) or 1/(1-p) (kept); 1 when dropout is off
    if (!d.on) return 1.f;
    uint32_t h = mix32(mix32(idx ^ d.s0) + d.s1);
    return h < d.thresh ? 0.f : d.scale;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dv3
