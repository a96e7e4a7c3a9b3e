// Shared declarations for the sm_100a kernels behind include/onda_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/onda_b200.h"

namespace onda {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n);
// optional event bracket around the dominant kernel (api.cu); both are no-ops unless timing is enabled
void timing_begin(cudaStream_t s);
void timing_end(cudaStream_t s);

#define ONDA_CUDA_TRY(expr)                                  \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::onda::cuda_fail(_e, #expr); \
    } while (0)

#define ONDA_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            ::onda::set_error(__VA_ARGS__); \
            return ONDA_EINVAL;            \
        }                                  \
    } while (0)

// ---- geometry ----------------------------------------------------------------
constexpr int kTilePixels = 128;     // pixels per CTA tile (= 4 warps x 32 lanes = 128 TMEM lanes)
constexpr int kStatSlots = ONDA_NUM_STATS;

__host__ __device__ inline int padded_classes(int C) { return C <= 20 ? 20 : 32; }
__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Distance-table layout (floats), built by onda_build_distance_table:
//   sigma[Dp] | w[Dp] | mu[Dp] | bias[32] | Q[Dp][CP]  (channel-major, CP = padded_classes(C))
//   | Bhi[Dp/4][32][4] | Blo[Dp/4][32][4]   TF32 split of -2*Q laid out as the tcgen05 B operand:
//     K-major, no swizzle, 8x16-byte core matrices: element (class n, channel c) sits at float index
//     (c/4)*128 + n*4 + (c%4), so LBO (next 4 channels) = 512 B and SBO (next 8 classes) = 128 B.
// Dp = D rounded up to 32; padded channels carry w = 0, Q = 0 so they contribute nothing.
struct TableLayout {
    int C, D, Dp, CP;
    size_t off_sigma, off_w, off_mu, off_bias, off_q, off_qhi, off_qlo, off_scratch, total;
};
__host__ __device__ inline TableLayout table_layout(int C, int D) {
    TableLayout t;
    t.C = C; t.D = D; t.Dp = round_up(D, 32); t.CP = padded_classes(C);
    t.off_sigma = 0;
    t.off_w = t.off_sigma + t.Dp;
    t.off_mu = t.off_w + t.Dp;
    t.off_bias = t.off_mu + t.Dp;
    t.off_q = t.off_bias + 32;
    t.off_qhi = t.off_q + (size_t)t.Dp * t.CP;
    t.off_qlo = t.off_qhi + (size_t)32 * t.Dp;
    t.off_scratch = t.off_qlo + (size_t)32 * t.Dp;          // per-CTA bias partials (doubles) + ticket of the table kernel
    t.total = t.off_scratch + (size_t)2 * 32 * (t.Dp / 32) + 8;
    return t;
}

// `sums` buffer layout: [C*D sum | C*D sumsq | C count | kStatSlots stats]
__host__ __device__ inline size_t sums_floats(int C, int D) { return (size_t)2 * C * D + C + kStatSlots; }

// ---- small device helpers ------------------------------------------------------
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// torch.max / argmax semantics: first maximal index, NaN beats everything (first NaN wins).
__device__ __forceinline__ bool torch_greater(float v, float best) {
    return (v > best) || (v != v && best == best);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace onda
