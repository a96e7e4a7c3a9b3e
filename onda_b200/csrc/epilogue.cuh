// Per-pixel tail of the fused pass: squared distances -> rectified posterior,
// label, statistics, and the coalesced write-out of the pixel-major results.
// Shared by the CUDA-core kernel, its split-D finishing kernel and the tcgen05 kernel.
#pragma once

#include "common.cuh"

namespace onda {

struct FusedParams {
    const float* feat;     // (B, D, HW)
    const float* prior;    // (B, C, HW) or null
    const float* logits;   // (B, C, HW) or null -> no class sums
    const float* table;    // distance table (common.cuh: TableLayout)
    int B, D, HW, C;
    long long N;           // B * HW
    float tau, thresh;
    long long* labels;     // [N] or null
    float* soft;           // [N][C] or null
    float* dist;           // [N][C] or null
    float* cta_partials;   // [gridDim.x][sums_floats(C, D)] class sums / counts (stats slots unused)
    float* stat_partials;  // [n_stat_ctas][kStatSlots]
    float* dots_scratch;   // [nslices][CP + 1][N] partial dot products when nslices > 1
    int nslices;           // channel slices (gridDim.y)
    int slice_channels;    // channels per slice, multiple of 32
    int tiles;             // ceil(N / 128)
};

// launch plan of the CUDA-core kernel (fused_simt.cu)
struct SimtPlan {
    int tiles, nslices, DS, grid_x, finish_grid;
    size_t smem_bytes;
};
SimtPlan plan_simt(int B, int D, int HW, int C, int sms, bool dist, bool sums);
int launch_fused_simt(const FusedParams& p, const SimtPlan& pl, bool dist, bool sums, cudaStream_t stream);

// tcgen05 kernel (fused_tc.cu)
bool tc_supported(int B, int D, int HW, int C);
int tc_grid(int tiles, int sms);
int launch_fused_tc(const FusedParams& p, int grid, bool sums, cudaStream_t stream);

struct PixelStats {
    float proto_conf = 0.f, prior_conf = 0.f, pl_conf = 0.f, entropy = 0.f;
    int pl_pixels = 0, pixels = 0;
};

// Everything after the squared distances for one pixel (one thread).  Follows
// prototype_handler.pseudo_labels (prototype_handler.py:140-166) step by step:
//   d' = d - min d (:124-125) ; q = softmax(-d'/tau) (:147) ; stat max q (:150) ;
//   r = q*prior ; r /= sum r (:159-160) ; (m, l) = max r ; l = 255 if m < thresh (:163-166)
// with torch's first-index / NaN-first max semantics.  d2 holds squared distances on entry;
// on exit r[] holds the rectified posterior and dsh[] the shifted distances.
template <int CP>
__device__ __forceinline__ void rectify_pixel(const float (&d2)[CP], const float (&pri)[CP], int C, float tau,
                                              float thresh, bool have_prior, float (&r)[CP], float (&dsh)[CP],
                                              int& label, float& m_out, float& maxq, float& maxprior,
                                              float& entropy) {
    float dmin = __int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            float v = d2[k];
            v = v < 0.f ? 0.f : v;  // keeps NaN (a negative pooled variance gives NaN like the reference)
            float d = sqrtf(v);
            dsh[k] = d;
            if (d < dmin || d != d) dmin = d;
        }
    }
    float zmax = -__int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            dsh[k] = dsh[k] - dmin;
            float z = __fdiv_rn(-dsh[k], tau);
            r[k] = z;
            if (z > zmax || z != z) zmax = z;
        }
    }
    float esum = 0.f;
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            float e = exp2f((r[k] - zmax) * 1.4426950408889634f);
            r[k] = e;
            esum += e;
        }
    }
    maxq = -1.f;
    maxprior = -__int_as_float(0x7f800000);
    float rsum = 0.f;
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            float q = __fdiv_rn(r[k], esum);
            if (torch_greater(q, maxq)) maxq = q;
            if (have_prior) {
                if (torch_greater(pri[k], maxprior)) maxprior = pri[k];
                q = q * pri[k];
            }
            r[k] = q;
            rsum += q;
        }
    }
    float best = -__int_as_float(0x7f800000);
    int arg = 0;
    entropy = 0.f;
    const float inv_log2c = 1.f / log2f((float)C);
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            float v = have_prior ? __fdiv_rn(r[k], rsum) : r[k];
            r[k] = v;
            if (torch_greater(v, best)) { best = v; arg = k; }
            entropy -= v * log2f(v + 1e-30f) * inv_log2c;
        }
    }
    m_out = best;
    label = (best < thresh) ? ONDA_IGNORE_LABEL : arg;
}

// Copies `rows` staged rows of C floats (row stride CP+1 in shared memory) to a pixel-major
// global array with fully coalesced 128-byte stores.  One warp, its own 32 rows.
template <int CP>
__device__ __forceinline__ void warp_copy_rows(const float* stage_rows, float* gdst, int rows, int C, int lane) {
    const int total = rows * C;
    const int dr = 32 / C, dk = 32 % C;  // element e -> (row, k) advanced incrementally: no division per store
    int row = lane / C, k = lane % C;
    for (int e = lane; e < total; e += 32) {
        gdst[e] = stage_rows[row * (CP + 1) + k];
        k += dk;
        row += dr;
        if (k >= C) { k -= C; row += 1; }
    }
}

// Tail for one tile row handled by thread `t` (pixel n = tile_base + t).  `stage` is a
// [128][CP+1] shared-memory slab; rows 32*warp .. 32*warp+31 belong to this warp.
template <int CP>
__device__ __forceinline__ void finish_pixel(const FusedParams& p, const float (&d2)[CP], long long tile_base, int t,
                                             float* stage, PixelStats& st) {
    const int lane = t & 31, warp = t >> 5;
    const long long n = tile_base + t;
    const bool valid = n < p.N;
    const bool want_post = (p.labels != nullptr) || (p.soft != nullptr);
    float pri[CP], r[CP], dsh[CP];
    const bool have_prior = p.prior != nullptr;
    if (valid && have_prior && want_post) {
        const long long b = n / p.HW;
        const long long q = n - b * p.HW;
        const float* pp = p.prior + (b * p.C) * (long long)p.HW + q;
#pragma unroll
        for (int k = 0; k < CP; ++k) pri[k] = (k < p.C) ? __ldg(pp + (long long)k * p.HW) : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < CP; ++k) pri[k] = 0.f;
    }
    int label = 0;
    float m = 0.f, maxq = 0.f, maxprior = 0.f, ent = 0.f;
    rectify_pixel<CP>(d2, pri, p.C, p.tau, p.thresh, have_prior && want_post, r, dsh, label, m, maxq, maxprior, ent);
    if (valid) {
        st.proto_conf += maxq;
        st.pixels += 1;
        if (have_prior && want_post) {
            st.prior_conf += maxprior;
            st.pl_conf += m;
            st.entropy += ent;
            st.pl_pixels += (label != ONDA_IGNORE_LABEL) ? 1 : 0;
        }
        if (p.labels != nullptr) p.labels[n] = (long long)label;
    }
    const long long warp_base = tile_base + 32 * warp;
    long long remain = p.N - warp_base;
    const int rows = remain <= 0 ? 0 : (remain < 32 ? (int)remain : 32);
    float* my_rows = stage + (32 * warp) * (CP + 1);
    if (p.soft != nullptr) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < p.C) stage[t * (CP + 1) + k] = r[k];
        __syncwarp();
        warp_copy_rows<CP>(my_rows, p.soft + warp_base * p.C, rows, p.C, lane);
    }
    if (p.dist != nullptr) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < p.C) stage[t * (CP + 1) + k] = dsh[k];
        __syncwarp();
        warp_copy_rows<CP>(my_rows, p.dist + warp_base * p.C, rows, p.C, lane);
    }
}

// Block-level, fixed-order reduction of the per-thread statistics into one partial row.
__device__ __forceinline__ void write_stat_partial(const PixelStats& st, float* red /* [nwarps][kStatSlots] smem */,
                                                   float* out_row, int nwarps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[kStatSlots];
    v[ONDA_STAT_PROTO_CONF] = st.proto_conf;
    v[ONDA_STAT_PRIOR_CONF] = st.prior_conf;
    v[ONDA_STAT_PL_CONF] = st.pl_conf;
    v[ONDA_STAT_PL_PIXELS] = (float)st.pl_pixels;
    v[ONDA_STAT_PIXELS] = (float)st.pixels;
    v[ONDA_STAT_ENTROPY] = st.entropy;
    v[6] = 0.f;
    v[7] = 0.f;
#pragma unroll
    for (int s = 0; s < kStatSlots; ++s) {
        float x = warp_sum(v[s]);
        if (lane == 0) red[warp * kStatSlots + s] = x;
    }
    __syncthreads();
    if (threadIdx.x < kStatSlots) {
        float x = 0.f;
        for (int w = 0; w < nwarps; ++w) x += red[w * kStatSlots + threadIdx.x];
        out_row[threadIdx.x] = x;
    }
}

}  // namespace onda
