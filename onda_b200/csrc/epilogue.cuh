// Per-pixel tail of the fused pass: squared distances -> rectified posterior,
// label, statistics, and the coalesced write-out of the pixel-major results.
// Shared by the CUDA-core kernel, its split-D finishing kernel and the tcgen05 kernel.
#pragma once

#include "common.cuh"

namespace onda {

struct FusedParams {
    const float* feat;     // (B, D, HW)
    const float* prior;    // (B, C, HW) or null
    const float* logits;   // (B, C, HW) or null -> no class sums (unless class_ids is given)
    const long long* class_ids;   // [N] or null: the class of every pixel given directly (values outside [0, C) are skipped)
    const float* table;    // distance table (common.cuh: TableLayout)
    int B, D, HW, C;
    long long N;           // B * HW
    float tau, inv_tau, thresh;   // inv_tau = 1/tau computed on the host
    long long* labels;     // [N] or null
    float* soft;           // [N][C] or null
    float* dist;           // [N][C] or null
    float* cta_partials;   // [gridDim.x][sums_floats(C, D)] class sums / counts (stats slots unused)
    float* stat_partials;  // [n_stat_ctas][kStatSlots]
    float* dots_scratch;   // [nslices][CP + 1][N] partial dot products when nslices > 1
    unsigned* sched;       // tile counters of the tcgen05 kernel (inside the distance table, TableLayout::off_sched)
    int dynamic_tiles;     // tcgen05 kernel: draw tiles from the counters (ONDA_TC_DYNAMIC_TILES=1) instead of the fixed round-robin
    int nslices;           // channel slices (gridDim.y)
    int slice_channels;    // channels per slice, multiple of 32
    int tiles;             // CUDA-core kernel: ceil(N / 128) tiles of the flattened pixel axis; tcgen05 kernel: B * tiles_per_img
    int tiles_per_img;     // tcgen05 kernel: ceil(HW / 128) -- its tiles never straddle two images (bulk row copies)
    int cluster;           // tcgen05 kernel: CTAs per cluster = channel slices of one tile (slice_channels = D / cluster)
    int nstage;            // tcgen05 kernel: stages of the feature ring in shared memory
    long long* debug;      // optional [gridDim.x][32 warps][8] cycle counters (onda_debug_set_buffer), else null
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
int tc_tiles(int B, int HW);
int tc_slices(int tiles, int D, int sms);
int tc_grid(int tiles, int sms, int slices);
int launch_fused_tc(const FusedParams& p, int grid, bool sums, cudaStream_t stream);
// completes pixels whose dot products were parked per channel slice (fused_simt.cu); returns the number of CTAs that
// wrote statistics partials
int launch_split_finish(const FusedParams& p, int sms, int* n_stat, cudaStream_t stream);

struct PixelStats {
    float proto_conf = 0.f, prior_conf = 0.f, pl_conf = 0.f, entropy = 0.f;
    int pl_pixels = 0, pixels = 0;
};

// Loads the C logits of pixel n (NCHW, coalesced across a warp) with every load in flight at once.
template <int CP>
__device__ __forceinline__ void load_pixel_row(const float* base, int C, int HW, long long n, float (&v)[CP]) {
    const unsigned nu = (unsigned)n, hw = (unsigned)HW;
    const unsigned b = nu / hw, q = nu - b * hw;
    const char* lp = reinterpret_cast<const char*>(base + ((size_t)b * C) * hw + q);
    const size_t plane = (size_t)hw * sizeof(float);
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        v[k] = (k < C) ? __ldg(reinterpret_cast<const float*>(lp)) : 0.f;
        lp += plane;
    }
}

// Same for a pixel whose class-0 element is already addressed: `lp` = &base[(img * C) * HW + pixel]; 32-bit offsets.
template <int CP>
__device__ __forceinline__ void load_pixel_row_at(const float* lp, int C, unsigned hw, float (&v)[CP]) {
#pragma unroll
    for (int k = 0; k < CP; ++k) v[k] = (k < C) ? __ldg(lp + (unsigned)k * hw) : 0.f;
}

// NaN-propagating min / max (one FMNMX each): what `(a < b || a != a) ? a : b` spells with two compares and a select
__device__ __forceinline__ float fmin_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmax_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// First maximal index with torch semantics (prototype_handler.onehot, prototype_handler.py:83-86).
template <int CP>
__device__ __forceinline__ int first_argmax(const float (&v)[CP], int C) {
    float best = v[0];
    int arg = 0;
#pragma unroll
    for (int k = 1; k < CP; ++k)
        if (k < C && torch_greater(v[k], best)) { best = v[k]; arg = k; }
    return arg;
}

// Same result, cheaper in the common case: a NaN-propagating maximum (one FMNMX per class), then the first index that
// holds it (scanned downwards: a compare and a select per class); only a row that contains a NaN takes the full torch
// rule (first NaN wins).  -0 == +0 like torch's `>` scan: the first of them wins.
template <int CP>
__device__ __forceinline__ int first_argmax_fast(const float (&v)[CP], int C) {
    float best = v[0];
#pragma unroll
    for (int k = 1; k < CP; ++k)
        if (k < C) best = fmax_nan(best, v[k]);
    if (best != best) return first_argmax<CP>(v, C);
    int arg = 0;
#pragma unroll
    for (int k = CP - 1; k >= 1; --k)
        if (k < C) arg = (v[k] == best) ? k : arg;
    return (v[0] == best) ? 0 : arg;
}

// ---- fast single-instruction math (MUFU, flush-to-zero forms: no denormal fix-up code): relative error ~2^-22,
// far inside the 1e-5 parity budget; a flushed denormal only ever replaces a value below 1.2e-38 by 0
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Everything after the squared distances for one pixel (one thread).  Follows
// prototype_handler.pseudo_labels (prototype_handler.py:140-166) step by step:
//   d' = d - min d (:124-125) ; q = softmax(-d'/tau) (:147) ; stat max q (:150) ;
//   r = q*prior ; r /= sum r (:159-160) ; (m, l) = max r ; l = 255 if m < thresh (:163-166)
// with torch's first-index max and NaN behaviour (a NaN row keeps label 0: `nan < thresh` is false).
// The divisions by the (row-constant) softmax and renormalisation sums are reciprocal multiplies.
// On entry v[] holds squared distances; on exit v[] holds the rectified posterior r (WANT_DIST:
// dsh[] additionally receives the shifted distances).
template <int CP, bool WANT_DIST>
__device__ __forceinline__ void rectify_pixel(float (&v)[CP], const float (&pri)[CP], int C, float inv_tau, float thresh,
                                              bool have_prior, float (&dsh)[WANT_DIST ? CP : 1], int& label,
                                              float& m_out, float& maxq, float& maxprior, float& entropy) {
    const float inf = __int_as_float(0x7f800000);
    float dmin = inf, dmax = 0.f;
    if (inv_tau > 0.f) {                                         // the usual case: the row maximum is not needed
#pragma unroll
        for (int k = 0; k < CP; ++k) {
            if (k < C) {
                const float d = fast_sqrt(fmax_nan(v[k], 0.f));  // rounding can make a ~0 squared distance negative; keeps NaN
                v[k] = d;
                dmin = fmin_nan(d, dmin);                        // a NaN distance makes the whole row NaN, like torch.min
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < CP; ++k) {
            if (k < C) {
                const float d = fast_sqrt(fmax_nan(v[k], 0.f));
                v[k] = d;
                dmin = fmin_nan(d, dmin);
                dmax = fmaxf(dmax, d);
            }
        }
    }
    // softmax(-d'/tau) = 2^((d' - ref) * scale) / sum, ref = the d' whose exponent is the row maximum:
    // 0 for tau > 0 (the usual case), max d' for a negative tau
    const float scale = -1.4426950408889634f * inv_tau;
    const float ref = inv_tau > 0.f ? 0.f : dmax - dmin;
    float esum = 0.f, emax = 0.f;
    if (WANT_DIST) {
#pragma unroll
        for (int k = 0; k < CP; ++k) {
            if (k < C) {
                const float ds = v[k] - dmin;
                dsh[k] = ds;
                const float e = fast_ex2((ds - ref) * scale);
                v[k] = e;
                esum += e;
                emax = fmaxf(emax, e);
            }
        }
    } else {
        // (d - dmin - ref) * scale as one fused multiply-add per class: the rounding of the row constant c0 is a factor
        // common to the whole row, which the normalisation by the row sum removes again
        const float c0 = -(dmin + ref) * scale;
#pragma unroll
        for (int k = 0; k < CP; ++k) {
            if (k < C) {
                const float e = fast_ex2(fmaf(v[k], scale, c0));
                v[k] = e;
                esum += e;
                emax = fmaxf(emax, e);
            }
        }
    }
    const float inv_e = fast_rcp(esum);
    maxq = emax * inv_e;
    if (esum != esum) maxq = esum;                         // NaN row
    maxprior = -inf;
    float rsum = 0.f;
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            float q = v[k] * inv_e;
            if (have_prior) {
                maxprior = fmaxf(maxprior, pri[k]);
                q = q * pri[k];
            }
            v[k] = q;
            rsum += q;
        }
    }
    const float inv_r = have_prior ? fast_rcp(rsum) : 1.f;
    float best = -inf;
    int arg = 0;
    entropy = 0.f;
#pragma unroll
    for (int k = 0; k < CP; ++k) {
        if (k < C) {
            const float r = v[k] * inv_r;
            v[k] = r;
            if (r > best) { best = r; arg = k; }
            entropy = fmaf(r, fast_lg2(r + 1e-30f), entropy);
        }
    }
    entropy = -entropy / log2f((float)C);
    // torch.max on a NaN row: value NaN, first index.  A row whose rectified sum is exactly zero (q * prior underflowed
    // for every class) is such a row in the reference too (0 / 0), so it keeps label 0 and a NaN posterior.
    if (rsum != rsum || rsum == 0.f) { best = __int_as_float(0x7fc00000); arg = 0; }
    m_out = best;
    label = (best < thresh) ? ONDA_IGNORE_LABEL : arg;
}

// Copies `rows` staged rows of C floats (row stride S in shared memory) to a pixel-major global array with fully
// coalesced 128-byte stores.  One warp, its own 32 rows.  S is the smallest ODD number >= C (conflict-free staging
// writes); for an odd class count (19 in every OnDA config) the staged rows are therefore contiguous and the copy is
// flat: one load and one store per 32 floats, no index arithmetic.
template <int CP>
__device__ __forceinline__ void warp_copy_rows(const float* stage_rows, float* gdst, int rows, int C, int S, int lane) {
    const int total = rows * C;
    if (S == C) {
        if (rows == 32) {
#pragma unroll
            for (int i = 0; i < CP; ++i)
                if (i < C) gdst[lane + 32 * i] = stage_rows[lane + 32 * i];
        } else {
            for (int e = lane; e < total; e += 32) gdst[e] = stage_rows[e];
        }
        return;
    }
    const int dr = 32 / C, dk = 32 % C;  // element e -> (row, k) advanced incrementally: no division per store
    int row = lane / C, k = lane % C;
    for (int e = lane; e < total; e += 32) {
        gdst[e] = stage_rows[row * S + k];
        k += dk;
        row += dr;
        if (k >= C) { k -= C; row += 1; }
    }
}

// Tail for one tile row handled by thread `t` (pixel n = tile_base + t).  `stage` is a
// [128][CP+1] shared-memory slab (rows are staged at the odd stride C | 1 <= CP + 1); rows 32*warp .. 32*warp+31 belong to this warp.
// `tile_rows` = number of valid rows of this tile (rows tile_base .. tile_base + tile_rows - 1 exist).
template <int CP, bool WANT_DIST>
__device__ __forceinline__ void finish_pixel_rows(const FusedParams& p, const int C, float (&d2)[CP], long long tile_base,
                                                  int tile_rows, int t, float* stage, PixelStats& st,
                                                  const float* preloaded_prior = nullptr) {
    const int lane = t & 31, warp = t >> 5;
    const long long n = tile_base + t;
    const bool valid = t < tile_rows;
    const bool want_post = (p.labels != nullptr) || (p.soft != nullptr);
    float pri[CP];
    float dsh[WANT_DIST ? CP : 1];
    const bool have_prior = p.prior != nullptr;
    if (preloaded_prior != nullptr) {
#pragma unroll
        for (int k = 0; k < CP; ++k) pri[k] = preloaded_prior[k];
    } else if (valid && have_prior && want_post) {
        load_pixel_row<CP>(p.prior, C, p.HW, n, pri);
    } else {
#pragma unroll
        for (int k = 0; k < CP; ++k) pri[k] = 0.f;
    }
    int label = 0;
    float m = 0.f, maxq = 0.f, maxprior = 0.f, ent = 0.f;
    rectify_pixel<CP, WANT_DIST>(d2, pri, C, p.inv_tau, p.thresh, have_prior && want_post, dsh, label, m, maxq,
                                 maxprior, ent);
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
    const int remain = tile_rows - 32 * warp;
    const int rows = remain <= 0 ? 0 : (remain < 32 ? remain : 32);
    const int S = C | 1;                                  // row stride of the staged rows: odd, <= CP + 1
    float* my_rows = stage + (32 * warp) * S;
    if (p.soft != nullptr) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < C) stage[t * S + k] = d2[k];
        __syncwarp();
        warp_copy_rows<CP>(my_rows, p.soft + warp_base * C, rows, C, S, lane);
    }
    if (WANT_DIST && p.dist != nullptr) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < C) stage[t * S + k] = dsh[k];
        __syncwarp();
        warp_copy_rows<CP>(my_rows, p.dist + warp_base * C, rows, C, S, lane);
    }
}

// Tile of the flattened pixel axis: rows beyond N do not exist.
template <int CP, bool WANT_DIST>
__device__ __forceinline__ void finish_pixel(const FusedParams& p, const int C, float (&d2)[CP], long long tile_base, int t,
                                             float* stage, PixelStats& st, const float* preloaded_prior = nullptr) {
    const long long remain = p.N - tile_base;
    const int tile_rows = remain <= 0 ? 0 : (remain < kTilePixels ? (int)remain : kTilePixels);
    finish_pixel_rows<CP, WANT_DIST>(p, C, d2, tile_base, tile_rows, t, stage, st, preloaded_prior);
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
