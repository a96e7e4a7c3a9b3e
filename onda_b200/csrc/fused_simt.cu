// CUDA-core implementation of the fused pass (any D, C <= 32).
//
// Tile = 128 consecutive pixels of the flattened (B*H*W) pixel axis.  Inside a CTA of four
// warps, warp g owns a quarter of the CTA's channel slice and walks ALL 128 pixels of the tile
// (4 pixels per lane, 32 consecutive pixels per load instruction = one coalesced 128-byte
// request along an NCHW channel plane).  That ownership gives
//   * register tiling: each Q row read from shared memory feeds 4 pixels x CP classes of FMAs;
//   * deterministic class sums without atomics: the per-class accumulators of a channel are
//     only ever touched by the one warp that owns the channel, in tile order.
// The four partial dot-product sets are combined through shared memory in warp order, then
// thread t finishes pixel t (epilogue.cuh).  With more than one channel slice (large D, or few
// tiles) the partial dots go to a scratch array and split_finish_kernel completes the pixels.
#include "epilogue.cuh"

namespace onda {

constexpr int kSimtThreads = 128;
constexpr int kChunk = 8;  // channels per register chunk

struct SimtSmem {
    int DS;  // channels in a slice (multiple of 32)
    size_t q, mu, w, acc, dots, ys, cnt, red, total;
};
__host__ __device__ inline SimtSmem simt_smem(int DS, int C, int CP, bool dist, bool sums) {
    SimtSmem s;
    s.DS = DS;
    s.q = 0;
    s.mu = s.q + (dist ? (size_t)DS * CP : 0);
    s.w = s.mu + (dist ? DS : 0);
    s.acc = s.w + (dist ? DS : 0);
    s.dots = s.acc + (sums ? (size_t)2 * C * DS : 0);
    s.ys = s.dots + (dist ? (size_t)4 * kTilePixels * (CP + 1) : 0);
    s.cnt = s.ys + kTilePixels;
    s.red = s.cnt + 32;
    s.total = s.red + 4 * kStatSlots;
    return s;
}

template <int CP, bool DIST, bool SUMS>
__global__ void __launch_bounds__(kSimtThreads, 2) fused_simt_kernel(const FusedParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, D = p.D, HW = p.HW;
    const int DS = p.slice_channels;
    const int s0 = blockIdx.y * DS;
    const SimtSmem L = simt_smem(DS, C, CP, DIST, SUMS);
    float* Qs = smem + L.q;
    float* mus = smem + L.mu;
    float* wsm = smem + L.w;
    float* acc = smem + L.acc;
    float* dots = smem + L.dots;
    int* ys = reinterpret_cast<int*>(smem + L.ys);
    int* cnt = reinterpret_cast<int*>(smem + L.cnt);
    float* red = smem + L.red;

    const TableLayout T = table_layout(C, D);
    if (DIST) {
        for (int i = tid; i < DS * CP; i += kSimtThreads) {
            int ch = s0 + i / CP;
            Qs[i] = ch < T.Dp ? p.table[T.off_q + (size_t)s0 * CP + i] : 0.f;
        }
        for (int i = tid; i < DS; i += kSimtThreads) {
            int ch = s0 + i;
            mus[i] = ch < T.Dp ? p.table[T.off_mu + ch] : 0.f;
            wsm[i] = ch < T.Dp ? p.table[T.off_w + ch] : 0.f;
        }
    }
    if (SUMS) {
        for (int i = tid; i < 2 * C * DS; i += kSimtThreads) acc[i] = 0.f;
        if (tid < 32) cnt[tid] = 0;
    }
    __syncthreads();

    // channel range of this warp: a quarter of the slice, clipped to D
    const int QS = DS / 4;  // multiple of 8
    const int cbeg = s0 + warp * QS;
    const int cend = min(cbeg + QS, D);
    const bool fused_tail = (p.nslices == 1);
    PixelStats st;

    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const long long tile_base = (long long)tile * kTilePixels;
        // ---- classes of the 128 pixels (first argmax of the EMA logits, prototype_handler.py:83-86)
        if (SUMS) {
            const long long n = tile_base + tid;
            int arg = -1;
            if (n < p.N) {
                const long long b = n / HW, q = n - b * HW;
                const float* lp = p.logits + (b * C) * (long long)HW + q;
                float best = __ldg(lp);
                arg = 0;
                for (int k = 1; k < C; ++k) {
                    float v = __ldg(lp + (long long)k * HW);
                    if (torch_greater(v, best)) { best = v; arg = k; }
                }
            }
            ys[tid] = arg;
        }
        __syncthreads();  // ys visible; previous tile's readers of dots/ys are done

        const float* fptr[4];
        bool valid[4];
        int y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long n = tile_base + 32 * i + lane;
            valid[i] = n < p.N;
            const long long nn = valid[i] ? n : 0;
            const long long b = nn / HW, q = nn - b * HW;
            fptr[i] = p.feat + (b * D) * (long long)HW + q;
            y[i] = SUMS ? ys[32 * i + lane] : -1;
        }
        unsigned class_mask = 0;
        if (SUMS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) class_mask |= (y[i] >= 0) ? (1u << y[i]) : 0u;
            class_mask = __reduce_or_sync(0xffffffffu, class_mask);
            if (warp == 0 && blockIdx.y == 0) {  // pixel counts per class, once per tile
                unsigned rem = class_mask;
                while (rem) {
                    const int k = __ffs(rem) - 1;
                    rem &= rem - 1;
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) c += __popc(__ballot_sync(0xffffffffu, y[i] == k));
                    if (lane == 0) cnt[k] += c;
                }
            }
        }
        const bool uniform = __popc(class_mask) <= 1;

        float dot[4][CP];
        float A[4];
        if (DIST) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                A[i] = 0.f;
#pragma unroll
                for (int k = 0; k < CP; ++k) dot[i][k] = 0.f;
            }
        }

        float xn[4][kChunk];
        auto load_chunk = [&](int c0) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const int c = c0 + j;
                const bool cok = c < cend;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    xn[i][j] = (cok && valid[i]) ? ldg_stream(fptr[i] + (long long)c * HW) : 0.f;
            }
        };
        if (cbeg < cend) load_chunk(cbeg);
        for (int c0 = cbeg; c0 < cend; c0 += kChunk) {
            float x[4][kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i][j] = xn[i][j];
            if (c0 + kChunk < cend) load_chunk(c0 + kChunk);  // prefetch: 32 loads in flight per thread

            if (DIST) {
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                    const int cl = c0 + j - s0;
                    const float m = mus[cl], wv = wsm[cl];
                    float qv[CP];
                    const float4* q4 = reinterpret_cast<const float4*>(Qs + cl * CP);
#pragma unroll
                    for (int k4 = 0; k4 < CP / 4; ++k4) {
                        float4 v = q4[k4];
                        qv[4 * k4 + 0] = v.x; qv[4 * k4 + 1] = v.y; qv[4 * k4 + 2] = v.z; qv[4 * k4 + 3] = v.w;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float xc = x[i][j] - m;
                        A[i] = fmaf(xc * xc, wv, A[i]);
#pragma unroll
                        for (int k = 0; k < CP; ++k) dot[i][k] = fmaf(xc, qv[k], dot[i][k]);
                    }
                }
            }

            if (SUMS) {
                // warp-level reduce-scatter of 8 channel sums + 8 channel sums of squares over the
                // 128 pixels of the tile, once per class present in the tile (once, unmasked, when
                // the tile is class-uniform).  Lane pair (2i, 2i+1) ends with value index i.
                unsigned rem = class_mask;
                while (rem) {
                    const int k = __ffs(rem) - 1;
                    rem &= rem - 1;
                    float v[16];
#pragma unroll
                    for (int j = 0; j < kChunk; ++j) {
                        float s = 0.f, s2 = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float xv = (uniform || y[i] == k) ? x[i][j] : 0.f;
                            s += xv;
                            s2 = fmaf(xv, xv, s2);
                        }
                        v[j] = s;
                        v[kChunk + j] = s2;
                    }
#pragma unroll
                    for (int half = 8; half >= 1; half >>= 1) {
                        const bool up = (lane & (2 * half)) != 0;
#pragma unroll
                        for (int t2 = 0; t2 < half; ++t2) {
                            const float send = up ? v[t2] : v[t2 + half];
                            const float keep = up ? v[t2 + half] : v[t2];
                            v[t2] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * half);
                        }
                    }
                    const float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                    const int idx = lane >> 1;
                    const int c = c0 + (idx & 7);
                    if ((lane & 1) == 0 && c < cend) {
                        float* a = acc + ((size_t)((idx >> 3) * C + k)) * DS + (c - s0);
                        *a += tot;
                    }
                }
            }
        }

        if (DIST) {
            // ---- combine the four channel quarters in warp order
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float* row = dots + ((size_t)warp * kTilePixels + 32 * i + lane) * (CP + 1);
#pragma unroll
                for (int k = 0; k < CP; ++k) row[k] = dot[i][k];
                row[CP] = A[i];
            }
            __syncthreads();
            float d2[CP];
            float a_tot = 0.f;
#pragma unroll
            for (int k = 0; k < CP; ++k) d2[k] = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float* row = dots + ((size_t)g * kTilePixels + tid) * (CP + 1);
#pragma unroll
                for (int k = 0; k < CP; ++k) d2[k] += row[k];
                a_tot += row[CP];
            }
            if (fused_tail) {
                const float* bias = p.table + T.off_bias;
#pragma unroll
                for (int k = 0; k < CP; ++k) d2[k] = fmaf(-2.f, d2[k], a_tot + __ldg(bias + k));
                finish_pixel<CP>(p, d2, tile_base, tid, dots, st);
            } else {
                const long long n = tile_base + tid;
                if (n < p.N) {
                    float* dst = p.dots_scratch + ((size_t)blockIdx.y * (CP + 1)) * p.N + n;
#pragma unroll
                    for (int k = 0; k < CP; ++k) dst[(size_t)k * p.N] = d2[k];
                    dst[(size_t)CP * p.N] = a_tot;
                }
            }
        } else {
            __syncthreads();  // every warp has read ys before the next tile overwrites it
        }
    }

    __syncthreads();
    if (SUMS) {
        float* out = p.cta_partials + (size_t)blockIdx.x * sums_floats(C, D);
        for (int i = tid; i < 2 * C * DS; i += kSimtThreads) {
            const int cl = i % DS;
            const int qk = i / DS;
            if (s0 + cl < D) out[(size_t)qk * D + s0 + cl] = acc[i];
        }
        if (blockIdx.y == 0 && tid < C) out[(size_t)2 * C * D + tid] = (float)cnt[tid];
    }
    if (DIST && fused_tail && blockIdx.y == 0)
        write_stat_partial(st, red, p.stat_partials + (size_t)blockIdx.x * kStatSlots, kSimtThreads / 32);
}

// Completes pixels whose dot products were split over channel slices: sums the slices in
// slice order, then the common per-pixel tail.
template <int CP>
__global__ void __launch_bounds__(kSimtThreads) split_finish_kernel(const FusedParams p) {
    extern __shared__ __align__(16) float smem[];
    float* stage = smem;                                   // [128][CP+1]
    float* red = smem + kTilePixels * (CP + 1);            // [4][kStatSlots]
    const int tid = threadIdx.x;
    const TableLayout T = table_layout(p.C, p.D);
    PixelStats st;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const long long tile_base = (long long)tile * kTilePixels;
        const long long n = tile_base + tid;
        float d2[CP];
        float a_tot = 0.f;
#pragma unroll
        for (int k = 0; k < CP; ++k) d2[k] = 0.f;
        if (n < p.N) {
            for (int s = 0; s < p.nslices; ++s) {
                const float* src = p.dots_scratch + ((size_t)s * (CP + 1)) * p.N + n;
#pragma unroll
                for (int k = 0; k < CP; ++k) d2[k] += src[(size_t)k * p.N];
                a_tot += src[(size_t)CP * p.N];
            }
        }
        const float* bias = p.table + T.off_bias;
#pragma unroll
        for (int k = 0; k < CP; ++k) d2[k] = fmaf(-2.f, d2[k], a_tot + __ldg(bias + k));
        __syncthreads();
        finish_pixel<CP>(p, d2, tile_base, tid, stage, st);
    }
    __syncthreads();
    write_stat_partial(st, red, p.stat_partials + (size_t)blockIdx.x * kStatSlots, kSimtThreads / 32);
}

// ---- host-side planning and launch ----------------------------------------------
SimtPlan plan_simt(int B, int D, int HW, int C, int sms, bool dist, bool sums) {
    SimtPlan pl;
    const long long N = (long long)B * HW;
    const int CP = padded_classes(C);
    const int Dp = round_up(D, 32);
    pl.tiles = (int)((N + kTilePixels - 1) / kTilePixels);
    int DS = Dp < 256 ? Dp : 256;
    // few tiles: slice the channels further so the grid still covers the machine
    while (DS > 32 && (long long)pl.tiles * ((Dp + DS - 1) / DS) < 2LL * sms) {
        int next = round_up(DS / 2, 32);
        if (next >= DS) break;
        DS = next;
    }
    pl.DS = DS;
    pl.nslices = (Dp + DS - 1) / DS;
    int per_slice = (2 * sms) / pl.nslices;
    if (per_slice < 1) per_slice = 1;
    pl.grid_x = pl.tiles < per_slice ? pl.tiles : per_slice;
    if (pl.grid_x < 1) pl.grid_x = 1;
    pl.finish_grid = pl.tiles < 4 * sms ? pl.tiles : 4 * sms;
    if (pl.finish_grid < 1) pl.finish_grid = 1;
    pl.smem_bytes = simt_smem(DS, C, CP, dist, sums).total * sizeof(float);
    return pl;
}

template <int CP, bool DIST, bool SUMS>
static int launch_one(const FusedParams& p, const SimtPlan& pl, cudaStream_t stream) {
    auto kern = fused_simt_kernel<CP, DIST, SUMS>;
    ONDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    timing_begin(stream);
    kern<<<dim3(pl.grid_x, pl.nslices), kSimtThreads, pl.smem_bytes, stream>>>(p);
    timing_end(stream);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    if (DIST && pl.nslices > 1) {
        auto fin = split_finish_kernel<CP>;
        const size_t sm = (size_t)(kTilePixels * (CP + 1) + 4 * kStatSlots) * sizeof(float);
        fin<<<pl.finish_grid, kSimtThreads, sm, stream>>>(p);
        ONDA_CUDA_TRY(cudaGetLastError());
        count_launch(1);
    }
    return ONDA_OK;
}

int launch_fused_simt(const FusedParams& p, const SimtPlan& pl, bool dist, bool sums, cudaStream_t stream) {
    const int CP = padded_classes(p.C);
#define ONDA_DISPATCH(CPV)                                                         \
    if (dist && sums) return launch_one<CPV, true, true>(p, pl, stream);           \
    if (dist) return launch_one<CPV, true, false>(p, pl, stream);                  \
    return launch_one<CPV, false, true>(p, pl, stream);
    if (CP == 20) { ONDA_DISPATCH(20) }
    ONDA_DISPATCH(32)
#undef ONDA_DISPATCH
}

}  // namespace onda
