// CUDA-core implementation of the fused pass (any D, C <= 32).
//
// Tile = 128 consecutive pixels of the flattened (B*H*W) pixel axis.  Inside a CTA of four
// warps, warp g owns a quarter of the CTA's channel slice and walks ALL 128 pixels of the tile
// (4 pixels per lane, 32 consecutive pixels per load instruction = one coalesced 128-byte
// request along an NCHW channel plane).  That ownership gives
//   * register tiling: each Q row read from shared memory feeds 4 pixels x CP classes of FMAs;
//   * deterministic class sums without atomics: the per-class accumulators of a channel are
//     only ever touched by the one warp that owns the channel, in tile order.
// The class sums use a warp-private shared-memory transposition tile: after a 32-channel block has
// been staged, lane = channel walks the 128 pixels with run-length accumulation (cost independent
// of how many classes a tile contains).
// The four partial dot-product sets are combined through shared memory in warp order, then
// thread t finishes pixel t (epilogue.cuh).  With more than one channel slice (large D, or few
// tiles) the partial dots go to a scratch array and split_finish_kernel completes the pixels.
#include "epilogue.cuh"

namespace onda {

constexpr int kSimtThreads = 128;
constexpr int kChunk = 8;  // channels per register chunk

constexpr int kTRow = kTilePixels + 4;  // padded row of the transposition tile (conflict-free LDS.128 by channel)

struct SimtSmem {
    int DS;  // channels in a slice (multiple of 32)
    size_t mu, w, acc, tile, ys, cnt, red, total;
};
__host__ __device__ inline SimtSmem simt_smem(int DS, int C, int CP, bool dist, bool sums) {
    SimtSmem s;
    s.DS = DS;
    s.mu = 0;
    s.w = s.mu + (dist ? DS : 0);
    s.acc = s.w + (dist ? DS : 0);
    s.tile = s.acc + (sums ? (size_t)2 * C * DS : 0);
    // the per-warp transposition tiles [4][32][kTRow] and the dot-product staging [4][128][CP+1] share storage
    const size_t t_tile = sums ? (size_t)4 * 32 * kTRow : 0;
    const size_t t_dots = dist ? (size_t)4 * kTilePixels * (CP + 1) : 0;
    s.ys = s.tile + (t_tile > t_dots ? t_tile : t_dots);
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
    float* mus = smem + L.mu;
    float* wsm = smem + L.w;
    float* acc = smem + L.acc;
    float* dots = smem + L.tile;
    float* Tw = smem + L.tile + (size_t)warp * 32 * kTRow;   // this warp's [32 channels][kTRow pixels] tile
    int* ys = reinterpret_cast<int*>(smem + L.ys);
    int* cnt = reinterpret_cast<int*>(smem + L.cnt);
    float* red = smem + L.red;

    const TableLayout T = table_layout(C, D);
    const float* Qg = p.table + T.off_q;   // [Dp][CP] rows, read through L1 (uniform 16-byte loads)
    if (DIST) {
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
                if (p.class_ids != nullptr) {        // source labels (calculate_prototypes): 255 / out-of-range = not counted
                    const long long id = p.class_ids[n];
                    arg = (id >= 0 && id < C) ? (int)id : -1;
                } else {
                    float lv[CP];
                    load_pixel_row<CP>(p.logits, C, HW, n, lv);
                    arg = first_argmax<CP>(lv, C);
                }
            }
            ys[tid] = arg;
        }
        __syncthreads();  // ys visible; previous tile's readers of the staging area are done

        const float* fptr[4];
        bool valid[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long n = tile_base + 32 * i + lane;
            valid[i] = n < p.N;
            const long long nn = valid[i] ? n : 0;
            const long long b = nn / HW, q = nn - b * HW;
            fptr[i] = p.feat + (b * D) * (long long)HW + q;
        }
        if (SUMS && warp == 0 && blockIdx.y == 0) {  // pixel counts per class, once per tile
            unsigned class_mask = 0;
            int y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[i] = ys[32 * i + lane];
                class_mask |= (y[i] >= 0) ? (1u << y[i]) : 0u;
            }
            class_mask = __reduce_or_sync(0xffffffffu, class_mask);
            while (class_mask) {
                const int k = __ffs(class_mask) - 1;
                class_mask &= class_mask - 1;
                int c = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) c += __popc(__ballot_sync(0xffffffffu, y[i] == k));
                if (lane == 0) cnt[k] += c;
            }
        }

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
        for (int blk = cbeg; blk < cend; blk += 32) {       // 32-channel block = one transposition tile
            const int bend = min(blk + 32, cend);
            for (int c0 = blk; c0 < bend; c0 += kChunk) {
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
                        const float4* q4 = reinterpret_cast<const float4*>(Qg + (size_t)(c0 + j) * CP);
#pragma unroll
                        for (int k4 = 0; k4 < CP / 4; ++k4) {
                            float4 v = __ldg(q4 + k4);
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
                if (SUMS) {   // stage the chunk channel-major: row = channel, consecutive lanes = consecutive pixels
#pragma unroll
                    for (int j = 0; j < kChunk; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) Tw[(c0 - blk + j) * kTRow + 32 * i + lane] = x[i][j];
                }
            }
            if (SUMS) {
                // ---- class sums of this 32-channel block: lane = channel, walks the tile's 128 pixels in
                // order keeping a running (sum, sum of squares) for the current class; a class change flushes
                // into the accumulators this warp alone owns.  Control flow is warp-uniform (every lane sees
                // the same class sequence) and the summation order is fixed -> deterministic, no atomics.
                __syncwarp();
                const int c = blk + lane;
                if (c < bend) {
                    const float4* row = reinterpret_cast<const float4*>(Tw + lane * kTRow);
                    const int4* ys4 = reinterpret_cast<const int4*>(ys);
                    float* a1 = acc + (c - s0);
                    float* a2 = acc + (size_t)C * DS + (c - s0);
                    int cur = -1;
                    float s1 = 0.f, s2 = 0.f;
                    auto flush = [&]() {
                        if (cur >= 0) {
                            a1[(size_t)cur * DS] += s1;
                            a2[(size_t)cur * DS] += s2;
                        }
                    };
                    auto one = [&](int yv, float xv) {
                        if (yv != cur) { flush(); cur = yv; s1 = 0.f; s2 = 0.f; }
                        s1 += xv;
                        s2 = fmaf(xv, xv, s2);
                    };
#pragma unroll 4
                    for (int p4 = 0; p4 < kTilePixels / 4; ++p4) {
                        const int4 yy = ys4[p4];
                        const float4 v = row[p4];
                        if (yy.x == cur && yy.y == cur && yy.z == cur && yy.w == cur) {
                            s1 += (v.x + v.y) + (v.z + v.w);
                            s2 += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
                        } else {
                            one(yy.x, v.x); one(yy.y, v.y); one(yy.z, v.z); one(yy.w, v.w);
                        }
                    }
                    flush();
                }
                __syncwarp();
            }
        }

        if (DIST) {
            // ---- combine the four channel quarters in warp order
            if (SUMS) __syncthreads();   // the staging area aliases the other warps' transposition tiles
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
                if (p.dist != nullptr) finish_pixel<CP, true>(p, p.C, d2, tile_base, tid, dots, st);
                else finish_pixel<CP, false>(p, p.C, d2, tile_base, tid, dots, st);
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
        if (p.dist != nullptr) finish_pixel<CP, true>(p, p.C, d2, tile_base, tid, stage, st);
        else finish_pixel<CP, false>(p, p.C, d2, tile_base, tid, stage, st);
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

int launch_split_finish(const FusedParams& p, int sms, int* n_stat, cudaStream_t stream) {
    int grid = p.tiles < 4 * sms ? p.tiles : 4 * sms;
    if (grid < 1) grid = 1;
    const int CP = padded_classes(p.C);
    const size_t sm = (size_t)(kTilePixels * (CP + 1) + 4 * kStatSlots) * sizeof(float);
    if (CP == 20) split_finish_kernel<20><<<grid, kSimtThreads, sm, stream>>>(p);
    else split_finish_kernel<32><<<grid, kSimtThreads, sm, stream>>>(p);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    *n_stat = grid;
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
