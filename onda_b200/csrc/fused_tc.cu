// tcgen05 implementation of the fused pass for sm_100a (D = 128 or 256, C <= 32).
//
// Why tensor cores: the CUDA-core kernel (fused_simt.cu) is issue-bound -- 21 FFMA per feature
// element for the 19-wide contraction alone (profiles/r1_simt_v1_ncu_full.txt) -- so the distance
// contraction moves to the 5th-generation tensor cores with an error-compensated 3xTF32 split:
//     x' . Q  =  hi(x').hi(Q) + hi(x').lo(Q) + lo(x').hi(Q)      (fp32 accumulate in TMEM)
// which keeps fp32-level accuracy (the dropped lo.lo term is 2^-22 relative).
//
// One persistent CTA per SM, 24 warps, over a single global chunk sequence (chunk = 128 pixels x
// 32 channels; a tile of 128 pixels is D/32 consecutive chunks):
//   warps  0-15  workers, four groups of four; warp w%4 is the pixel quarter (the only TMEM lanes a warp
//                may touch are 32*(w%4)..+31), w/4 the group; group g takes chunks g, g+4, g+8, ...
//                Per chunk: (1) lane = pixel: coalesced 4-byte loads along the NCHW channel planes (the planes
//                are only 4-byte aligned: no TMA), issued eight at a time between the phases below so that
//                the warp never sits blocked behind the SM's miss queue; (2) raw values -> the group's
//                shared-memory tile [pixel][channel]; (3) centred values, split into TF32 hi/lo with packed
//                f32x2 math -> TMEM (tcgen05.st) as the A operand (lane = pixel, column = channel), and
//                sum_j w_j x'_j^2 -> a spare TMEM column of the same lane; (4) class sums of the chunk:
//                lane = channel, each of the group's four warps walks 32 entries of the class-sorted pixel
//                list and adds every segment's (sum, sum of squares) to that class's shared-memory
//                accumulators; a range that starts inside a class parks that first segment in a spare row
//                which the warp that started the class adds after the group's barrier.  One writer per
//                accumulator at a time, fixed summation order, no atomics.
//   warps 16-19  epilogue: tcgen05.ld of the 32 accumulator columns and the partial columns of their pixel,
//                then the common per-pixel tail (epilogue.cuh): sqrt, softmax, prior rectification, label,
//                statistics.
//   warps 20-21  sorter: per tile, class of every pixel (first argmax of the EMA logits) and a stable
//                counting sort of the 128 pixels by class, published for the workers (double-buffered).
//   warp   22    MMA issuer: per chunk 4 K-steps x 3 tcgen05.mma (kind::tf32, M=128, N=32, K=8),
//                A from TMEM, B = TF32 split of -2*w*(P-mu) resident in shared memory (one bulk
//                asynchronous copy in the prologue).
// Every mbarrier has one producer side and one consumer side that visit it phase by phase, in order.
// Two accumulator buffers of 32 TMEM columns; tcgen05.commit frees A stages / publishes accumulators.
// Shared memory is kept under 195 KB on purpose (tc_smem): the next carve-out step leaves 28 KB of L1 and
// costs a quarter of the speed.
#include <utility>
#include "epilogue.cuh"

namespace onda {

constexpr int kTcWorkerWarps = 16;
constexpr int kTcEpiWarp0 = 16;
constexpr int kTcSortWarp0 = 20;                 // two sorter warps, two pixels per lane
constexpr int kTcMmaWarp = 22;                   // warp 23 is idle: 24 warps = six full warpgroups -> 80 registers per thread
constexpr int kTcThreads = 24 * 32;
constexpr int kTcGroups = 4;                     // worker groups = TMEM A stages = class-sum tiles
constexpr int kTcChunkC = 32;                    // channels per chunk
constexpr int kTcTRow = 34;                      // floats per pixel row of a class-sum tile: 32 channels + 2 (conflict-free STS.64 and LDS.32)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol0 = kTcGroups * 64;    // accumulators after the A stages
constexpr uint32_t kApartCol0 = kAccCol0 + 64;   // then sum_j w_j x'_j^2 of each chunk: 2 tile parities x 8 chunks, lane = pixel
constexpr int kTcHeadFloats = kTcGroups * 3 * kTcChunkC;   // per statistic: head partials of ranges 1..3 of every group's current chunk
constexpr uint32_t kSpinLimit = 20000000u;       // failed probes (each followed by a <=256 ns sleep) before giving up

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One non-blocking probe of an mbarrier phase (test_wait: try_wait may park in the shared-memory pipeline
// for a hardware time-out, and ~20 parked waiters starve the LDS/STS of the working warps).
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Wait with a fixed sleep between probes.  Polling costs issue slots and shared-memory pipeline slots that the
// working warps need, so roles that wait for long events (epilogue, sorter) pass a long sleep.
template <int kSleepNs = 200>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    uint32_t tries = 0;
    while (!mbar_try(bar, parity)) {
        __nanosleep(kSleepNs);
        if (++tries > kSpinLimit) __trap();   // a stuck pipeline traps instead of hanging the GPU
    }
}
// wait that adds its duration to a diagnostic counter when profiling is on
template <int kSleepNs = 200>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool prof, long long& acc_cycles) {
    if (!prof) { mbar_wait<kSleepNs>(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait<kSleepNs>(bar, parity);
    acc_cycles += clock64() - t0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t cvt_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// packed fp32 pairs (sm_100 FADD2 / FFMA2): one instruction, two channels
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint64_t v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem descriptor]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_st1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

// x[J] = feature at plane J of a pixel: 32-bit element index off0 + J * plane (one IMAD with an immediate plane index,
// no table of offsets in registers, no dependent chain), then one IMAD.WIDE onto the base; the load bypasses L1
// allocation (streamed once)
template <int J>
__device__ __forceinline__ float ldg_plane(const float* feat, unsigned off0, unsigned plane) {
    float v;
    const float* a = feat + (off0 + (unsigned)J * plane);      // 32-bit element index (tc_supported: the map has < 2^32 elements)
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(a));
    return v;
}
template <int N, int... Js>
__device__ __forceinline__ void ldg_planes(float (&x)[N], const float* feat, unsigned off0, unsigned plane, std::integer_sequence<int, Js...>) {
    ((x[Js] = ldg_plane<Js>(feat, off0, plane)), ...);
}
template <int J0, int N, int... Js>
__device__ __forceinline__ void ldg_planes_part(float (&x)[N], const float* feat, unsigned off0, unsigned plane, std::integer_sequence<int, Js...>) {
    ((x[J0 + Js] = ldg_plane<J0 + Js>(feat, off0, plane)), ...);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_bdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, dense, M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- shared-memory carve-up ------------------------------------------------------------------------
struct TcSmem {
    size_t bhi, blo, tiles, acc, out, mu, w, eoff, ecls, cuts, wc, cnt, red, bars, tmem_ptr, total;  // byte offsets
};
__host__ __device__ inline TcSmem tc_smem(int D, int C, int CP, bool sums) {
    // Everything whose size does not depend on D or C comes first: its addresses are compile-time offsets from the
    // shared-memory base and cost no registers (the workers have none to spare).
    // The total matters beyond fitting: L1 is what the SM's 256 KB leave after the shared-memory carve-out, the
    // carve-out is one of a few sizes (.., 164, 196, 228 KB), and the kernel is measurably slower with the 28 KB of L1
    // that 228 KB leave than with the 60 KB of the 196 KB step (the feature loads straddle 128-byte lines; the
    // second line is the next quarter's first).  Keep total + 1 KB (system) <= 196 KB.
    TcSmem s;
    size_t o = 0;
    s.bars = o; o += 256;                              // (2 * kTcGroups + 8) mbarriers
    s.tmem_ptr = o; o += 128;
    s.eoff = o; o += (size_t)2 * kTilePixels * 4;      // per tile parity: row offset of every class-sorted entry | segment-end flag
    s.ecls = o; o += (size_t)2 * kTilePixels * 4;      // ... and its accumulator row (class * D; -1 = head segment of its range)
    s.cuts = o; o += 128;                              // ... [0] live entries, [1..3] the classes cut by the range starts 32, 64, 96
    s.wc = o; o += 640;                                // per-warp class histograms of the sorter (4 x 36)
    s.cnt = o; o += 128;
    s.red = o; o += 4 * kStatSlots * 4;
    s.tiles = o; o += sums ? (size_t)kTcGroups * kTilePixels * kTcTRow * 4 : 0;
    s.out = o; o += (size_t)kTilePixels * (CP + 1) * 4;
    s.mu = o; o += (size_t)D * 4;
    s.w = o; o += (size_t)D * 4;
    s.bhi = o; o += (size_t)32 * D * 4;
    s.blo = o; o += (size_t)32 * D * 4;
    s.acc = o; o += sums ? (size_t)2 * (C * D + kTcHeadFloats) * 4 : 0;   // [class * D/32 + chunk | then 3 heads per group][sum | sum of squares][32 channels]
    s.total = o;
    return s;
}

template <int CP, int CE, bool SUMS, bool WANT_DIST, bool PROF>
__global__ void __launch_bounds__(kTcThreads, 1) fused_tc_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = CE > 0 ? CE : p.C;      // CE: class count known at compile time (19 in every OnDA config) -> no k < C predication
    const int D = p.D, HW = p.HW;
    const int NB = D / kTcChunkC;
    const int nb_shift = NB == 8 ? 3 : 2;              // D = 256 or 128 (tc_supported)
    const TcSmem L = tc_smem(D, C, CP, SUMS);
    float* Bhi = reinterpret_cast<float*>(smem_raw + L.bhi);
    float* Blo = reinterpret_cast<float*>(smem_raw + L.blo);
    float* Tst = reinterpret_cast<float*>(smem_raw + L.tiles);
    float* acc = reinterpret_cast<float*>(smem_raw + L.acc);
    float* out_stage = reinterpret_cast<float*>(smem_raw + L.out);
    float* mus = reinterpret_cast<float*>(smem_raw + L.mu);
    float* wsm = reinterpret_cast<float*>(smem_raw + L.w);
    int* eoff = reinterpret_cast<int*>(smem_raw + L.eoff);
    int* ecls = reinterpret_cast<int*>(smem_raw + L.ecls);
    int* cuts = reinterpret_cast<int*>(smem_raw + L.cuts);
    int* wc = reinterpret_cast<int*>(smem_raw + L.wc);
    int* cnt = reinterpret_cast<int*>(smem_raw + L.cnt);
    float* red = reinterpret_cast<float*>(smem_raw + L.red);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + L.tmem_ptr);
    const uint32_t bars = smem_u32(smem_raw + L.bars);
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto empty_a = [&](int s) { return bars + 8u * (kTcGroups + s); };
    auto acc_full = [&](int i) { return bars + 8u * (2 * kTcGroups + i); };
    auto acc_empty = [&](int i) { return bars + 8u * (2 * kTcGroups + 2 + i); };
    auto sort_ready = [&](int i) { return bars + 8u * (2 * kTcGroups + 4 + i); };
    auto sort_free = [&](int i) { return bars + 8u * (2 * kTcGroups + 6 + i); };
    const uint32_t btab_bar = bars + 8u * (2 * kTcGroups + 8);

    const TableLayout T = table_layout(C, D);
    // ---- one-time setup: tables to shared memory, barriers, tensor memory
    for (int i = tid; i < D; i += kTcThreads) {
        mus[i] = -p.table[T.off_mu + i];      // negated: the workers centre with a packed add
        wsm[i] = p.table[T.off_w + i];
    }
    if (SUMS) {
        for (int i = tid; i < 2 * (C * D + kTcHeadFloats); i += kTcThreads) acc[i] = 0.f;
        if (tid < 32) cnt[tid] = 0;
    }
    if (tid == 0) {
        for (int s = 0; s < kTcGroups; ++s) {
            mbar_init(full_a(s), 128);
            mbar_init(empty_a(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), 128);
            mbar_init(sort_ready(i), 64);
            mbar_init(sort_free(i), kTcWorkerWarps);
        }
        mbar_init(btab_bar, 1);
        fence_barrier_init();
        // B operand tables (hi | lo, adjacent in the table and in shared memory): one bulk asynchronous copy, off
        // everybody's critical path -- only the MMA issuer waits for it, before its first MMA
        const uint32_t bytes = (uint32_t)(2 * 32 * D) * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(btab_bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(Bhi)), "l"(p.table + T.off_qhi), "r"(bytes), "r"(btab_bar) : "memory");
    }
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();       // B tables were written through the generic proxy; the tensor core reads them through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr bool prof = PROF;       // per-warp wait counters (onda_debug_set_buffer); compiled out of the production kernel
    long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = PROF ? clock64() : 0;
    const int my_tiles = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_chunks = my_tiles * NB;
    const unsigned HWu = (unsigned)HW, Nu = (unsigned)p.N;

    if (warp < kTcWorkerWarps) {
        // =========================== workers ===========================================
        const int quarter = warp & 3, group = warp >> 2;
        const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
        float* Tg = Tst + (size_t)group * kTilePixels * kTcTRow;
        const uint32_t tg_lane = smem_u32(Tg) + 4u * lane;           // class sums: lane = channel of the chunk
        const int gbar = 3 + group;                                   // named barrier of the group (128 threads)
        float x[kTcChunkC];
        const unsigned plane = HWu;
        auto load_base = [&](int q) {   // address of channel b*32 of this lane's pixel in chunk q
            const int t = q >> nb_shift, b = q & (NB - 1);
            const unsigned tile = blockIdx.x + (unsigned)t * gridDim.x;
            unsigned n = tile * kTilePixels + 32 * quarter + lane;
            n = n < Nu ? n : Nu - 1;         // clamp: results of padded rows are never stored (epilogue / sorter guard them)
            const unsigned bimg = n / HWu, pix = n - bimg * HWu;
            return (bimg * (unsigned)D + (unsigned)(b * kTcChunkC)) * HWu + pix;
        };
        auto issue_loads = [&](int q) { ldg_planes(x, p.feat, load_base(q), plane, std::make_integer_sequence<int, kTcChunkC>{}); };
        if (group < total_chunks) issue_loads(group);
        for (int q = group; q < total_chunks; q += kTcGroups) {
            const int t = q >> nb_shift, b = q & (NB - 1);
            const int par = t & 1;
            const uint32_t use = (uint32_t)q >> 2;
            const long long t_it0 = prof ? clock64() : 0;
            if (SUMS) {   // raw values into the group's tile, pixel-major, two channels per 8-byte store
                float2* trow = reinterpret_cast<float2*>(Tg + (size_t)(32 * quarter + lane) * kTcTRow);
#pragma unroll
                for (int j = 0; j < kTcChunkC / 2; ++j) trow[j] = make_float2(x[2 * j], x[2 * j + 1]);
                named_bar_sync(gbar, 128);          // all 128 pixels of the chunk are staged
            }
            if (prof) dbg[4] += clock64() - t_it0;
            if (t >= 2) mbar_wait_t(acc_empty(par), (((uint32_t)t >> 1) - 1) & 1, prof, dbg[0]);   // accumulator and partials [par] of tile t-2 consumed
            mbar_wait_t(empty_a(group), (use & 1) ^ 1, prof, dbg[2]);
            tc_fence_after();
            const long long t_cv0 = prof ? clock64() : 0;
            const bool more = q + kTcGroups < total_chunks;
            const unsigned nsrc = load_base(more ? q + kTcGroups : q);
            uint64_t a2 = 0;                       // sum_j w_j x'_j^2 of the even / odd channels (packed f32x2 math)
            const uint32_t tcol = tmem_base + lane_base + (uint32_t)group * 64;
            const ulonglong2* mu4 = reinterpret_cast<const ulonglong2*>(mus + b * kTcChunkC);     // -mu, two pairs per load
            const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(wsm + b * kTcChunkC);
            const uint64_t neg1 = pack2(-1.f, -1.f);
#pragma unroll
            for (int part = 0; part < 4; ++part) {       // eight channels at a time: bounds the live registers
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                    const ulonglong2 m = mu4[part * 2 + j4], wv = w4[part * 2 + j4];
                    const uint64_t mm[2] = {m.x, m.y}, ww[2] = {wv.x, wv.y};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = j4 * 4 + 2 * e;
                        const uint64_t xc = fadd2(pack2(x[part * 8 + j], x[part * 8 + j + 1]), mm[e]);
                        a2 = ffma2(fmul2(xc, xc), ww[e], a2);
                        const uint32_t h0 = ((uint32_t)xc + 0x1000u) & 0xffffe000u;            // round to TF32 (10-bit mantissa)
                        const uint32_t h1 = ((uint32_t)(xc >> 32) + 0x1000u) & 0xffffe000u;
                        const uint64_t l2 = ffma2((uint64_t)h0 | ((uint64_t)h1 << 32), neg1, xc);     // exact remainder
                        hi[j] = h0;
                        hi[j + 1] = h1;
                        lo[j] = (uint32_t)l2;
                        lo[j + 1] = (uint32_t)(l2 >> 32);
                    }
                }
                tc_st8(tcol + part * 8, hi);
                tc_st8(tcol + 32 + part * 8, lo);
                if (more) {   // the registers of channels 0..15 are free again: their next loads start here
                    if (part == 1) ldg_planes_part<0>(x, p.feat, nsrc, plane, std::make_integer_sequence<int, 8>{});
                    if (part == 3) ldg_planes_part<8>(x, p.feat, nsrc, plane, std::make_integer_sequence<int, 8>{});
                }
            }
            const float a = __uint_as_float((uint32_t)a2) + __uint_as_float((uint32_t)(a2 >> 32));
            tc_st1(tmem_base + lane_base + kApartCol0 + (uint32_t)(par * 8 + b), __float_as_uint(a));     // read by this pixel's epilogue thread
            const long long t_cv1 = prof ? clock64() : 0;
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(full_a(group));
            const long long t_cv2 = prof ? clock64() : 0;
            // The next chunk's loads fly during the class sums.  With class sums they are issued eight at a time between
            // the batches of the summation: 32 back-to-back loads from every warp of a group fill the SM's miss queue and
            // the warp would sit blocked at the issue (measured: a quarter of the worker's time).
            if (!SUMS && more) {
                ldg_planes_part<16>(x, p.feat, nsrc, plane, std::make_integer_sequence<int, 16>{});
            }
            if (prof) {
                dbg[5] += t_cv1 - t_cv0;                 // centring / splitting / tcgen05.st issue
                dbg[6] += t_cv2 - t_cv1;                 // tcgen05.wait::st + arrive
                dbg[0] += clock64() - t_cv2;             // issuing the next chunk's loads (acc_empty waits are negligible)
            }

            if (SUMS) {
                // ---- class sums of this chunk's 32 channels (lane = channel).  The class-sorted pixels of the tile are cut
                // into four ranges of 32 entries, one per warp of the group: balanced whatever the label map looks like.
                // Lane e keeps entry e's row offset and the offset of its accumulator row in registers (broadcast by
                // shuffle: no dependent shared-memory loads), eight loads are in flight at once, and the segment ends
                // are a warp-uniform bit mask: an end adds the running (sum, sum of squares) to that row.  A range that
                // starts inside a class accumulates that first segment (its "head") into a spare row -- same code, the
                // sorter just marks those entries -- which the warp that started the class adds to the class row after
                // the group's barrier, heads in range order.  One writer per accumulator at a time, fixed order, no
                // atomics.
                mbar_wait_t(sort_ready(par), ((uint32_t)t >> 1) & 1, prof, dbg[1]);
                const long long t_seg0 = prof ? clock64() : 0;
                const int* eo = eoff + par * kTilePixels;
                const int* ec = ecls + par * kTilePixels;
                const int* ct = cuts + par * 8;
                // accumulators: row (class * NB + b), then [sum | sum of squares][32 channels]: the pair of a flush is 128 bytes apart
                const uint32_t a1 = smem_u32(acc) + (uint32_t)(b * 64 + lane) * 4u;     // + row offset of the class (bytes, from the sorter)
                const uint32_t head_base = (uint32_t)(C * NB + group * 3 - b) * 256u;   // head j of this group, relative to a1: + (j-1)*256
                const int idx = 32 * quarter + lane;                                    // this warp's range: entries 32*quarter .. +31
                int nlive = ct[0] - 32 * quarter;                                       // ct[0] = live entries of the tile
                nlive = nlive < 0 ? 0 : (nlive > 32 ? 32 : nlive);
                const int eo_i = eo[idx], er_i = ec[idx];
                const unsigned endbits = __ballot_sync(0xffffffffu, eo_i & 1);          // segment ends (sorter: class end or entry 31)
                const int myoff = eo_i & ~1;
                const uint32_t myrow = er_i < 0 ? head_base + (uint32_t)(quarter - 1) * 256u : (uint32_t)er_i;   // accumulator row (bytes) of this entry's segment
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int e0 = 0; e0 < 32; e0 += 8) {
                    if (more) {
                        if (e0 == 0) ldg_planes_part<16>(x, p.feat, nsrc, plane, std::make_integer_sequence<int, 8>{});
                        if (e0 == 16) ldg_planes_part<24>(x, p.feat, nsrc, plane, std::make_integer_sequence<int, 8>{});
                    }
                    if (e0 >= nlive) continue;
                    float xv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t ad = tg_lane + (uint32_t)__shfl_sync(0xffffffffu, myoff, e0 + e);
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[e]) : "r"(ad));
                    }
                    // entries past the last live one need no guard: that one ends a segment, so whatever they add to
                    // the running pair is never flushed
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        s1 += xv[e];
                        s2 = fmaf(xv[e], xv[e], s2);
                        if (endbits & (1u << (e0 + e))) {
                            const uint32_t ad = a1 + __shfl_sync(0xffffffffu, myrow, e0 + e);
                            sts32(ad, lds32(ad) + s1);
                            sts32(ad + 128u, lds32(ad + 128u) + s2);
                            s1 = 0.f;
                            s2 = 0.f;
                        }
                    }
                }
                if (prof) dbg[3] += clock64() - t_seg0;
                named_bar_sync(gbar, 128);          // the group is done reading its tile; all heads are complete
                {   // move the heads of the classes this warp started into their class rows
                    const int hj = (lane >= 1 && lane < 4) ? ct[lane] : -1;
                    unsigned m = __ballot_sync(0xffffffffu, hj >= 0 && (hj >> 8) == quarter);
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        const int k = __shfl_sync(0xffffffffu, hj, j) & 0xff;
                        const uint32_t hd = a1 + head_base + (uint32_t)(j - 1) * 256u, ad = a1 + (uint32_t)(k * NB) * 256u;
                        sts32(ad, lds32(ad) + lds32(hd));
                        sts32(ad + 128u, lds32(ad + 128u) + lds32(hd + 128u));
                        sts32(hd, 0.f);
                        sts32(hd + 128u, 0.f);
                    }
                }
                if (q + kTcGroups >= total_chunks || ((q + kTcGroups) >> nb_shift) != t) {   // last chunk of this tile for the group
                    if (lane == 0) mbar_arrive(sort_free(par));
                }
            }
        }
    } else if (warp == kTcMmaWarp) {
        // =========================== MMA issuer ========================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, 32);
            const uint32_t bhi_addr = smem_u32(Bhi), blo_addr = smem_u32(Blo);
            mbar_wait<32>(btab_bar, 0);             // the B tables have landed (bulk copy of the prologue)
            for (int t = 0; t < my_tiles; ++t) {
                const int par = t & 1;
                if (t >= 2) mbar_wait_t(acc_empty(par), (((uint32_t)t >> 1) - 1) & 1, prof, dbg[0]);
                tc_fence_after();
                const uint32_t dcol = tmem_base + kAccCol0 + (uint32_t)par * 32;
                for (int b = 0; b < NB; ++b) {
                    const int q = t * NB + b;
                    const int stage = q & 3;
                    const uint32_t use = (uint32_t)q >> 2;
                    mbar_wait_t<32>(full_a(stage), use & 1, prof, dbg[1]);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t a_hi = tmem_base + (uint32_t)stage * 64 + ks * 8;
                        const uint32_t a_lo = a_hi + 32;
                        const uint32_t koff = (uint32_t)((b * kTcChunkC + ks * 8) / 4) * 512u;   // 4 channels = one 512-byte slab
                        const uint64_t d_hi = make_bdesc(bhi_addr + koff, 512, 128);
                        const uint64_t d_lo = make_bdesc(blo_addr + koff, 512, 128);
                        tc_mma_tf32_ts(dcol, a_hi, d_hi, idesc, (b | ks) != 0);
                        tc_mma_tf32_ts(dcol, a_hi, d_lo, idesc, 1);
                        tc_mma_tf32_ts(dcol, a_lo, d_hi, idesc, 1);
                    }
                    tc_commit(empty_a(stage));          // A stage reusable once these MMAs retire
                }
                tc_commit(acc_full(par));               // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp >= kTcEpiWarp0 && warp < kTcEpiWarp0 + 4) {
        // =========================== epilogue ===========================================
        const int et = tid - kTcEpiWarp0 * 32;          // 0..127 = pixel row of the tile = TMEM lane
        const uint32_t lane_base = (uint32_t)(et & ~31) << 16;
        PixelStats st;
        const float* bias = p.table + T.off_bias;
        for (int t = 0; t < my_tiles; ++t) {
            const int par = t & 1;
            const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
            float pri[CP];
            {   // prior row of this pixel: issued before the wait so its latency hides behind the MMAs
                const long long n = tile * kTilePixels + et;
                if (n < p.N && p.prior != nullptr && (p.labels != nullptr || p.soft != nullptr)) {
                    load_pixel_row<CP>(p.prior, C, HW, n, pri);
                } else {
#pragma unroll
                    for (int k = 0; k < CP; ++k) pri[k] = 0.f;
                }
            }
            mbar_wait_t<800>(acc_full(par), ((uint32_t)t >> 1) & 1, prof, dbg[0]);
            tc_fence_after();
            uint32_t dv[32];
            tc_ld32(tmem_base + lane_base + kAccCol0 + (uint32_t)par * 32, dv);
            uint32_t av[8];
            tc_ld8(tmem_base + lane_base + kApartCol0 + (uint32_t)par * 8, av);
            tc_wait_ld();
            float a_tot = 0.f;
#pragma unroll
            for (int b = 0; b < 8; ++b) a_tot += b < NB ? __uint_as_float(av[b]) : 0.f;
            tc_fence_before();
            mbar_arrive(acc_empty(par));
            float d2[CP];
#pragma unroll
            for (int k = 0; k < CP; ++k) d2[k] = (a_tot + __ldg(bias + k)) + __uint_as_float(dv[k]);
            finish_pixel<CP, WANT_DIST>(p, C, d2, tile * kTilePixels, et, out_stage, st, pri);
        }
        // fixed-order reduction of the statistics over the four epilogue warps
        {
            float v[kStatSlots] = {st.proto_conf, st.prior_conf, st.pl_conf, (float)st.pl_pixels, (float)st.pixels,
                                   st.entropy, 0.f, 0.f};
            const int ew = et >> 5;
#pragma unroll
            for (int s = 0; s < kStatSlots; ++s) {
                const float xs = warp_sum(v[s]);
                if (lane == 0) red[ew * kStatSlots + s] = xs;
            }
            named_bar_sync(2, 128);
            if (et < kStatSlots) {
                float xs = 0.f;
                for (int w = 0; w < 4; ++w) xs += red[w * kStatSlots + et];
                p.stat_partials[(size_t)blockIdx.x * kStatSlots + et] = xs;
            }
        }
    } else if (SUMS && warp >= kTcSortWarp0 && warp < kTcSortWarp0 + 2) {
        // =========================== sorter ================================================
        // Per tile: class of every pixel = first argmax of the EMA logits (prototype_handler.py:83-86), then a stable
        // counting sort of the 128 pixels by class (padding pixels form bucket 32, last).  Two warps, two pixels per
        // lane: "virtual warp" v = 2*sw + r owns pixels 32*v .. 32*v+31.  Published per tile parity: row offset and
        // class of every sorted entry, and the cuts of the sorted order into four ranges at class boundaries (one
        // range per warp of a worker group).
        const int sw = warp - kTcSortWarp0;
        float lv[2][CP];
        auto fetch_logits = [&](int t) {   // the logits of tile t+1 are fetched while tile t is being sorted
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const long long n = ((long long)blockIdx.x + (long long)t * gridDim.x) * kTilePixels + 32 * (2 * sw + r) + lane;
                if (t < my_tiles && n < p.N) load_pixel_row<CP>(p.logits, C, HW, n, lv[r]);
            }
        };
        fetch_logits(0);
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int t = 0; t < my_tiles; ++t) {
            const int par = t & 1;
            const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
            int y[2], bucket[2];
            unsigned peers[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int v = 2 * sw + r;
                const long long n = tile * kTilePixels + 32 * v + lane;
                y[r] = (n < p.N) ? first_argmax<CP>(lv[r], C) : -1;
                bucket[r] = y[r] < 0 ? 32 : y[r];
                peers[r] = __match_any_sync(0xffffffffu, bucket[r]);
                wc[v * 36 + lane] = 0;
                if (lane < 4) wc[v * 36 + 32 + lane] = 0;
                __syncwarp();
                if ((peers[r] & lt_mask) == 0) wc[v * 36 + bucket[r]] = __popc(peers[r]);   // lowest lane of each class present
            }
            fetch_logits(t + 1);
            named_bar_sync(1, 64);
            // lane l: size of class l over the tile, then an exclusive prefix over classes
            const int tot = wc[lane] + wc[36 + lane] + wc[72 + lane] + wc[108 + lane];
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int vv = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += vv;
            }
            const int n_valid = __shfl_sync(0xffffffffu, incl, 31);
            const int cstart = incl - tot;
            const bool has = tot > 0 && lane < C;
            // the ranges of the four worker warps start at entries 32, 64, 96: cutcls[c] = the class that runs across
            // entry 32*c | the worker warp holding that class's first entry << 8, or -1 when a class starts there
            int cutcls[4];
#pragma unroll
            for (int c = 1; c < 4; ++c) {
                const bool inside = has && cstart < 32 * c && 32 * c < cstart + tot;
                const unsigned who = __ballot_sync(0xffffffffu, inside);
                const int val = __shfl_sync(0xffffffffu, lane | ((cstart >> 5) << 8), who ? __ffs(who) - 1 : 0);
                cutcls[c] = who ? val : -1;
            }
            if (t >= 2) mbar_wait_t<800>(sort_free(par), (((uint32_t)t >> 1) - 1) & 1, prof, dbg[0]);   // workers are done with tile t-2
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int v = 2 * sw + r;
                const int cs = __shfl_sync(0xffffffffu, cstart, bucket[r] & 31);      // first entry and size of this pixel's class
                const int ct_ = __shfl_sync(0xffffffffu, tot, bucket[r] & 31);
                int base = bucket[r] == 32 ? n_valid : cs;
                for (int v2 = 0; v2 < v; ++v2) base += wc[v2 * 36 + bucket[r]];
                const int pos = base + __popc(peers[r] & lt_mask);
                const bool valid = bucket[r] != 32;
                const bool seg_end = valid && (pos == cs + ct_ - 1 || (pos & 31) == 31);      // last of its class, or of its range
                const bool in_head = valid && cs < (pos & ~31);                                // the class began in an earlier range
                eoff[par * kTilePixels + pos] = ((32 * v + lane) * (kTcTRow * 4)) | (seg_end ? 1 : 0);   // row offset (even) | end flag
                ecls[par * kTilePixels + pos] = in_head ? -1 : y[r] * NB * 256;                // accumulator row offset (bytes), -1 = the range's head
            }
            if (sw == 0) {
                if (lane < 4) cuts[par * 8 + lane] = lane == 0 ? n_valid : (lane == 1 ? cutcls[1] : (lane == 2 ? cutcls[2] : cutcls[3]));
                if (lane < C) cnt[lane] += tot;                  // pixel counts per class
            }
            mbar_arrive(sort_ready(par));
            named_bar_sync(1, 64);                               // wc is reused by the next tile
        }
    }

    // ---- teardown: publish the class partials, release tensor memory
    if (PROF && p.debug != nullptr && lane == 0) {
        long long* d = p.debug + ((size_t)blockIdx.x * 32 + warp) * 8;
        dbg[7] = clock64() - t_start;
        for (int i = 0; i < 8; ++i) d[i] = dbg[i];
    }
    tc_fence_before();
    __syncthreads();
    if (SUMS) {
        float* out = p.cta_partials + (size_t)blockIdx.x * sums_floats(C, D);
        const int cd = C * D;
        for (int i = tid; i < 2 * cd; i += kTcThreads) {     // out: [sum | sum of squares][class][D]
            const int stat = i >= cd, rem = i - stat * cd, k = rem / D, j = rem - k * D;
            out[i] = acc[((k * NB + (j >> 5)) * 2 + stat) * 32 + (j & 31)];
        }
        if (tid < C) out[(size_t)2 * C * D + tid] = (float)cnt[tid];
    }
    if (warp == kTcMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------
bool tc_supported(int B, int D, int HW, int C) {
    if ((unsigned long long)B * D * HW >= (1ull << 32)) return false;            // 32-bit element offsets in the workers
    if (!(D % 128 == 0 && D >= 128 && D <= 256 && C >= 1 && C <= 32)) return false;   // D/64 block pairs: 2 or 4
    return tc_smem(D, C, padded_classes(C), true).total <= 227 * 1024;   // per-CTA shared-memory limit on sm_100
}

int tc_grid(int tiles, int sms) { return tiles < sms ? tiles : sms; }

template <int CP, int CE, bool SUMS, bool WANT_DIST, bool PROF>
static int launch_tc(const FusedParams& p, int grid, cudaStream_t stream) {
    auto kern = fused_tc_kernel<CP, CE, SUMS, WANT_DIST, PROF>;
    const size_t smem = tc_smem(p.D, p.C, CP, SUMS).total;
    ONDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timing_begin(stream);
    kern<<<grid, kTcThreads, smem, stream>>>(p);
    timing_end(stream);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

template <int CP, int CE>
static int launch_tc_cp(const FusedParams& p, int grid, bool sums, cudaStream_t stream) {
    const bool dist = p.dist != nullptr;
    if (p.debug != nullptr && sums && !dist) return launch_tc<CP, CE, true, false, true>(p, grid, stream);   // diagnostics build
    if (sums) return dist ? launch_tc<CP, CE, true, true, false>(p, grid, stream) : launch_tc<CP, CE, true, false, false>(p, grid, stream);
    return dist ? launch_tc<CP, CE, false, true, false>(p, grid, stream) : launch_tc<CP, CE, false, false, false>(p, grid, stream);
}

int launch_fused_tc(const FusedParams& p, int grid, bool sums, cudaStream_t stream) {
    if (p.C == 19) return launch_tc_cp<20, 19>(p, grid, sums, stream);       // the class count of every OnDA config
    return padded_classes(p.C) == 20 ? launch_tc_cp<20, 0>(p, grid, sums, stream) : launch_tc_cp<32, 0>(p, grid, sums, stream);
}

}  // namespace onda
