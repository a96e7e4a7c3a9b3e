// tcgen05 implementation of the fused pass for sm_100a (D % 64 == 0, C <= 32; widths above 256 and small batches run as
// channel slices of at most 256 channels per CTA whose partial dot products split_finish_kernel adds up).
//
// Why tensor cores: the CUDA-core kernel (fused_simt.cu) is issue-bound -- 21 FFMA per feature
// element for the 19-wide contraction alone (profiles/r1_simt_v1_ncu_full.txt) -- so the distance
// contraction moves to the 5th-generation tensor cores with an error-compensated TF32 split:
//     x' . Q  =  hi(x').hi(Q) + hi(x').lo(Q) + lo(x').hi(Q)      (fp32 accumulate in TMEM; the dropped
// lo.lo term is 2^-22 relative)
//
// Feature ingest (round 2): tensor TMA.  The NCHW channel planes are only 4-byte aligned (H*W is odd) and a tensor map
// needs 16-byte aligned strides -- which the planes of every FOURTH channel have (4*H*W floats apart).  So the feature
// map is described by four 3-D tensor maps, one per residue r = channel mod 4: {float index u, channel group a =
// channel / 4, image}, strides {4, 16*H*W, 4*D*H*W} bytes, base = the 16-byte boundary at or below channel r of
// image 0, pixel p of that channel sitting at u = p + shift_r (shift_r = 0..3 floats, fixed per map).  Box
// coordinates are element indices and need no alignment, so ONE cp.async.bulk.tensor fetches the 128-pixel rows of
// 8 channels (box 132 x 8 x 1, 4 extra floats of padding per row, out-of-range elements zero-filled) and four of
// them fetch a 32-channel chunk into a ring stage: rows grouped by residue, slot 8*r + a, 528-byte pitch, the row
// of residue r starting shift_r floats into its slot (boxes start at float index pix0: 16-byte aligned addresses).  No thread issues a global load for the features, no registers
// hold loads in flight, and the memory-level parallelism is the ring depth (6 stages of 16.5 KB), not the warp
// count.  A tile is 128 consecutive pixels of ONE image (the last tile of an image is partial).
//
// One persistent CTA per SM, 32 warps, over a single chunk sequence (chunk = 128 pixels x 32 channels; a tile is D/32
// consecutive chunks; chunk q uses ring stage q % nstage -- nstage is even -- and TMEM A stage q % 4):
//   warp   31    producer: one thread; publishes the CTA's tile schedule in a small shared-memory queue (fixed
//                round-robin, or drawn from a global counter two steps ahead: onda_set_tile_schedule), waits for the
//                ring stage to be free, then issues the chunk's four tensor copies (mbarrier expect_tx / complete_tx).
//   warps  0-7   converters, two groups of four; warp w%4 is the pixel quarter (the only TMEM lanes a warp may
//                touch are 32*(w%4)..+31), group g takes chunks g, g+2, ...  Per chunk: lane = pixel reads the
//                32 channel rows of the stage (consecutive lanes = consecutive words: conflict-free), releases
//                the stage, centres, accumulates sum_j w_j x'_j^2, splits into TF32 hi/lo with packed f32x2 math
//                and writes both with tcgen05.st into TMEM as the A operand (lane = pixel, column = channel).
//   warps  8-23  summers, four groups of four; group g takes the chunk PAIRS g, g+4, .. of every tile: the class sums
//                of both chunks straight from their two (neighbouring) ring stages, lane = channel of either chunk, the
//                two values of an entry travelling through the packed f32x2 pipe together.  With slot 8*(c%4) + c/4,
//                528-byte pitch and the row of residue c%4 starting shift(c%4) floats into its slot, the 32 lanes of a
//                "same pixel, 32 channels" read hit 32 different banks when H*W is odd (bank = 4*(c/4) + shift(c%4) +
//                pixel; the four shifts are then distinct; an even H*W costs bank conflicts here, nothing else).  Each
//                warp walks 32 entries of the class-sorted pixel list and adds every segment's (sum, sum of squares) to
//                that class's shared-memory accumulators; a range that starts inside a class parks that first segment
//                in a spare block which the warp that started the class adds after the group's barrier.  One
//                writer per accumulator at a time, fixed summation order, no atomics.
//   warps 24-27  epilogue: tcgen05.ld of the accumulator columns and the partial columns of their pixel,
//                then the common per-pixel tail (epilogue.cuh): sqrt, softmax, prior rectification, label,
//                statistics.
//   warps 28-29  sorter: per tile, class of every pixel (first argmax of the EMA logits) and a stable
//                counting sort of the 128 pixels by class, published for the summers (double-buffered).
//   warp   30    MMA issuer: per chunk 4 K-steps x 2 tcgen05.mma (kind::tf32, M=128, K=8): hi(x') against the hi and
//                lo rows of B at once (N=64) and lo(x') against the hi rows (N=32); A from TMEM (an MMA this narrow
//                is bound by reading A, so the two hi terms share one read), B = TF32 split of -2*w*(P-mu) resident
//                in shared memory (one bulk copy in the prologue that only this warp waits for).
// The register file is re-split after the prologue (setmaxnreg): converters 72, summers 56, epilogue 80, the rest 64.
// Inside a group of warps ONE warp watches the mbarriers (try_wait) and the others sleep in the group's named barrier.
// A parity wait is only sound on a barrier the waiter is at most one phase away from: every mbarrier here has one
// producer side and consumers that visit it phase by phase -- except the ring's "full" barriers, of whose fills a
// summer group consumes only some: there the watchers publish fill counts and check them first (see the summers).
// Two accumulator buffers of 96 TMEM columns; tcgen05.commit frees A stages / publishes accumulators.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <utility>

#include "epilogue.cuh"

namespace onda {

// Warp roles are aligned to warpgroups (four warps) because the register budget is re-split per warpgroup
// (setmaxnreg): 1024 threads start with 64 registers each.
constexpr int kTcConvGroups = 2;
constexpr int kTcConvWarps = 4 * kTcConvGroups;  // warps 0..7    (warpgroups 0-1)
constexpr int kTcSumWarp0 = 8;                   // warps 8..23   (warpgroups 2-5)
constexpr int kTcSumGroups = 4;
constexpr int kTcSumWarps = 4 * kTcSumGroups;
constexpr int kTcEpiWarp0 = 24;                  // warps 24..27  (warpgroup 6); a multiple of 4: warp w may touch TMEM lanes 32*(w%4)..+31
constexpr int kTcSortWarp0 = 28;                 // warps 28..29: two sorter warps, two pixels per lane
constexpr int kTcMmaWarp = 30;
constexpr int kTcProdWarp = 31;
constexpr int kTcThreads = 32 * 32;
constexpr int kTcRegsConv = 72, kTcRegsSum = 56, kTcRegsEpi = 80;   // x 256 / 512 / 128 threads, + 128 x 64 for the last warpgroup = 65536
constexpr int kTcAStages = 4;                    // TMEM A stages (hi | lo, 64 columns each)
constexpr int kTcChunkC = 32;                    // channels per chunk
constexpr int kTcRowBytes = 528;                 // 16-byte cover of 128 floats at any 4-byte phase
constexpr int kTcStageBytes = kTcChunkC * kTcRowBytes;
constexpr int kTcMaxStages = 8;
constexpr int kTcTileQ = 16;                    // tile queue slots (a power of two, more than the roles can drift apart)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol0 = kTcAStages * 64;   // accumulators after the A stages
constexpr uint32_t kAccSet = 96;                 // per tile parity: three accumulators of 32 columns (one per term of the split, so
                                                 // that consecutive MMAs do not wait for each other's accumulator)
constexpr uint32_t kApartCol0 = kAccCol0 + 2 * kAccSet;   // then sum_j w_j x'_j^2 of each chunk: 2 tile parities x 8 chunks, lane = pixel
constexpr int kTcHeadFloats = kTcSumGroups * 2 * 3 * 2 * kTcChunkC;   // per statistic: head partials of ranges 1..3 of every group's current chunk pair, double-buffered over the group's pairs
constexpr uint32_t kSpinLimit = 4000000u;        // failed probes (each up to ~1 us of hardware suspension) before giving up

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// One probe of an mbarrier phase; the hardware may suspend the warp until the phase completes or a time limit passes.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000u) : "memory");
    return ok != 0;
}
// Wait for a phase.  Every wake-up costs issue slots the working warps need (and any mbarrier event of the CTA
// wakes a suspended warp), so inside a group of warps only ONE warp waits on the mbarrier and the others sleep in
// the group's named barrier -- see the call sites.  A stuck pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t tries = 0;
    while (!mbar_try(bar, parity))
        if (++tries > kSpinLimit) __trap();
}
// Wait of a role with slack (MMA issuer, producer, sorter): plain probes a fixed time apart.  A suspended try_wait is
// woken by mbarrier traffic and every wake-up is a handful of instructions in issue slots the working warps need;
// these roles can afford up to `ns` of extra latency per wait instead.
__device__ __forceinline__ void mbar_wait_slack(uint32_t bar, uint32_t parity, unsigned ns) {
    uint32_t tries = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(ns);
        if (++tries > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_slack_t(uint32_t bar, uint32_t parity, unsigned ns, bool prof, long long& acc_cycles) {
    if (!prof) { mbar_wait_slack(bar, parity, ns); return; }
    const long long t0 = clock64();
    mbar_wait_slack(bar, parity, ns);
    acc_cycles += clock64() - t0;
}
// wait that adds its duration to a diagnostic counter when profiling is on
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool prof, long long& acc_cycles) {
    if (!prof) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc_cycles += clock64() - t0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint64_t v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v)); }
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok));
    return ok != 0;
}
// D[tmem] (+)= A[tmem] . B[smem descriptor]^T, kind::tf32, issued by one thread; ACC = 0 overwrites D
template <int ACC>
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "n"(ACC) : "memory");
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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]);
// the first CP (20 or 32) columns of an accumulator
template <int CP>
__device__ __forceinline__ void tc_ld_cols(uint32_t taddr, uint32_t (&v)[CP]) {
    if constexpr (CP == 32) {
        tc_ld32(taddr, v);
    } else {
        static_assert(CP == 20, "padded class count is 20 or 32");
        tc_ld16(taddr, v);
        tc_ld4(taddr + 16, v + 16);
    }
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
// 3-D tensor copy global -> shared (tile mode), completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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

// slot of channel row j (0..31) of a chunk inside a ring stage: rows grouped by (j mod 4), see the header
__host__ __device__ constexpr int ring_slot(int j) { return 8 * (j & 3) + (j >> 2); }

// ---- shared-memory carve-up ------------------------------------------------------------------------
struct TcSmem {
    size_t bars, tmem_ptr, tq, bias, eoff, ecls, cuts, wc, cnt, red, out, mu, w, btab, acc, ring, total;  // byte offsets
    int nstage;
};
__host__ __device__ inline TcSmem tc_smem(int Dc, int C, int CP, bool sums, int nstage) {
    const int BR = C <= 24 ? 24 : 32;        // rows of the B tables (TableLayout::BR)
    // Everything whose size does not depend on D or C comes first: its addresses are compile-time offsets from the
    // shared-memory base and cost no registers.
    TcSmem s;
    size_t o = 0;
    s.bars = o; o += 512;                              // 2 * kTcMaxStages + 2 * kTcAStages + 9 + kTcTileQ mbarriers
    s.tmem_ptr = o; o += 64;
    s.tq = o; o += 64;                                 // tile queue: the tile of this CTA's t-th step, slot t % kTcTileQ (-1: no more)
    s.eoff = o; o += (size_t)2 * kTilePixels * 4;      // per tile parity: row offset of every class-sorted entry | segment-end flag
    s.ecls = o; o += (size_t)2 * kTilePixels * 4;      // ... and its accumulator row (class * Dc; -1 = head segment of its range)
    s.cuts = o; o += 128;                              // ... [0] live entries, [1..3] the classes cut by the range starts 32, 64, 96
    s.wc = o; o += 640;                                // per-warp class histograms of the sorter (4 x 36)
    s.cnt = o; o += 128;
    s.red = o; o += 4 * kStatSlots * 4;
    s.bias = o; o += 128;                              // per-class bias of the distance (TableLayout::off_bias)
    s.out = o; o += (size_t)kTilePixels * (CP + 1) * 4;
    o = (o + 127) / 128 * 128;
    s.mu = o; o += (size_t)Dc * 4;
    s.w = o; o += (size_t)Dc * 4;
    s.btab = o; o += (size_t)2 * BR * Dc * 4;
    s.acc = o; o += sums ? (size_t)2 * (C * Dc + kTcHeadFloats) * 4 : 0;   // blocks [class * Dc/64 + chunk pair | then 3 heads x 2 buffers per group] of [sum | sum of squares][32 lanes][chunk of the pair]
    o = (o + 127) / 128 * 128;
    s.ring = o; o += (size_t)nstage * kTcStageBytes;
    s.total = o;
    s.nstage = nstage;
    return s;
}
// as many ring stages as fit under the per-CTA limit (227 KB): an even number (the summers take the chunks in pairs
// whose stages must be neighbours), at least 4, at most kTcMaxStages; 0 = does not fit
__host__ inline int tc_ring_stages(int Dc, int C, int CP, bool sums) {
    const size_t fixed = tc_smem(Dc, C, CP, sums, 0).total;
    const size_t limit = 227 * 1024;
    if (fixed + 4 * (size_t)kTcStageBytes > limit) return 0;
    size_t n = (limit - fixed) / kTcStageBytes;
    n = n > (size_t)kTcMaxStages ? (size_t)kTcMaxStages : n;
    return (int)(n & ~(size_t)1);
}

// the four tensor maps of the feature map (one per channel residue mod 4) and each map's shift
struct TcMaps {
    CUtensorMap map[4];
    int shift[4];
};

template <int CP, int CE, bool SUMS, bool WANT_DIST, bool PROF, bool PARTIAL>
__global__ void __launch_bounds__(kTcThreads, 1) fused_tc_kernel(const FusedParams p, const __grid_constant__ TcMaps maps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const long long t_entry = PROF ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = CE > 0 ? CE : p.C;      // CE: class count known at compile time (19 in every OnDA config) -> no k < C predication
    const int D = p.D, HW = p.HW;
    const int Dc = p.slice_channels;      // channels this CTA contracts over
    const int NB = Dc / kTcChunkC;        // even (tc_supported)
    const int nstage = p.nstage;
    const int c_base = (int)blockIdx.y * Dc;     // first channel of this CTA's slice (gridDim.y slices of Dc channels; PARTIAL when > 1)
    const TcSmem L = tc_smem(Dc, C, CP, SUMS, nstage);
    float* Btab = reinterpret_cast<float*>(smem_raw + L.btab);
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
    const uint32_t ring = smem_u32(smem_raw + L.ring);
    const uint32_t bars = smem_u32(smem_raw + L.bars);
    auto ring_full = [&](int s) { return bars + 8u * s; };
    auto ring_empty = [&](int s) { return bars + 8u * (kTcMaxStages + s); };
    constexpr int kB0 = 2 * kTcMaxStages;
    auto full_a = [&](int s) { return bars + 8u * (kB0 + s); };
    auto empty_a = [&](int s) { return bars + 8u * (kB0 + kTcAStages + s); };
    auto acc_full = [&](int i) { return bars + 8u * (kB0 + 2 * kTcAStages + i); };
    auto acc_empty = [&](int i) { return bars + 8u * (kB0 + 2 * kTcAStages + 2 + i); };
    auto sort_ready = [&](int i) { return bars + 8u * (kB0 + 2 * kTcAStages + 4 + i); };
    auto sort_free = [&](int i) { return bars + 8u * (kB0 + 2 * kTcAStages + 6 + i); };
    const uint32_t btab_bar = bars + 8u * (kB0 + 2 * kTcAStages + 8);
    auto tq_full = [&](int s) { return bars + 8u * (kB0 + 2 * kTcAStages + 9 + s); };
    volatile int* tq = reinterpret_cast<volatile int*>(smem_raw + L.tq);

    const TableLayout T = table_layout(C, D);
    // ---- one-time setup: barriers, tensor memory, zeroed accumulators -- this CTA's own shared / tensor memory only --
    // then (pdl_wait: the kernel may have been scheduled while its predecessor was still running) the tables
    float* bias_s = reinterpret_cast<float*>(smem_raw + L.bias);
    if (SUMS) {
        for (int i = tid; i < 2 * (C * Dc + kTcHeadFloats); i += kTcThreads) acc[i] = 0.f;
        if (tid < 32) cnt[tid] = 0;
        if (tid < 16) cuts[16 + tid] = 0;       // landed[] of the summers' watchers
    }
    if (tid == 0) {
        mbar_init(btab_bar, 1);
        for (int s = 0; s < nstage; ++s) {
            mbar_init(ring_full(s), 1);                  // the producer's expect_tx arrive; the four tensor copies complete the bytes
            mbar_init(ring_empty(s), SUMS ? 8 : 4);      // the four converter warps and the four summer warps of the chunk
        }
        for (int s = 0; s < kTcAStages; ++s) {
            mbar_init(full_a(s), 4);                     // one arrival per converter warp
            mbar_init(empty_a(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), 4);                  // one arrival per epilogue warp
            mbar_init(sort_ready(i), 2);                 // one arrival per sorter warp
            mbar_init(sort_free(i), 4 * (NB / 2 < kTcSumGroups ? NB / 2 : kTcSumGroups));   // the summer warps that have chunk pairs
        }
        for (int i = 0; i < kTcTileQ; ++i) mbar_init(tq_full(i), 1);
        fence_barrier_init();
    }
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    if (tid == 0) {
        // B operand table of this CTA's channel slice: one bulk asynchronous copy, off everybody's critical path -- only
        // the MMA issuer waits for it, before its first MMA
        const uint32_t bytes = (uint32_t)(2 * T.BR * Dc) * 4u;
        mbar_arrive_tx(btab_bar, bytes);
        bulk_g2s_plain(smem_u32(Btab), p.table + T.off_b + (size_t)c_base * 2 * T.BR, bytes, btab_bar);
    }
    if (tid < 32) bias_s[tid] = tid < C ? p.table[T.off_bias + tid] : 0.f;
    for (int i = tid; i < Dc; i += kTcThreads) {
        mus[i] = -p.table[T.off_mu + c_base + i];      // negated: the converters centre with a packed add
        wsm[i] = p.table[T.off_w + c_base + i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr bool prof = PROF;       // per-warp wait counters (onda_debug_set_buffer); compiled out of the production kernel
    long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = PROF ? clock64() : 0;
    const unsigned tpi = (unsigned)p.tiles_per_img;
    const unsigned HWu = (unsigned)HW;
    // The producer publishes the CTA's tile schedule in a small queue (fixed round-robin by default, optionally drawn
    // from a global counter).  The tile of this CTA's t-th step, or -1 when there is none:
    auto tile_at = [&](int t) -> int {
        mbar_wait(tq_full(t & (kTcTileQ - 1)), ((uint32_t)t / kTcTileQ) & 1);
        return tq[t & (kTcTileQ - 1)];
    };
    // tile -> image, first pixel, live pixels
    auto tile_of = [&](int tile_, unsigned& img, unsigned& pix0, int& npx) {
        const unsigned tile = (unsigned)tile_;
        img = tile / tpi;
        pix0 = (tile - img * tpi) * kTilePixels;
        const unsigned rem = HWu - pix0;
        npx = rem < (unsigned)kTilePixels ? (int)rem : kTilePixels;
    };
    // Each warpgroup first takes its share of the register file (the releases let the requests through).
    if (warp < kTcConvWarps) {
        // =========================== converters ========================================
        reg_inc<kTcRegsConv>();
        const int quarter = warp & 3, group = warp >> 2;
        const int cbar = 7 + group;          // named barrier of the group (128 threads)
        const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
        const uint32_t lane_word = 4u * (uint32_t)(32 * quarter + lane);
        int stage = group;                   // ring stage of chunk q = q % nstage, phase (q / nstage) & 1
        uint32_t rphase = 0;
        int t = 0, blk = group;
        float x[kTcChunkC];
        for (int q = group;; q += kTcConvGroups) {
            if (blk >= NB) { blk -= NB; ++t; }               // NB is even and >= 2: one tile at most
            if (blk < kTcConvGroups && tile_at(t) < 0) break;     // the group's first chunk of a tile: is there a tile?
            const int par = t & 1;
            const int as = q & (kTcAStages - 1);
            const uint32_t use = (uint32_t)q >> 2;
            if (quarter == 0) {              // one warp of the group watches the mbarriers, the others sleep in the barrier
                mbar_wait_t(ring_full(stage), rphase, prof, dbg[0]);
                // ... including the two the conversion needs (practically always complete by now: the MMAs of four chunks
                // ago, the epilogue of two tiles ago), so the group meets once per chunk
                if (t >= 2) mbar_wait_t(acc_empty(par), (((uint32_t)t >> 1) - 1) & 1, prof, dbg[1]);   // partial columns [par] of tile t-2 consumed
                mbar_wait_t(empty_a(as), (use & 1) ^ 1, prof, dbg[2]);
            }
            named_bar_sync(cbar, 128);
            tc_fence_after();
            {   // the 32 channel rows of this lane's pixel: row j sits in slot 8*(j%4) + j/4 and starts shift[j%4] floats in
                const uint32_t sbase = ring + (uint32_t)stage * kTcStageBytes + lane_word;
#pragma unroll
                for (int j = 0; j < kTcChunkC; ++j) x[j] = lds32(sbase + (uint32_t)ring_slot(j) * kTcRowBytes + 4u * (uint32_t)maps.shift[j & 3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ring_empty(stage));      // this warp's reads of the stage are issued and ordered before the arrive
            stage += kTcConvGroups;
            if (stage >= nstage) { stage -= nstage; rphase ^= 1; }
            const long long t_cv0 = prof ? clock64() : 0;
            uint64_t a2 = 0;                       // sum_j w_j x'_j^2 of the even / odd channels (packed f32x2 math)
            const uint32_t tcol = tmem_base + lane_base + (uint32_t)as * 64;
            const ulonglong2* mu4 = reinterpret_cast<const ulonglong2*>(mus + blk * kTcChunkC);     // -mu, two pairs per load
            const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(wsm + blk * kTcChunkC);
            const uint64_t neg1 = pack2(-1.f, -1.f), vk = pack2(8193.f, 8193.f);
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
                        // Veltkamp split: hi = x' rounded to 11 significant bits (a TF32 value), lo = the exact remainder
                        const uint64_t cc = fmul2(xc, vk);                 // x' * (2^13 + 1)
                        const uint64_t h2 = ffma2(ffma2(xc, neg1, cc), neg1, cc);      // cc - (cc - x')
                        const uint64_t l2 = ffma2(h2, neg1, xc);
                        hi[j] = (uint32_t)h2;
                        hi[j + 1] = (uint32_t)(h2 >> 32);
                        lo[j] = (uint32_t)l2;
                        lo[j + 1] = (uint32_t)(l2 >> 32);
                    }
                }
                tc_st8(tcol + part * 8, hi);
                tc_st8(tcol + 32 + part * 8, lo);
            }
            const float a = __uint_as_float((uint32_t)a2) + __uint_as_float((uint32_t)(a2 >> 32));
            tc_st1(tmem_base + lane_base + kApartCol0 + (uint32_t)(par * 8 + blk), __float_as_uint(a));     // read by this pixel's epilogue thread
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(as));      // one arrival per warp: 32 times fewer barrier events to wake the waiters
            if (prof) dbg[3] += clock64() - t_cv0;         // centring / splitting / tcgen05.st / wait::st
            blk += kTcConvGroups;
        }
    } else if (warp < kTcSumWarp0 + kTcSumWarps) {
        // =========================== summers ===========================================
        reg_dec<kTcRegsSum>();
        if (SUMS) {
        // Class sums of a PAIR of chunks (2 x 32 channels; lane = channel of either chunk), read straight from the two
        // ring stages: the pair shares the walk over the sorted list -- offsets, address arithmetic, segment tests -- and
        // its two values per entry go through the packed f32x2 pipe together.  The class-sorted pixels of the tile are
        // cut into four ranges of 32 entries, one per warp of the group: balanced whatever the label map looks like.
        // The pixel offsets of the entries come four at a time from warp-uniform (broadcast) 16-byte loads, sixteen
        // feature loads are in flight at once, and the segment ends are a warp-uniform bit mask: an end adds the
        // running (sum, sum of squares) to the segment's accumulator block.  A range that starts inside a class
        // accumulates that first segment (its "head") into a spare block -- same code, the sorter just marks those
        // entries -- which the warp that started the class adds to the class block after the group's barrier, heads
        // in range order.  One writer per accumulator at a time, fixed order, no atomics.
        // Accumulator block of (class, pair): [sum | sum of squares][32 lanes][chunk 0 | chunk 1] = 512 bytes.
        const int sw = warp - kTcSumWarp0;
        const int quarter = sw & 3, group = sw >> 2;        // quarter = which 32 sorted entries; group g takes pairs g, g+4, .. of a tile
        const int gbar = 3 + group;                         // named barrier of the group (128 threads)
        const int NP = NB >> 1;                             // chunk pairs per tile (NB is even)
        int stage = 0, qprev = 0;                           // ring stage of the pair's first chunk: even (the ring has an even number of stages)
        int hp = 0;                                         // head buffer of this pair (alternates over the group's pairs)
        int use = 0;                                        // how many times that stage has been filled before (phase = use & 1)
        volatile int* landed = cuts + 16;                   // per stage: fills seen complete by their consumers' watchers
        const uint32_t lane_off = (uint32_t)ring_slot(lane) * kTcRowBytes + 4u * (uint32_t)maps.shift[lane & 3];   // pixel 0 of this lane's channel row
        const int idx = 32 * quarter + lane;                // this warp's range: entries 32*quarter .. +31
        for (int t = 0; tile_at(t) >= 0; ++t)
        for (int pr = group; pr < NP; pr += kTcSumGroups, hp ^= 1) {
            {   // ring stage of chunk q = t * NB + 2 * pr
                const int q = t * NB + 2 * pr;
                stage += q - qprev;
                qprev = q;
                while (stage >= nstage) { stage -= nstage; ++use; }
            }
            const int par = t & 1;
            if (quarter == 0) {              // one warp of the group watches the mbarriers, the others sleep in the barrier
                mbar_wait_t(sort_ready(par), ((uint32_t)t >> 1) & 1, prof, dbg[1]);
                // A parity wait is only sound on a barrier that is at most one phase away, and this group consumes only
                // some of the fills of a stage (the others belong to other groups).  So every watcher publishes the
                // fills it has seen complete (landed[stage] = fills so far), and a watcher first makes sure the fill
                // before its own has been seen -- which never waits longer than the data itself: its own fill cannot
                // even be requested before that one's consumers have released the stage.
                if (use > 0) {
                    long long spins = 0;
                    while (landed[stage] < use || landed[stage + 1] < use) {
                        __nanosleep(64);
                        if (++spins > kSpinLimit) __trap();
                    }
                }
                mbar_wait_t(ring_full(stage), (uint32_t)use & 1u, prof, dbg[0]);
                mbar_wait_t(ring_full(stage + 1), (uint32_t)use & 1u, prof, dbg[0]);
                if (lane == 0) {
                    landed[stage] = use + 1;
                    landed[stage + 1] = use + 1;
                }
            }
            named_bar_sync(gbar, 128);
            const long long t_seg0 = prof ? clock64() : 0;
            const uint32_t row_lane = ring + (uint32_t)stage * kTcStageBytes + lane_off;     // second chunk: + kTcStageBytes
            const uint32_t eo = smem_u32(eoff + par * kTilePixels + 32 * quarter);     // byte offsets of this range's 32 pixels in a row
            const int* ct = cuts + par * 8;
            const uint32_t a1 = smem_u32(acc) + (uint32_t)(pr * 512 + lane * 8);      // + block offset of the class (bytes, from the sorter)
            const uint32_t head_base = (uint32_t)(C * NP + (group * 2 + hp) * 3 - pr) * 512u;   // head j of this group's pair, relative to a1: + (j-1)*512
            int nlive = ct[0] - 32 * quarter;                                         // ct[0] = live entries of the tile
            nlive = nlive < 0 ? 0 : (nlive > 32 ? 32 : nlive);
            const int er_i = ecls[par * kTilePixels + idx];                           // bit 0: segment end, bit 1: head segment, else class block offset
            const unsigned endbits = __ballot_sync(0xffffffffu, er_i & 1);            // segment ends (sorter: class end or entry 31)
            const uint32_t myrow = (er_i & 2) ? head_base + (uint32_t)(quarter - 1) * 512u : (uint32_t)(er_i & ~3);   // accumulator block (bytes) of this entry's segment
            // The accumulators of the running segment are fetched when the segment starts, so a flush is two packed adds
            // and two stores; entries past the last live one need no guard: that one ends a segment, so whatever they add
            // to the running sums is never flushed (their rows are the zero-filled padding of the tile anyway).
            uint32_t cur = a1 + (uint32_t)__shfl_sync(0xffffffffu, myrow, 0);
            uint64_t c1 = lds64(cur), c2 = lds64(cur + 256u);
            uint64_t s1 = 0ull, s2 = 0ull;            // (chunk 0, chunk 1) running sum / sum of squares
#pragma unroll 1
            for (int h = 0; h < 4; ++h) {            // 8 entries x 2 chunks at a time: sixteen loads in flight (kept rolled: code size)
                if (8 * h >= nlive) break;
                int4 ov[2];
#pragma unroll
                for (int i = 0; i < 2; ++i)          // warp-uniform address: one broadcast 16-byte load gives four offsets
                    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(ov[i].x), "=r"(ov[i].y), "=r"(ov[i].z), "=r"(ov[i].w) : "r"(eo + 32u * h + 16u * i));
                uint64_t xv[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t o0 = row_lane + (uint32_t)ov[i].x, o1 = row_lane + (uint32_t)ov[i].y;
                    const uint32_t o2 = row_lane + (uint32_t)ov[i].z, o3 = row_lane + (uint32_t)ov[i].w;
                    xv[4 * i + 0] = pack2(lds32(o0), lds32(o0 + (uint32_t)kTcStageBytes));
                    xv[4 * i + 1] = pack2(lds32(o1), lds32(o1 + (uint32_t)kTcStageBytes));
                    xv[4 * i + 2] = pack2(lds32(o2), lds32(o2 + (uint32_t)kTcStageBytes));
                    xv[4 * i + 3] = pack2(lds32(o3), lds32(o3 + (uint32_t)kTcStageBytes));
                }
                if (h == 3 || nlive <= 8 * h + 8) {  // the stages' last reads are issued: release them (ordered before the arrives)
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(ring_empty(stage));
                        mbar_arrive(ring_empty(stage + 1));
                    }
                }
                // four entries at a time: without a segment end among the first three, the quad is a small tree (short
                // dependent chains, one test); otherwise entry by entry.  Either way the order of the additions is a
                // function of the sorted list alone.
                const unsigned hb = endbits >> (8 * h);
                auto flush = [&](int e) {            // entry e (0..7 of this step) ends a segment
                    sts64(cur, fadd2(c1, s1));
                    sts64(cur + 256u, fadd2(c2, s2));
                    s1 = 0ull;
                    s2 = 0ull;
                    if (8 * h + e < 31) {            // the next segment's accumulators
                        cur = a1 + (uint32_t)__shfl_sync(0xffffffffu, myrow, 8 * h + e + 1);
                        c1 = lds64(cur);
                        c2 = lds64(cur + 256u);
                    }
                };
#pragma unroll
                for (int qd = 0; qd < 2; ++qd) {
                    const uint64_t x0 = xv[4 * qd], x1 = xv[4 * qd + 1], x2 = xv[4 * qd + 2], x3 = xv[4 * qd + 3];
                    const unsigned qb = (hb >> (4 * qd)) & 0xfu;
                    if ((qb & 7u) == 0u) {
                        s1 = fadd2(s1, fadd2(fadd2(x0, x1), fadd2(x2, x3)));
                        s2 = fadd2(s2, fadd2(ffma2(x1, x1, fmul2(x0, x0)), ffma2(x3, x3, fmul2(x2, x2))));
                        if (qb & 8u) flush(4 * qd + 3);
                    } else {
                        s1 = fadd2(s1, x0); s2 = ffma2(x0, x0, s2);
                        if (qb & 1u) flush(4 * qd);
                        s1 = fadd2(s1, x1); s2 = ffma2(x1, x1, s2);
                        if (qb & 2u) flush(4 * qd + 1);
                        s1 = fadd2(s1, x2); s2 = ffma2(x2, x2, s2);
                        if (qb & 4u) flush(4 * qd + 2);
                        s1 = fadd2(s1, x3); s2 = ffma2(x3, x3, s2);
                        if (qb & 8u) flush(4 * qd + 3);
                    }
                }
            }
            if (nlive == 0) {
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(ring_empty(stage));
                    mbar_arrive(ring_empty(stage + 1));
                }
            }
            if (prof) dbg[3] += clock64() - t_seg0;
            named_bar_sync(gbar, 128);          // all heads of this pair are complete
            {   // move the heads of the classes this warp started into their class blocks
                const int hj = (lane >= 1 && lane < 4) ? ct[lane] : -1;
                unsigned m = __ballot_sync(0xffffffffu, hj >= 0 && (hj >> 8) == quarter);
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int k = __shfl_sync(0xffffffffu, hj, j) & 0xff;
                    const uint32_t hd = a1 + head_base + (uint32_t)(j - 1) * 512u, ad = a1 + (uint32_t)(k * NP) * 512u;
                    sts64(ad, fadd2(lds64(ad), lds64(hd)));
                    sts64(ad + 256u, fadd2(lds64(ad + 256u), lds64(hd + 256u)));
                    sts64(hd, 0ull);
                    sts64(hd + 256u, 0ull);
                }
            }
            // The next pair's segment flushes (possibly into the very same class blocks) and head blocks (the other
            // buffer) start behind the group barrier that opens that pair: every warp has finished its moves by then.
            if (pr + kTcSumGroups >= NP) {      // that was this group's last pair of the tile
                if (lane == 0) mbar_arrive(sort_free(par));
            }
        }
        }   // SUMS
    } else if (warp >= kTcEpiWarp0 && warp < kTcEpiWarp0 + 4) {
        // =========================== epilogue ===========================================
        reg_inc<kTcRegsEpi>();
        const int et = tid - kTcEpiWarp0 * 32;          // 0..127 = pixel row of the tile = TMEM lane
        const uint32_t lane_base = (uint32_t)(et & ~31) << 16;
        PixelStats st;
        for (int t = 0;; ++t) {
            const int tile = tile_at(t);
            if (tile < 0) break;
            const int par = t & 1;
            unsigned img, pix0;
            int npx;
            tile_of(tile, img, pix0, npx);
            const long long n0 = (long long)img * HW + pix0;
            float pri[CP];
            {   // prior row of this pixel: issued before the wait so its latency hides behind the MMAs
                if (!PARTIAL && et < npx && p.prior != nullptr && (p.labels != nullptr || p.soft != nullptr)) {
                    load_pixel_row_at<CP>(p.prior + ((size_t)img * C) * HW + pix0 + et, C, HWu, pri);
                } else {
#pragma unroll
                    for (int k = 0; k < CP; ++k) pri[k] = 0.f;
                }
            }
            if (warp == kTcEpiWarp0) mbar_wait_t(acc_full(par), ((uint32_t)t >> 1) & 1, prof, dbg[0]);
            named_bar_sync(2, 128);
            tc_fence_after();
            // the three terms of the split: hi.hi in columns [0, BR), hi.lo in [BR, 2 BR), lo.hi in [64, 64 + BR): small terms first
            const uint32_t dcol = tmem_base + lane_base + kAccCol0 + (uint32_t)par * kAccSet;
            float d2[CP];
            {
                uint32_t u[CP], v[CP];
                tc_ld_cols<CP>(dcol + (uint32_t)T.BR, u);
                tc_ld_cols<CP>(dcol + 64, v);
                tc_wait_ld();
#pragma unroll
                for (int k = 0; k < CP; ++k) d2[k] = __uint_as_float(u[k]) + __uint_as_float(v[k]);
            }
            uint32_t dv[CP];
            tc_ld_cols<CP>(dcol, dv);
            uint32_t av[8];
            tc_ld8(tmem_base + lane_base + kApartCol0 + (uint32_t)par * 8, av);
            tc_wait_ld();
            float a_tot = 0.f;
#pragma unroll
            for (int b = 0; b < 8; ++b) a_tot += b < NB ? __uint_as_float(av[b]) : 0.f;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(par));
            if (PARTIAL) {
                // this CTA covers one channel slice: park the partial dot products (scaled back by -1/2: B holds -2 Q)
                // and the partial sum_j w_j x'_j^2 for split_finish_kernel, which adds the slices in order
                if (et < npx) {
                    float* dst = p.dots_scratch + ((size_t)blockIdx.y * (CP + 1)) * p.N + (n0 + et);
#pragma unroll
                    for (int k = 0; k < CP; ++k) dst[(size_t)k * p.N] = -0.5f * (d2[k] + __uint_as_float(dv[k]));
                    dst[(size_t)CP * p.N] = a_tot;
                }
            } else {
#pragma unroll
                for (int k = 0; k < CP; ++k) d2[k] = (a_tot + bias_s[k]) + (d2[k] + __uint_as_float(dv[k]));
                finish_pixel_rows<CP, WANT_DIST>(p, C, d2, n0, npx, et, out_stage, st, pri);
            }
        }
        // fixed-order reduction of the statistics over the four epilogue warps
        if (!PARTIAL) {
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
    } else {
    // sorter, MMA issuer and producer share the last warpgroup, which keeps its 64 registers
    if (warp == kTcMmaWarp) {
        // =========================== MMA issuer ========================================
        // The whole warp walks the loop (uniform control flow); one elected lane issues.  Descriptors are the chunk's
        // base descriptor plus a constant per K-step (the address field counts 16-byte units; 8 channels = 64 units).
        {
            const uint32_t idesc64 = make_idesc_tf32(128, 64), idesc32 = make_idesc_tf32(128, 32);
            const uint32_t lbo = 32u * (uint32_t)T.BR;       // bytes between 4-channel slabs of the B table
            const uint64_t b0 = make_bdesc(smem_u32(Btab), lbo, 128);
            const uint32_t kstep = 4u * (uint32_t)T.BR;      // one K-step = 8 channels = two slabs, in the descriptor's 16-byte units
            mbar_wait(btab_bar, 0);             // the B table has landed (bulk copy of the prologue)
            int q = 0;
            for (int t = 0; tile_at(t) >= 0; ++t) {
                const int par = t & 1;
                if (t >= 2) mbar_wait_slack_t(acc_empty(par), (((uint32_t)t >> 1) - 1) & 1, 64, prof, dbg[0]);
                const uint32_t d0 = tmem_base + kAccCol0 + (uint32_t)par * kAccSet, d1 = d0 + 64;
                for (int b = 0; b < NB; ++b, ++q) {
                    const int as = q & (kTcAStages - 1);
                    const uint32_t use = (uint32_t)q >> 2;
                    mbar_wait_slack_t(full_a(as), use & 1, 64, prof, dbg[1]);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_hi = tmem_base + (uint32_t)as * 64, a_lo = a_hi + 32;
                        const uint64_t bd = b0 + (uint64_t)((uint32_t)b * 4u * kstep);
                        // hi(x') against the hi and lo rows at once (64 classes wide: columns [0, BR) and [BR, 2 BR) of d0),
                        // lo(x') against the hi rows (d1); the first MMAs of a tile overwrite the accumulators
                        if (b == 0) {
                            tc_mma_tf32<0>(d0, a_hi, bd, idesc64);
                            tc_mma_tf32<0>(d1, a_lo, bd, idesc32);
                        } else {
                            tc_mma_tf32<1>(d0, a_hi, bd, idesc64);
                            tc_mma_tf32<1>(d1, a_lo, bd, idesc32);
                        }
#pragma unroll
                        for (int ks = 1; ks < 4; ++ks) {
                            tc_mma_tf32<1>(d0, a_hi + ks * 8, bd + (uint64_t)(ks * kstep), idesc64);
                            tc_mma_tf32<1>(d1, a_lo + ks * 8, bd + (uint64_t)(ks * kstep), idesc32);
                        }
                        tc_commit(empty_a(as));             // A stage reusable once these MMAs retire
                        if (b == NB - 1) tc_commit(acc_full(par));      // accumulator complete
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kTcProdWarp) {
        // =========================== producer ===========================================
        // One thread.  Chunk (tile, 32 channels from cb) = four tensor copies, one per channel residue r: box of 132
        // floats x 8 channel groups starting at float index pix0 (a 16-byte aligned address; pixel pix0 lands shift_r
        // floats into the row, indices past the row's end are zero-filled), channel group cb / 4, image img.
        if (lane == 0) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
#pragma unroll
            for (int r = 0; r < 4; ++r) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.map[r]) : "memory");
            int stage = 0;
            uint32_t rphase = 0;
            // Tile schedule, published through the queue one step before anybody needs it (the sorter reads one step
            // ahead).  Default: the first two tiles are fixed (blockIdx.x, blockIdx.x + gridDim.x), later ones are drawn
            // from this slice's counter while the copies of two steps earlier are being issued -- SMs do not run at
            // the same speed and tiles / SMs is rarely an integer: 4-5 % faster at 14.3 tiles per SM.  Which CTA
            // accumulates which tile then varies from run to run, and with it the last bits of the class sums (labels,
            // soft predictions and statistics of the step do not depend on them).  onda_set_tile_schedule(0): the
            // fixed round-robin tile blockIdx.x + t * gridDim.x, bit-reproducible.
            const int G = (int)gridDim.x;
            unsigned* counter = p.sched + blockIdx.y;
            // (not for channel slices: the slices of a tile should run at the same time, and measured 0.8 % slower at D = 2048)
            const bool drawing = p.dynamic_tiles != 0 && p.tiles > 2 * G && gridDim.y == 1;      // else: tile blockIdx.x + t * gridDim.x, as a fixed schedule
            int cur = (int)blockIdx.x, nxt = (int)blockIdx.x + G < p.tiles ? (int)blockIdx.x + G : -1;
            tq[0] = cur;
            tq[1] = nxt;
            mbar_arrive(tq_full(0));
            mbar_arrive(tq_full(1));
            for (int t = 0; cur >= 0; ++t) {
                int after = -1;
                if (nxt >= 0) after = drawing ? 2 * G + (int)atomicAdd(counter, 1u) : nxt + G;
                unsigned img, pix0;
                int npx;
                tile_of(cur, img, pix0, npx);
                for (int b = 0; b < NB; ++b) {
                    mbar_wait_slack_t(ring_empty(stage), rphase ^ 1, 100, prof, dbg[0]);
                    const uint32_t dst = ring + (uint32_t)stage * kTcStageBytes;
                    const uint32_t bar = ring_full(stage);
                    mbar_arrive_tx(bar, (uint32_t)kTcStageBytes);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        tma_load_3d(dst + (uint32_t)(8 * r) * kTcRowBytes, &maps.map[r], (int)pix0, (c_base + b * kTcChunkC) / 4,
                                    (int)img, bar, policy);
                    if (++stage == nstage) { stage = 0; rphase ^= 1; }
                }
                if (after >= p.tiles) after = -1;
                tq[(t + 2) & (kTcTileQ - 1)] = after;
                mbar_arrive(tq_full((t + 2) & (kTcTileQ - 1)));
                cur = nxt;
                nxt = after;
            }
            // This CTA has drawn its last tile.  The last producer to get here zeroes the counters for the next launch
            // (launches that never draw -- every tile assigned statically -- skip the ticket).
            if (drawing) {
                __threadfence();
                if (atomicAdd(p.sched + kTcSchedSlices, 1u) == gridDim.x * gridDim.y - 1) {
                    __threadfence();
                    for (unsigned y = 0; y < gridDim.y; ++y) p.sched[y] = 0u;
                    p.sched[kTcSchedSlices] = 0u;
                }
            }
        }
        __syncwarp();
    } else if (SUMS && warp >= kTcSortWarp0 && warp < kTcSortWarp0 + 2) {
        // =========================== sorter ================================================
        // Per tile: class of every pixel = first argmax of the EMA logits (prototype_handler.py:83-86), then a stable
        // counting sort of the 128 pixels by class (padding pixels form bucket 32, last).  Two warps, two pixels per
        // lane: "virtual warp" v = 2*sw + r owns pixels 32*v .. 32*v+31.  Published per tile parity: pixel offset and
        // class of every sorted entry, and the cuts of the sorted order into four ranges at class boundaries (one
        // range per warp of a summer group).
        const int sw = warp - kTcSortWarp0;
        float lv[2][CP];
        auto fetch_logits = [&](int t) {   // the logits of tile t+1 are fetched while tile t is being sorted
            const int tile = tile_at(t);
            if (tile < 0) return;
            unsigned img, pix0;
            int npx;
            tile_of(tile, img, pix0, npx);
            const float* lp = p.logits + ((size_t)img * C) * HW + pix0 + 64 * sw + lane;
            if (64 * sw + lane < npx) load_pixel_row_at<CP>(lp, C, HWu, lv[0]);
            if (64 * sw + 32 + lane < npx) load_pixel_row_at<CP>(lp + 32, C, HWu, lv[1]);
        };
        fetch_logits(0);
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int t = 0;; ++t) {
            const int tile = tile_at(t);
            if (tile < 0) break;
            const int par = t & 1;
            unsigned img, pix0;
            int npx;
            tile_of(tile, img, pix0, npx);
            int y[2], bucket[2];
            unsigned peers[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int v = 2 * sw + r;
                y[r] = (32 * v + lane < npx) ? first_argmax_fast<CP>(lv[r], C) : -1;
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
            // the ranges of the four summer warps start at entries 32, 64, 96: cutcls[c] = the class that runs across
            // entry 32*c | the summer warp holding that class's first entry << 8, or -1 when a class starts there
            int cutcls[4];
#pragma unroll
            for (int c = 1; c < 4; ++c) {
                const bool inside = has && cstart < 32 * c && 32 * c < cstart + tot;
                const unsigned who = __ballot_sync(0xffffffffu, inside);
                const int val = __shfl_sync(0xffffffffu, lane | ((cstart >> 5) << 8), who ? __ffs(who) - 1 : 0);
                cutcls[c] = who ? val : -1;
            }
            if (t >= 2) {                                                                    // summers are done with tile t-2
                if (sw == 0) mbar_wait_slack_t(sort_free(par), (((uint32_t)t >> 1) - 1) & 1, 200, prof, dbg[0]);
                named_bar_sync(1, 64);
            }
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
                eoff[par * kTilePixels + pos] = (32 * v + lane) * 4;                          // byte offset of the pixel in a channel row
                // accumulator row offset of the entry's class (bytes, a multiple of 256) | bit 1: the class began in an
                // earlier range (head segment) | bit 0: segment end
                ecls[par * kTilePixels + pos] = (valid ? y[r] * NB * 256 : 0) | (in_head || !valid ? 2 : 0) | (seg_end ? 1 : 0);
            }
            if (sw == 0) {
                if (lane < 4) cuts[par * 8 + lane] = lane == 0 ? n_valid : (lane == 1 ? cutcls[1] : (lane == 2 ? cutcls[2] : cutcls[3]));
                if (lane < C) cnt[lane] += tot;                  // pixel counts per class
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sort_ready(par));
            named_bar_sync(1, 64);                               // wc is reused by the next tile
        }
    }

    }   // last warpgroup

    // ---- teardown: publish the class partials, release tensor memory
    const long long t_role_end = PROF ? clock64() : 0;
    tc_fence_before();
    __syncthreads();
    if (tid == 0) pdl_launch_dependents();      // the partial combine may be scheduled; it waits for this grid to complete
    const long long t_sync_end = PROF ? clock64() : 0;
    if (SUMS) {
        float* out = p.cta_partials + (size_t)blockIdx.x * sums_floats(C, D);
        // accumulator row r = (class * NP + pair) * 2 + statistic: 32 lanes x (chunk 2 pair | chunk 2 pair + 1)
        // -> out: [sum | sum of squares][class][D]
        const int NP = NB >> 1;
        const unsigned inv_np = (65536u + (unsigned)NP - 1u) / (unsigned)NP;      // kc / NP = (kc * inv_np) >> 16 for kc < 2^13
        for (int r = warp; r < 2 * C * NP; r += kTcThreads / 32) {
            const int stat = r & 1, kc = r >> 1;
            const int k = (int)(((unsigned)kc * inv_np) >> 16), pr = kc - k * NP;
            const float2 v = reinterpret_cast<const float2*>(acc)[r * 32 + lane];
            float* dst = out + (size_t)(stat * C + k) * D + c_base + pr * 64 + lane;
            dst[0] = v.x;
            dst[32] = v.y;
        }
        if (tid < C && blockIdx.y == 0) out[(size_t)2 * C * D + tid] = (float)cnt[tid];
    }
    if (warp == kTcMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (PROF && p.debug != nullptr && lane == 0) {
        long long* d = p.debug + ((size_t)blockIdx.x * 32 + warp) * 8;
        dbg[7] = t_role_end - t_start;
        dbg[4] = t_start - t_entry;                 // prologue
        dbg[5] = t_sync_end - t_role_end;           // waiting for the slowest role
        dbg[6] = clock64() - t_sync_end;            // writing the partials / releasing tensor memory
        for (int i = 0; i < 8; ++i) d[i] = dbg[i];
    }
}

// ---- host side -----------------------------------------------------------------------------------------
// channels per CTA: the whole width up to 256, else the largest of 256 / 192 / 128 / 64 that divides D (a tile is then
// split over D / slice CTAs whose partial dot products meet in split_finish_kernel)
int tc_slice_channels(int D) {
    if (D % 64 != 0 || D < 64) return 0;
    if (D <= 256) return D;
    for (int dc = 256; dc >= 64; dc -= 64)
        if (D % dc == 0) return dc;
    return 0;
}

bool tc_supported(int B, int D, int HW, int C) {
    (void)B; (void)HW;
    const int dc = tc_slice_channels(D);
    if (dc == 0 || D / dc > 16 || C < 1 || C > 32) return false;
    return tc_ring_stages(dc, C, padded_classes(C), true) >= 4;
}

int tc_tiles(int B, int HW) { return B * ((HW + kTilePixels - 1) / kTilePixels); }

// Channel slices of a launch: D / tc_slice_channels(D).  (Halving the slice once more for small batches -- 66 tiles at
// B = 1 -- was measured: the kernel drops from 20 to 16 us but the finishing launch it needs costs 5 us: not done.)
int tc_slices(int tiles, int D, int sms) {
    (void)tiles; (void)sms;
    return D / tc_slice_channels(D);
}

int tc_grid(int tiles, int sms, int slices) {
    const int per_slice = sms / slices > 0 ? sms / slices : 1;
    return tiles < per_slice ? tiles : per_slice;
}

// ---- tensor maps of the feature map (see the header): built with the driver's cuTensorMapEncodeTiled, fetched through
// the runtime so the library does not link libcuda; the last few encodings are kept (a training step alternates
// between a handful of feature buffers).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tc_make_maps(const float* feat, int B, int D, int HW, TcMaps* out) {
    struct Entry { const float* feat; int B, D, HW, dev; TcMaps maps; };
    constexpr int kEntries = 8;
    static Entry cache[kEntries];
    static int used = 0, next = 0;
    static std::mutex mu;
    static EncodeTiledFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    const int dev = current_device();
    for (int i = 0; i < used; ++i) {
        const Entry& e = cache[i];
        if (e.feat == feat && e.B == B && e.D == D && e.HW == HW && e.dev == dev) { *out = e.maps; return ONDA_OK; }
    }
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        ONDA_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        ONDA_REQUIRE(q == cudaDriverEntryPointSuccess && fn != nullptr, "tcgen05 kernel: the driver does not export cuTensorMapEncodeTiled");
        encode = (EncodeTiledFn)fn;
    }
    Entry e;
    e.feat = feat; e.B = B; e.D = D; e.HW = HW; e.dev = dev;
    for (int r = 0; r < 4; ++r) {
        const uintptr_t first = reinterpret_cast<uintptr_t>(feat + (size_t)r * HW);       // channel r of image 0, pixel 0
        const uintptr_t base = first & ~(uintptr_t)15;
        const int shift = (int)((first - base) >> 2);
        const cuuint64_t dims[3] = {(cuuint64_t)HW + (cuuint64_t)shift, (cuuint64_t)(D / 4), (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)16 * HW, (cuuint64_t)4 * D * HW};
        const cuuint32_t box[3] = {kTcRowBytes / 4, 8, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult rc = encode(&e.maps.map[r], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, reinterpret_cast<void*>(base), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ONDA_REQUIRE(rc == CUDA_SUCCESS, "tcgen05 kernel: cuTensorMapEncodeTiled failed (%d) for B=%d D=%d HW=%d", (int)rc, B, D, HW);
        e.maps.shift[r] = shift;
    }
    cache[next] = e;
    next = (next + 1) % kEntries;
    if (used < kEntries) ++used;
    *out = e.maps;
    return ONDA_OK;
}

template <int CP, int CE, bool SUMS, bool WANT_DIST, bool PROF, bool PARTIAL>
static int launch_tc(FusedParams p, int grid, cudaStream_t stream) {
    auto kern = fused_tc_kernel<CP, CE, SUMS, WANT_DIST, PROF, PARTIAL>;
    p.nstage = tc_ring_stages(p.slice_channels, p.C, CP, SUMS);
    const size_t smem = tc_smem(p.slice_channels, p.C, CP, SUMS, p.nstage).total;
    static size_t smem_set[kMaxDevices] = {};     // per instantiation and device: raise the attribute only when a launch needs more
    const int dev = current_device();
    if (smem > smem_set[dev]) {
        ONDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev] = smem;
    }
    p.dynamic_tiles = tile_schedule_dynamic() ? 1 : 0;
    TcMaps maps;
    const int rc = tc_make_maps(p.feat, p.B, p.D, p.HW, &maps);
    if (rc != ONDA_OK) return rc;
    timing_begin(stream);
    const cudaError_t le = launch_chained(kern, dim3((unsigned)grid, (unsigned)(p.D / p.slice_channels)), dim3(kTcThreads), smem, stream, p, maps);
    timing_end(stream);
    ONDA_CUDA_TRY(le);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

template <int CP, int CE>
static int launch_tc_cp(const FusedParams& p, int grid, bool sums, cudaStream_t stream) {
    const bool dist = p.dist != nullptr;
    if (p.slice_channels < p.D)       // channel slices: the per-pixel tail runs in split_finish_kernel
        return sums ? launch_tc<CP, CE, true, false, false, true>(p, grid, stream) : launch_tc<CP, CE, false, false, false, true>(p, grid, stream);
    if (p.debug != nullptr && sums && !dist) return launch_tc<CP, CE, true, false, true, false>(p, grid, stream);   // diagnostics build
    if (sums) return dist ? launch_tc<CP, CE, true, true, false, false>(p, grid, stream) : launch_tc<CP, CE, true, false, false, false>(p, grid, stream);
    return dist ? launch_tc<CP, CE, false, true, false, false>(p, grid, stream) : launch_tc<CP, CE, false, false, false, false>(p, grid, stream);
}

int launch_fused_tc(const FusedParams& p, int grid, bool sums, cudaStream_t stream) {
    if (p.C == 19) return launch_tc_cp<20, 19>(p, grid, sums, stream);       // the class count of every OnDA config
    return padded_classes(p.C) == 20 ? launch_tc_cp<20, 0>(p, grid, sums, stream) : launch_tc_cp<32, 0>(p, grid, sums, stream);
}

}  // namespace onda
