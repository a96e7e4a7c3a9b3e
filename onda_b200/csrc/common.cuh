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
bool tile_schedule_dynamic();      // onda_set_tile_schedule / ONDA_TC_DYNAMIC_TILES
// optional event bracket around the dominant kernel (api.cu); both are no-ops unless timing is enabled
void timing_begin(cudaStream_t s);
void timing_end(cudaStream_t s);
// per-device caches (kernel attributes, SM counts) are indexed by the current device, clamped to kMaxDevices - 1
constexpr int kMaxDevices = 64;
int current_device();
int cached_sm_count();

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

// ---- programmatic dependent launch ------------------------------------------------------------
// The kernels of a step (fused pass -> partial combine -> EMA/table -> next fused pass) are launched with the
// programmatic-stream-serialization attribute: a kernel may be SCHEDULED while its predecessor is still running -- once
// every CTA of the predecessor has executed pdl_launch_dependents() or exited -- and runs its private prologue
// (shared/tensor memory only) until pdl_wait(), which returns when the predecessor has completed and its writes are
// visible.  That hides the launch latency of three kernels per step.  Rules kept by every kernel launched this way:
// nothing before pdl_wait() reads or writes global memory; every CTA executes pdl_wait() (so completion is transitive
// along the chain).  A kernel whose predecessor does not take part simply starts when that one has finished.
bool pdl_enabled();      // api.cu: on unless ONDA_PDL=0
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- geometry ----------------------------------------------------------------
constexpr int kTilePixels = 128;     // pixels per CTA tile (= 4 warps x 32 lanes = 128 TMEM lanes)
constexpr int kStatSlots = ONDA_NUM_STATS;

__host__ __device__ inline int padded_classes(int C) { return C <= 20 ? 20 : 32; }
__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Distance-table layout (floats), built by onda_build_distance_table:
//   sigma[Dp] | w[Dp] | mu[Dp] | bias[32] | Q[Dp][CP]  (channel-major, CP = padded_classes(C))
//   | B[Dp/4][2*BR][4]   TF32 split of -2*Q laid out as the tcgen05 B operand: K-major, no swizzle, 8x16-byte
//     core matrices.  Per 4-channel slab: BR rows of hi parts, then BR rows of lo parts; element (class n, channel c)
//     of part s (0 = hi, 1 = lo) sits at float index (c/4)*8*BR + (s*BR + n)*4 + (c%4), so LBO (next 4 channels) =
//     32*BR bytes and SBO (next 8 classes) = 128 B.  BR = 24 rows when C <= 24 (else 32).  One 64-class-wide MMA on
//     a slab start therefore yields hi.hi in accumulator columns 0..BR-1 and hi.lo in columns BR..2*BR-1 (the rest
//     of its 64 columns is the next slab's beginning: garbage nobody reads), and a 32-wide MMA yields lo.hi.
// Dp = D rounded up to 32; padded channels carry w = 0, Q = 0 so they contribute nothing.
struct TableLayout {
    int C, D, Dp, CP, BR;
    size_t off_sigma, off_w, off_mu, off_bias, off_q, off_b, off_scratch, off_sched, total;
};
constexpr int kTcSchedSlices = 16;      // tile counters of the tcgen05 kernel: one per channel slice, then the exit ticket
__host__ __device__ inline TableLayout table_layout(int C, int D) {
    TableLayout t;
    t.C = C; t.D = D; t.Dp = round_up(D, 32); t.CP = padded_classes(C); t.BR = C <= 24 ? 24 : 32;
    t.off_sigma = 0;
    t.off_w = t.off_sigma + t.Dp;
    t.off_mu = t.off_w + t.Dp;
    t.off_bias = t.off_mu + t.Dp;
    t.off_q = t.off_bias + 32;
    t.off_b = t.off_q + (size_t)t.Dp * t.CP;
    t.off_scratch = t.off_b + (size_t)2 * t.BR * t.Dp + 64;      // (+ 64: the over-read of the last slab stays inside the table); per-CTA bias partials (doubles) + ticket of the table kernel
    t.off_sched = t.off_scratch + (size_t)2 * 32 * (t.Dp / 32) + 8;      // zero between launches (re-armed by the kernel that uses them)
    t.total = t.off_sched + kTcSchedSlices + 8;
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

// ---- peer memory (one-shot all-reduce over NVLink; allreduce.cu, api.cu) ----------------------------
constexpr int kMaxPeers = 8;
struct PeerTable {
    const float* buf[kMaxPeers];      // rank r's input slot, mapped into this process
    uint32_t* flags[kMaxPeers];       // rank r's flag words for this slot: flags[r][src] = epoch when src is ready
    uint32_t* done[kMaxPeers];        // rank r's "done" words (or null): done[r][src] = epoch once src has read r's slot
};
constexpr long long kPeerSpinLimit = 200000000LL;      // probes 256 ns apart: about a minute, then the kernel traps

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

inline PeerTable make_peer_table(int rank, int world, void* const* bufs, void* const* flags, void* const* done = nullptr) {
    PeerTable t;
    for (int r = 0; r < kMaxPeers; ++r) {
        t.buf[r] = (const float*)bufs[r < world ? r : rank];    // unused entries stay dereferenceable and local (peer_gather4)
        t.flags[r] = r < world ? (uint32_t*)flags[r] : nullptr;
        t.done[r] = (r < world && done != nullptr) ? (uint32_t*)done[r] : nullptr;
    }
    return t;
}
// After the last read of the peers' slots: tell every peer "rank has read your slot of this epoch".  One thread per peer.
__device__ __forceinline__ void peer_publish_done(const PeerTable& peers, int rank, int world, uint32_t epoch) {
    uint32_t* done_of_peer = nullptr;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
        if (r == (int)threadIdx.x) done_of_peer = peers.done[r];
    if ((int)threadIdx.x < world && done_of_peer != nullptr) st_release_sys(done_of_peer + rank, epoch);
}
// Handshake of a one-shot exchange: publish "my input is ready" (epoch) on every peer, wait for all peers'.
// Called by every thread of every CTA; CTA 0 publishes.  Ends with __syncthreads().
__device__ __forceinline__ void peer_handshake(const PeerTable& peers, int rank, int world, uint32_t epoch) {
    // constant-index walks of the pointer tables keep them in the kernel parameter bank (no local-memory copy)
    uint32_t* flag_of_peer = nullptr;      // for thread r < world: peer r's flag array
    const uint32_t* my_flags = nullptr;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) {
        if (r == (int)threadIdx.x) flag_of_peer = peers.flags[r];
        if (r == rank) my_flags = peers.flags[r];
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();                                   // this rank's input (written by the previous kernel) first
        st_release_sys(flag_of_peer + rank, epoch);
    }
    if ((int)threadIdx.x < world) {
        const uint32_t* mine = my_flags + threadIdx.x;
        long long spins = 0;
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {         // epochs only grow: a later one also means "ready"
            __nanosleep(256);
            if (++spins > kPeerSpinLimit) __trap();               // a missing peer traps instead of hanging the GPU
        }
    }
    __syncthreads();
}
// sum over the ranks of element i of the exchanged buffers, in rank order: identical on every rank
__device__ __forceinline__ float peer_sum(const PeerTable& peers, int world, size_t i) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
        if (r < world) s += ld_relaxed_sys(peers.buf[r] + i);
    return s;
}

// Gather of a CTA's slice: element groups of four floats (16-byte aligned) at indices gi[0..ITEMS), summed over the
// ranks in rank order.  ALL loads (ITEMS x W, W = world rounded up to 1/2/4/8) are issued before the first add, so
// the slice costs one NVLink round trip.  Entries >= world of the peer table alias this rank's own buffer.
template <int W, int ITEMS>
__device__ __forceinline__ void peer_gather4(const PeerTable& peers, int world, const size_t (&gi)[ITEMS], const bool (&on)[ITEMS],
                                             float4 (&out)[ITEMS]) {
    float4 v[ITEMS][W];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (on[it])
                asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[it][r].x), "=f"(v[it][r].y), "=f"(v[it][r].z), "=f"(v[it][r].w) : "l"(peers.buf[r] + gi[it]));
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (on[it] && r < world) { s.x += v[it][r].x; s.y += v[it][r].y; s.z += v[it][r].z; s.w += v[it][r].w; }
        out[it] = s;
    }
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
