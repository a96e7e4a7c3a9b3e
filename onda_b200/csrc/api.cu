// extern "C" entry points of include/onda_b200.h plus the small kernels around the fused pass:
// distance-table build, fixed-order partial reduction, EMA / cumulative prototype updates and
// the prior-mix / switch-statistics kernel.
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "epilogue.cuh"

namespace onda {

// ---- error plumbing -----------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return ONDA_ECUDA;
}

static unsigned long long g_launches = 0;
// Tile schedule of the tcgen05 kernel: 1 = tiles drawn from a device counter (default; fastest), 0 = fixed round-robin
// (class sums bit-reproducible from run to run).  ONDA_TC_DYNAMIC_TILES=0 in the environment changes the default.
static int g_dynamic_tiles = [] { const char* e = getenv("ONDA_TC_DYNAMIC_TILES"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
bool tile_schedule_dynamic() { return g_dynamic_tiles != 0; }
static long long* g_debug = nullptr;       // optional device buffer for cycle / time stamps (onda_debug_set_buffer)
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// ---- optional kernel timing (roofline report) ----------------------------------------------
constexpr int kMaxTimed = 4096;
static std::mutex g_timing_mu;      // the diagnostics below may be driven from several host threads
static int g_timing = 0;            // 0 = off, n > 0 = bracket every n-th launch of the dominant kernel
static int g_timing_seen = 0;
static int g_timed = 0;
static cudaEvent_t g_ev[kMaxTimed][2];
static int g_ev_made = 0;
static thread_local int t_open_slot = -1;      // the event pair this thread's current launch is bracketed by

void timing_begin(cudaStream_t s) {
    t_open_slot = -1;
    std::lock_guard<std::mutex> lock(g_timing_mu);
    if (g_timing <= 0 || g_timed >= kMaxTimed) return;
    if ((g_timing_seen++ % g_timing) != 0) return;
    while (g_ev_made <= g_timed) {
        cudaEventCreate(&g_ev[g_ev_made][0]);
        cudaEventCreate(&g_ev[g_ev_made][1]);
        ++g_ev_made;
    }
    t_open_slot = g_timed++;
    cudaEventRecord(g_ev[t_open_slot][0], s);
}
void timing_end(cudaStream_t s) {
    if (t_open_slot < 0) return;
    cudaEventRecord(g_ev[t_open_slot][1], s);
    t_open_slot = -1;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("ONDA_PDL");
        return !(e != nullptr && e[0] == '0');
    }();
    return on;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
    return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int cached_sm_count() {
    static int sms[kMaxDevices] = {};
    const int dev = current_device();
    if (sms[dev] > 0) return sms[dev];
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    sms[dev] = v;
    return v;
}

// ---- distance table (optionally fused with the EMA blend) ---------------------------------------
// One 128-thread CTA per 32 channels: lane = channel, warp g handles the classes k = g, g+4, g+8, ... of that
// channel (four short dependent chains instead of one long one); per-channel partial moments meet in shared
// memory and are added in warp order.  Every quantity of the table is per channel except the per-class bias,
// whose per-CTA partial sums are folded by the last CTA to finish, in CTA order (deterministic).
// With `sums` != null the moving-average blend of ma() (prototype_handler.py:88-99) is applied to P and S first,
// by the same thread that then rebuilds the table entries of that (class, channel).
constexpr int kTableThreads = 128;

__global__ void __launch_bounds__(kTableThreads) table_kernel(float* __restrict__ P, float* __restrict__ S,
                                                              const float* __restrict__ cnt, int C, int D, int metric,
                                                              float* __restrict__ table, const float* __restrict__ sums,
                                                              float lam, PeerTable peers, int rank, int world,
                                                              uint32_t epoch, float* __restrict__ sums_out,
                                                              uint32_t* __restrict__ epoch_counter,
                                                              long long* __restrict__ stamps) {
    // world > 0: `sums` of every rank sit in peer-mapped slots; this kernel is also the all-reduce -- handshake,
    // then every use of sums[i] is the rank-ordered sum over the peers (identical on every rank), and the reduced
    // buffer is written to sums_out for the statistics readers.
    pdl_launch_dependents();      // the next fused pass may be scheduled (its prologue touches no global memory)
    pdl_wait();                   // the partial combine has completed
    const TableLayout T = table_layout(C, D);
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const bool mahal = metric == ONDA_METRIC_MAHALANOBIS;
    const bool gather = world > 0;
    __shared__ float cn[32];               // pixel count of every class (the EMA's cnt_k)
    // the epoch of this exchange: a host argument, or (graph replay: arguments are frozen) a device counter that the
    // last CTA of this grid bumps for the next call
    // optional time stamps of CTA 0 (exchange breakdown, profiles/exchange_probe.py): start | handshake done | gather done | end
    auto stamp = [&](int i) {
        if (stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            stamps[i] = t;
        }
    };
    stamp(0);
    if (gather && epoch_counter != nullptr) epoch = *reinterpret_cast<volatile uint32_t*>(epoch_counter);
    if (gather) peer_handshake(peers, rank, world, epoch);
    stamp(1);
    __shared__ float gs[2 * 32 * 32];      // gather: the reduced (sum | sum of squares)[class][this CTA's 32 channels]
    if (gather) {
        // all threads fetch the CTA's slice of every rank's sums with independent (16-byte when possible) loads:
        // one NVLink round trip per thread instead of one per class
        const int j0 = blockIdx.x * 32;
        if (D % 32 == 0) {
            constexpr int kItems = 2 * ONDA_MAX_CLASSES * 8 / kTableThreads;      // float4 groups per thread (4)
            size_t gi[kItems];
            bool on[kItems];
            float4 v[kItems];
#pragma unroll
            for (int it = 0; it < kItems; ++it) {
                const int item = threadIdx.x + it * kTableThreads;
                on[it] = item < 2 * C * 8;
                gi[it] = (size_t)(item >> 3) * D + j0 + (item & 7) * 4;            // row = stat * C + class
            }
            if (world <= 2) peer_gather4<2>(peers, world, gi, on, v);
            else if (world <= 4) peer_gather4<4>(peers, world, gi, on, v);
            else peer_gather4<8>(peers, world, gi, on, v);
#pragma unroll
            for (int it = 0; it < kItems; ++it) {
                const int item = threadIdx.x + it * kTableThreads;
                if (on[it]) {
                    *reinterpret_cast<float4*>(gs + (item >> 3) * 32 + (item & 7) * 4) = v[it];
                    *reinterpret_cast<float4*>(sums_out + gi[it]) = v[it];
                }
            }
        } else {
            for (int item = threadIdx.x; item < 2 * C * 32; item += kTableThreads) {
                const int row = item >> 5, jj = item & 31;
                if (j0 + jj < D) {
                    const size_t gi = (size_t)row * D + j0 + jj;
                    const float v = peer_sum(peers, world, gi);
                    gs[row * 32 + jj] = v;
                    sums_out[gi] = v;
                }
            }
        }
    }
    if (sums != nullptr || gather) {
        const size_t tail0 = (size_t)2 * C * D;
        if ((int)threadIdx.x < C + kStatSlots) {
            const float v = gather ? peer_sum(peers, world, tail0 + threadIdx.x) : sums[tail0 + threadIdx.x];
            if ((int)threadIdx.x < C) cn[threadIdx.x] = v;
            if (gather && blockIdx.x == 0) sums_out[tail0 + threadIdx.x] = v;
        }
    }
    if (gather) __syncthreads();
    stamp(2);
    __shared__ double wk[32];              // c_k / sum_k c_k
    __shared__ double mom[4][3][32];       // per warp: partial (gm, gsq, mean) of each channel
    __shared__ double bpart[4][32];        // per warp: partial bias of its classes
    if (g == 0) {
        double ck = (mahal && lane < C) ? (double)cnt[lane] : 0.0, total = ck;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        wk[lane] = ck / total;
    }
    __syncthreads();
    float pk[8];                            // this thread's classes g, g+4, ... (at most 8 of 32)
    double gm = 0.0, gsq = 0.0, mean = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = g + 4 * i;
        pk[i] = 0.f;
        if (j < D && k < C) {
            float pv = P[(size_t)k * D + j];
            float sv = (mahal || sums != nullptr || gather) ? S[(size_t)k * D + j] : 0.f;
            if (sums != nullptr || gather) {     // ma(): P <- P*rho + (1-rho)*sum/max(cnt,1), rho = lambda if cnt > 0 else 1
                const float n = cn[k];
                const float rho = n > 0.f ? lam : 1.f;
                const float one_m = __fsub_rn(1.f, rho);
                const float den = n > 0.f ? n : 1.f;
                const float sx = gather ? gs[k * 32 + lane] : sums[(size_t)k * D + j];
                const float sxx = gather ? gs[(C + k) * 32 + lane] : sums[(size_t)(C + k) * D + j];
                pv = __fadd_rn(__fmul_rn(pv, rho), __fmul_rn(one_m, __fdiv_rn(sx, den)));
                sv = __fadd_rn(__fmul_rn(sv, rho), __fmul_rn(one_m, __fdiv_rn(sxx, den)));
                P[(size_t)k * D + j] = pv;
                S[(size_t)k * D + j] = sv;
            }
            pk[i] = pv;
            mean += (double)pv;
            if (mahal) {
                gm += (double)pv * wk[k];          // global_var(): prototype_handler.py:57-59
                gsq += (double)sv * wk[k];         // :54-56
            }
        }
    }
    mom[g][0][lane] = gm;
    mom[g][1][lane] = gsq;
    mom[g][2][lane] = mean;
    __syncthreads();
    gm = (mom[0][0][lane] + mom[1][0][lane]) + (mom[2][0][lane] + mom[3][0][lane]);
    gsq = (mom[0][1][lane] + mom[1][1][lane]) + (mom[2][1][lane] + mom[3][1][lane]);
    mean = (mom[0][2][lane] + mom[1][2][lane]) + (mom[2][2][lane] + mom[3][2][lane]);
    float sigma = 0.f, wf = 0.f, muf = 0.f;
    if (j < D) {
        if (mahal) {
            sigma = (float)sqrt(gsq - gm * gm);        // :60
            wf = (float)(1.0 / ((double)sigma * (double)sigma));
            muf = (float)gm;
        } else {
            sigma = 1.f;
            wf = 1.f;
            muf = (float)(mean / C);
        }
    }
    if (g == 0 && j < T.Dp) {
        table[T.off_sigma + j] = sigma;
        table[T.off_w + j] = wf;
        table[T.off_mu + j] = muf;
    }
    // table entries and bias of this thread's classes
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = g + 4 * i;
        float qf = 0.f;
        double bk = 0.0;
        if (j < D && k < C) {
            const double diff = (double)pk[i] - (double)muf;
            qf = (float)((double)wf * diff);
            bk = (double)wf * diff * diff;
        }
        if (j < T.Dp) {
            if (k < T.CP) table[T.off_q + (size_t)j * T.CP + k] = qf;
            // TF32 split of -2*Q in the tcgen05 B-operand layout (common.cuh); classes >= C are zero rows
            const float v = -2.f * qf;
            // (both parts rounded to nearest TF32 here: the tensor core would truncate the low 13 bits)
            const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
            if (k < T.BR) {
                const size_t bi = (size_t)(j >> 2) * (8 * T.BR) + (size_t)k * 4 + (j & 3);
                table[T.off_b + bi] = hi;
                table[T.off_b + bi + 4 * T.BR] = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & 0xffffe000u);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bk += __shfl_xor_sync(0xffffffffu, bk, o);   // over the CTA's 32 channels
        if (lane == 0) bpart[g][i] = bk;
    }
    __syncthreads();
    // per-class bias: this CTA's partial, then the last CTA folds all CTAs in order
    double* part = reinterpret_cast<double*>(table + T.off_scratch);
    unsigned* ticket = reinterpret_cast<unsigned*>(table + T.off_scratch + (size_t)2 * 32 * gridDim.x);
    __shared__ unsigned last_flag;
    if (threadIdx.x < 32) part[(size_t)blockIdx.x * 32 + threadIdx.x] = bpart[threadIdx.x & 3][threadIdx.x >> 2];   // class k = g + 4i
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (last_flag && threadIdx.x < 32) {
        __threadfence();
        double v = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) v += __ldcg(part + (size_t)c * 32 + threadIdx.x);
        table[T.off_bias + threadIdx.x] = (float)v;
        if (threadIdx.x == 0) {
            *ticket = 0u;        // re-arm for the next build
            if (gather && epoch_counter != nullptr) *epoch_counter = epoch + 1u;     // every CTA has read it by now
        }
    }
    if (last_flag && gather) peer_publish_done(peers, rank, world, epoch);          // every CTA's peer reads are complete
    stamp(3);
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

__global__ void class_std_kernel(const float* __restrict__ P, const float* __restrict__ S, float* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sqrtf(__fsub_rn(S[i], __fmul_rn(P[i], P[i])));  // prototype_var(): prototype_handler.py:49-51
}

// ---- fixed-order combine of per-CTA partials ----------------------------------------------
// Block = 32 elements x 8 partial-lanes; lane g sums CTAs g, g+8, ... then the 8 lane sums are
// added in lane order: the summation tree depends only on (n_cta), never on scheduling.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ cta_partials, int n_cta,
                                                              size_t stride, int class_elems, int has_sums,
                                                              const float* __restrict__ stat_partials, int n_stat,
                                                              int has_stats, float* __restrict__ out,
                                                              const uint32_t* __restrict__ peer_done,
                                                              const uint32_t* __restrict__ peer_epoch, int peer_world) {
    __shared__ double part[8][32];
    pdl_launch_dependents();      // the EMA/table kernel may be scheduled now; it waits for this grid to complete
    pdl_wait();                   // the fused pass has completed
    if (peer_done != nullptr && (int)threadIdx.x < peer_world) {     // `out` is peer-visible: its previous contents must have
        const uint32_t want = *peer_epoch - 1u;                      // been read by every rank (one flag per thread: one round trip)
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(peer_done + threadIdx.x) - want) < 0) {
            __nanosleep(200);
            if (++spins > 200000000u) __trap();
        }
    }
    const int el = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + el;
    const int total = class_elems + kStatSlots;
    double s = 0.0;
    if (e < class_elems) {
        if (has_sums) {
            const float* src = cta_partials + e;
#pragma unroll 8
            for (int c = g; c < n_cta; c += 8) s += (double)__ldcg(src + (size_t)c * stride);   // independent loads, one add chain
        }
    } else if (e < total) {
        if (has_stats)
            for (int c = g; c < n_stat; c += 8) s += (double)stat_partials[(size_t)c * kStatSlots + (e - class_elems)];
    }
    part[g][el] = s;
    __syncthreads();
    if (g == 0 && e < total) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][el];
        out[e] = (float)t;
    }
}

// ---- prototype updates -----------------------------------------------------------------------
__global__ void ema_kernel(float* __restrict__ P, float* __restrict__ S, const float* __restrict__ sums, int C, int D,
                           float lam) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * D) return;
    const int k = i / D;
    const float cnt = sums[(size_t)2 * C * D + k];
    const float rho = cnt > 0.f ? lam : 1.f;        // lambda ** (cnt > 0), prototype_handler.py:92
    const float one_m = __fsub_rn(1.f, rho);
    const float den = cnt > 0.f ? cnt : 1.f;        // mask(), :21, :93
    const float m1 = __fdiv_rn(sums[i], den);
    const float m2 = __fdiv_rn(sums[(size_t)C * D + i], den);
    P[i] = __fadd_rn(__fmul_rn(P[i], rho), __fmul_rn(one_m, m1));  // :94-96
    S[i] = __fadd_rn(__fmul_rn(S[i], rho), __fmul_rn(one_m, m2));  // :97-99
}

__global__ void append_kernel(float* __restrict__ P, float* __restrict__ S, float* __restrict__ counter,
                              const float* __restrict__ sums, int C, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * D) return;
    const int k = i / D;
    const float cnt = sums[(size_t)2 * C * D + k];
    const float tot = __fadd_rn(counter[k], cnt);   // counter += sums, prototype_handler.py:66
    const float den = tot > 0.f ? tot : 1.f;        // :67
    const float d1 = __fsub_rn(sums[i], __fmul_rn(P[i], cnt));                  // :71
    const float d2 = __fsub_rn(sums[(size_t)C * D + i], __fmul_rn(S[i], cnt));  // :72
    P[i] = __fadd_rn(P[i], __fdiv_rn(d1, den));     // :73
    S[i] = __fadd_rn(S[i], __fdiv_rn(d2, den));     // :74
}

__global__ void counter_add_kernel(float* __restrict__ counter, const float* __restrict__ sums, int C, int D) {
    const int k = threadIdx.x;
    if (k < C) counter[k] = __fadd_rn(counter[k], sums[(size_t)2 * C * D + k]);
}

// ---- prior mix + switch statistics ----------------------------------------------------------
constexpr int kPriorThreads = 256;

template <int CP>
__global__ void __launch_bounds__(kPriorThreads, CP == 20 ? 4 : 2) prior_mix_kernel(const float* __restrict__ l0, const float* __restrict__ l1,
                                                                     const float* __restrict__ l2, float c0, float c1, float c2,
                                                                     float scale01, int B, int C, int HW, float* __restrict__ prior_out,
                                                                     float* __restrict__ partials, unsigned* __restrict__ ticket,
                                                                     float* __restrict__ stats_out) {
    __shared__ float red[kPriorThreads / 32][kStatSlots];
    __shared__ bool last;
    const unsigned N = (unsigned)B * (unsigned)HW, hw = (unsigned)HW;       // 32-bit indices (the host checks B*C*HW < 2^32)
    const float* src[3] = {l0, l1, l2};
    const float coef[3] = {c0, c1, c2};
    float st[kStatSlots];
#pragma unroll
    for (int s = 0; s < kStatSlots; ++s) st[s] = 0.f;
    for (unsigned n = blockIdx.x * kPriorThreads + threadIdx.x; n < N; n += gridDim.x * kPriorThreads) {
        const unsigned b = n / hw, q = n - b * hw;
        const unsigned off = (b * (unsigned)C) * hw + q;
        float mix[CP];
#pragma unroll
        for (int k = 0; k < CP; ++k) mix[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (src[i] == nullptr) {
                if (i == 1) {
#pragma unroll
                    for (int k = 0; k < CP; ++k) mix[k] = __fmul_rn(mix[k], scale01);
                }
                continue;
            }
            float z[CP];
            float zmax = -__int_as_float(0x7f800000);
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                if (k < C) {
                    z[k] = __ldg(src[i] + off + (unsigned)k * hw);
                    if (z[k] > zmax || z[k] != z[k]) zmax = z[k];
                }
            }
            float esum = 0.f;
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                if (k < C) {
                    z[k] = fast_ex2((z[k] - zmax) * 1.4426950408889634f);    // single MUFU, 2^-22 relative
                    esum += z[k];
                }
            }
            const float inv = __frcp_rn(esum);            // one rounded reciprocal per pixel instead of C divisions (<= 1 ulp apart)
            float pmax = -1.f;
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                if (k < C) {
                    const float pk = __fmul_rn(z[k], inv);    // softmax(axis=1), prototypes_hybrid_switch.py:53
                    if (torch_greater(pk, pmax)) pmax = pk;
                    mix[k] = __fadd_rn(mix[k], __fmul_rn(coef[i], pk));  // prior (+)= lambda * softmax, :56, :64, :84
                }
            }
            st[i] += pmax;                                    // .max(axis=1)[0].mean(), :54, :62, :81
            if (i == 1) {                                     // h-switch: prior *= percentage_static (prototypes_hswitch.py:56)
#pragma unroll
                for (int k = 0; k < CP; ++k) mix[k] = __fmul_rn(mix[k], scale01);
            }
        }
        float mmax = -__int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < CP; ++k) {
            if (k < C) {
                if (torch_greater(mix[k], mmax)) mmax = mix[k];
                if (prior_out != nullptr) prior_out[off + (unsigned)k * hw] = mix[k];
            }
        }
        st[3] += mmax;                                        // {"prior": prior.max(axis=1)[0].mean()}, :88
        st[4] += 1.f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < kStatSlots; ++s) {
        const float v = warp_sum(st[s]);
        if (lane == 0) red[warp][s] = v;
    }
    __syncthreads();
    if (threadIdx.x < kStatSlots) {
        float v = 0.f;
        for (int w = 0; w < kPriorThreads / 32; ++w) v += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * kStatSlots + threadIdx.x] = v;
    }
    // last CTA to arrive folds the per-CTA rows in CTA order (deterministic), then re-arms the ticket
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        __threadfence();
        if (threadIdx.x < kStatSlots) {
            double v = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) v += (double)__ldcg(partials + (size_t)c * kStatSlots + threadIdx.x);
            stats_out[threadIdx.x] = (float)v;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}


// ---- per-step log reductions (prototypes.py:341-352) ---------------------------------------------
// agreement count of the pseudo-labels with the first argmax of the student logits, number of non-ignored labels,
// sum of squared prototype entries.  Integer counts are exact; everything is combined in CTA order by the last CTA.
constexpr int kLogSlots = 4;
template <int CP>
__global__ void __launch_bounds__(kPriorThreads) step_log_kernel(const long long* __restrict__ labels,
                                                                 const float* __restrict__ logits,
                                                                 const float* __restrict__ protos, int B, int C, int HW,
                                                                 int n_proto, double* __restrict__ partials,
                                                                 unsigned* __restrict__ ticket, float* __restrict__ out) {
    __shared__ double red[kPriorThreads / 32][kLogSlots];
    __shared__ bool last;
    const long long N = (long long)B * HW;
    double agree = 0.0, valid = 0.0, sq = 0.0;
    for (long long n = (long long)blockIdx.x * kPriorThreads + threadIdx.x; n < N; n += (long long)gridDim.x * kPriorThreads) {
        float z[CP];
        load_pixel_row<CP>(logits, C, HW, n, z);
        const long long lab = labels[n];
        agree += lab == (long long)first_argmax<CP>(z, C) ? 1.0 : 0.0;      // == out.argmax(axis=1), :346-347
        valid += (lab >= 0 && lab != 255) ? 1.0 : 0.0;                       // :342
    }
    for (int i = blockIdx.x * kPriorThreads + threadIdx.x; i < n_proto; i += gridDim.x * kPriorThreads) {
        const double v = (double)protos[i];
        sq += v * v;                                                         // (prototypes**2).mean(), :351
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v3[kLogSlots] = {agree, valid, sq, 0.0};
#pragma unroll
    for (int s = 0; s < kLogSlots; ++s) {
        double v = v3[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][s] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLogSlots) {
        double v = 0.0;
        for (int w = 0; w < kPriorThreads / 32; ++w) v += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * kLogSlots + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        __threadfence();
        if (threadIdx.x < kLogSlots) {
            double v = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) v += __ldcg(partials + (size_t)c * kLogSlots + threadIdx.x);
            out[threadIdx.x] = threadIdx.x == 3 ? (float)N : (float)v;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}


// ---- model-weight EMA (prototypes.py:407-416) as one multi-tensor launch --------------------------------
// The host cuts every parameter / buffer into chunks of at most kEmaChunkBytes and uploads the table once; a CTA
// walks chunks grid-stride.  mode 0: dst = dst*keep + src*take on floats, two roundings then the add exactly like
// `param_k.clone() * a + param_q.clone() * (1 - a)`; mode 1: byte copy (buffers, any dtype).
constexpr int kEmaThreads = 256;
__global__ void __launch_bounds__(kEmaThreads) weight_ema_kernel(const onda_ema_chunk* __restrict__ chunks, int n_chunks,
                                                                 float keep, float take) {
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const onda_ema_chunk ch = chunks[c];
        const uintptr_t sa = (uintptr_t)ch.src, da = (uintptr_t)ch.dst;
        if (ch.mode == 0) {
            const float* s = (const float*)ch.src;
            float* d = (float*)ch.dst;
            const unsigned n = ch.count;
            if (((sa | da) & 15) == 0) {
                const unsigned n4 = n >> 2;
                const float4* s4 = (const float4*)s;
                float4* d4 = (float4*)d;
                for (unsigned i = threadIdx.x; i < n4; i += kEmaThreads) {
                    const float4 q = __ldcs(s4 + i);
                    float4 k = d4[i];
                    k.x = __fadd_rn(__fmul_rn(k.x, keep), __fmul_rn(q.x, take));
                    k.y = __fadd_rn(__fmul_rn(k.y, keep), __fmul_rn(q.y, take));
                    k.z = __fadd_rn(__fmul_rn(k.z, keep), __fmul_rn(q.z, take));
                    k.w = __fadd_rn(__fmul_rn(k.w, keep), __fmul_rn(q.w, take));
                    d4[i] = k;
                }
                for (unsigned i = (n4 << 2) + threadIdx.x; i < n; i += kEmaThreads)
                    d[i] = __fadd_rn(__fmul_rn(d[i], keep), __fmul_rn(s[i], take));
            } else {
                for (unsigned i = threadIdx.x; i < n; i += kEmaThreads)
                    d[i] = __fadd_rn(__fmul_rn(d[i], keep), __fmul_rn(s[i], take));
            }
        } else {
            const unsigned n = ch.count;
            if (((sa | da) & 15) == 0) {
                const unsigned n16 = n >> 4;
                const uint4* s16 = (const uint4*)ch.src;
                uint4* d16 = (uint4*)ch.dst;
                for (unsigned i = threadIdx.x; i < n16; i += kEmaThreads) d16[i] = s16[i];
                const unsigned char* sb = (const unsigned char*)ch.src;
                unsigned char* db = (unsigned char*)ch.dst;
                for (unsigned i = (n16 << 4) + threadIdx.x; i < n; i += kEmaThreads) db[i] = sb[i];
            } else {
                const unsigned char* sb = (const unsigned char*)ch.src;
                unsigned char* db = (unsigned char*)ch.dst;
                for (unsigned i = threadIdx.x; i < n; i += kEmaThreads) db[i] = sb[i];
            }
        }
    }
}


// ---- evaluation: upsample + argmax + confusion matrix in one pass (adaptation_model.py:143-160) ------------------
// The reference materialises interp(pred) at full resolution (19 x 1024 x 2048 floats per image), softmaxes it, takes the
// argmax, copies it to the host and bincounts there.  Here a thread owns one full-resolution column of a strip of rows:
// it keeps, per class, the two source rows of the bilinear interpolation already interpolated along x (align_corners=
// True, torch's index/weight arithmetic) in registers, so walking down the strip costs two multiplies and an add per
// class and pixel and new loads only when the source row changes (every ~8 pixels).  First argmax (softmax is monotone
// per pixel, so the argmax is that of the interpolated logits), then (label, prediction) is counted in a shared-memory
// histogram; integer atomics only, so the result does not depend on scheduling.  The upsampled tensor never exists.
constexpr int kConfThreads = 256;
template <int CP>
__global__ void __launch_bounds__(kConfThreads) confusion_kernel(const float* __restrict__ logits, int B, int C, int h, int w,
                                                                 const long long* __restrict__ labels, int H, int W,
                                                                 int strip_rows, unsigned long long* __restrict__ hist,
                                                                 unsigned char* __restrict__ pred_out) {
    __shared__ unsigned int sh[ONDA_MAX_CLASSES * ONDA_MAX_CLASSES];
    for (int i = threadIdx.x; i < C * C; i += kConfThreads) sh[i] = 0u;
    __syncthreads();
    const float scale_h = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;      // area_pixel_compute_scale, align_corners
    const float scale_w = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const long long hw = (long long)h * w;
    const int n_colblocks = (W + kConfThreads - 1) / kConfThreads;
    const int n_strips = (H + strip_rows - 1) / strip_rows;
    const long long items = (long long)B * n_strips * n_colblocks;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int cb = (int)(it % n_colblocks);
        const long long t2 = it / n_colblocks;
        const int strip = (int)(t2 % n_strips), b = (int)(t2 / n_strips);
        const int x = cb * kConfThreads + threadIdx.x;
        if (x >= W) continue;
        const float w1r = __fmul_rn(scale_w, (float)x);
        const int w1 = (int)w1r, w1p = w1 < w - 1 ? 1 : 0;
        const float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.f, w1l);
        const float* img = logits + ((long long)b * C) * hw + w1;
        float top[CP], bot[CP];              // rows rowT / rowB of every class, interpolated along x
        int rowT = -1, rowB = -1;
        auto load_row = [&](int r, float (&dst)[CP]) {
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                if (k < C) {
                    const float* p = img + (long long)k * hw + (long long)r * w;
                    dst[k] = __fadd_rn(__fmul_rn(w0l, __ldg(p)), __fmul_rn(w1l, __ldg(p + w1p)));
                }
            }
        };
        const int y_end = (strip + 1) * strip_rows < H ? (strip + 1) * strip_rows : H;
        for (int y = strip * strip_rows; y < y_end; ++y) {
            const float h1r = __fmul_rn(scale_h, (float)y);
            const int h1 = (int)h1r, h1p = h1 < h - 1 ? 1 : 0;
            const float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.f, h1l);
            const int rT = h1, rB = h1 + h1p;
            if (rT != rowT) {
                if (rT == rowB) {
#pragma unroll
                    for (int k = 0; k < CP; ++k) top[k] = bot[k];
                } else {
                    load_row(rT, top);
                }
                rowT = rT;
            }
            if (rB != rowB) {
                if (rB == rowT) {
#pragma unroll
                    for (int k = 0; k < CP; ++k) bot[k] = top[k];
                } else {
                    load_row(rB, bot);
                }
                rowB = rB;
            }
            float best = 0.f;
            int arg = 0;
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                if (k < C) {
                    const float v = __fadd_rn(__fmul_rn(h0l, top[k]), __fmul_rn(h1l, bot[k]));
                    if (k == 0 || torch_greater(v, best)) { best = v; arg = k; }
                }
            }
            const long long n = ((long long)b * H + y) * W + x;
            if (pred_out != nullptr) pred_out[n] = (unsigned char)arg;
            const long long lab = labels[n];
            if (lab >= 0 && lab < C) atomicAdd(&sh[(int)lab * C + arg], 1u);          // fast_hist: func.py:77-79
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += kConfThreads)
        if (sh[i] != 0u) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

}  // namespace onda

namespace onda {
int launch_allreduce_oneshot(float* out, size_t n, int rank, int world, void* const* bufs, void* const* flags,
                             uint32_t epoch, cudaStream_t stream);
}
using namespace onda;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int onda_abi_version(void) { return 1; }

const char* onda_last_error(void) { return g_err; }

int onda_sm_count(void) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return ONDA_ECUDA;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return ONDA_ECUDA;
    return v;
}

unsigned long long onda_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int onda_set_tile_schedule(int dynamic) {
    const int prev = g_dynamic_tiles;
    if (dynamic == 0 || dynamic == 1) g_dynamic_tiles = dynamic;
    return prev;
}

int onda_debug_set_buffer(void* device_buffer) {
    g_debug = (long long*)device_buffer;
    return ONDA_OK;
}

int onda_kernel_timing_enable(int enable) {
    std::lock_guard<std::mutex> lock(g_timing_mu);
    g_timing = enable > 0 ? enable : 0;
    g_timing_seen = 0;
    g_timed = 0;
    return ONDA_OK;
}

int onda_kernel_timing_read(float* total_ms_host, int* launches_host) {
    ONDA_REQUIRE(total_ms_host && launches_host, "onda_kernel_timing_read: null pointer");
    std::lock_guard<std::mutex> lock(g_timing_mu);
    float total = 0.f;
    for (int i = 0; i < g_timed; ++i) {
        ONDA_CUDA_TRY(cudaEventSynchronize(g_ev[i][1]));
        float ms = 0.f;
        ONDA_CUDA_TRY(cudaEventElapsedTime(&ms, g_ev[i][0], g_ev[i][1]));
        total += ms;
    }
    *total_ms_host = total;
    *launches_host = g_timed;
    g_timed = 0;
    return ONDA_OK;
}

size_t onda_table_floats(int C, int D) { return table_layout(C, D).total; }

size_t onda_sums_floats(int C, int D) { return sums_floats(C, D); }

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct Workspace {
    size_t off_cta, off_stat, off_dots, total;
    int max_cta, max_stat;
};

static Workspace fused_workspace(int B, int D, int HW, int C) {
    const int sms = cached_sm_count();
    const SimtPlan pl = plan_simt(B, D, HW, C, sms, true, true);
    Workspace w;
    w.max_cta = 2 * sms > pl.grid_x ? 2 * sms : pl.grid_x;
    w.max_stat = 4 * sms;
    w.off_cta = 0;
    w.off_stat = align256(w.off_cta + (size_t)w.max_cta * sums_floats(C, D) * sizeof(float));
    w.off_dots = align256(w.off_stat + (size_t)w.max_stat * kStatSlots * sizeof(float));
    int slices = pl.nslices;
    if (tc_supported(B, D, HW, C)) {
        const int ts = tc_slices(tc_tiles(B, HW), D, sms);
        slices = ts > slices ? ts : slices;
    }
    const size_t dots = slices > 1 ? (size_t)slices * (padded_classes(C) + 1) * (size_t)B * HW * sizeof(float) : 0;
    w.total = align256(w.off_dots + dots);
    return w;
}

int onda_impl_supported(int B, int D, int HW, int C, int impl) {
    if (B <= 0 || D <= 0 || HW <= 0 || C <= 0 || C > ONDA_MAX_CLASSES) return 0;
    if (impl == ONDA_IMPL_TCGEN05) return tc_supported(B, D, HW, C) ? 1 : 0;
    return (impl == ONDA_IMPL_SIMT || impl == ONDA_IMPL_AUTO) ? 1 : 0;
}

size_t onda_fused_workspace_bytes(int B, int D, int HW, int C, int impl) {
    (void)impl;
    if (B <= 0 || D <= 0 || HW <= 0 || C <= 0 || C > ONDA_MAX_CLASSES) return 0;
    return fused_workspace(B, D, HW, C).total;
}

int onda_build_distance_table(const float* prototypes, const float* squared_mean, const float* counter, int C, int D,
                              int metric, float* table, void* stream) {
    ONDA_REQUIRE(prototypes && table, "onda_build_distance_table: null pointer");
    ONDA_REQUIRE(C > 0 && C <= ONDA_MAX_CLASSES && D > 0, "onda_build_distance_table: unsupported shape C=%d D=%d", C, D);
    ONDA_REQUIRE(metric == ONDA_METRIC_EUCLIDEAN || metric == ONDA_METRIC_MAHALANOBIS,
                 "onda_build_distance_table: unexpected value for attribute distance_metric (%d)", metric);
    if (metric == ONDA_METRIC_MAHALANOBIS)
        ONDA_REQUIRE(squared_mean && counter, "onda_build_distance_table: mahalanobis needs squared_mean and counter");
    table_kernel<<<round_up(D, 32) / 32, kTableThreads, 0, (cudaStream_t)stream>>>(const_cast<float*>(prototypes), const_cast<float*>(squared_mean), counter, C, D, metric,
                                                                                   table, nullptr, 0.f, PeerTable{}, 0, 0, 0u, nullptr, nullptr, nullptr);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_table_global_std(const float* table, int C, int D, float* sigma_out, void* stream) {
    ONDA_REQUIRE(table && sigma_out && C > 0 && D > 0, "onda_table_global_std: bad argument");
    const TableLayout T = table_layout(C, D);
    copy_kernel<<<(D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table + T.off_sigma, sigma_out, D);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_prototype_std(const float* prototypes, const float* squared_mean, int C, int D, float* out, void* stream) {
    ONDA_REQUIRE(prototypes && squared_mean && out && C > 0 && D > 0, "onda_prototype_std: bad argument");
    const int n = C * D;
    class_std_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(prototypes, squared_mean, out, n);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_pseudolabel_fused(const float* feat, const float* prior, const float* logits, const float* table, int B, int D,
                           int HW, int C, float tau, float thresh, int64_t* labels, float* soft, float* dist,
                           float* sums, void* workspace, size_t workspace_bytes, int impl, void* stream_) {
    return onda_pseudolabel_fused_guarded(feat, prior, logits, table, B, D, HW, C, tau, thresh, labels, soft, dist, sums,
                                          workspace, workspace_bytes, impl, nullptr, nullptr, 0, stream_);
}

int onda_pseudolabel_fused_guarded(const float* feat, const float* prior, const float* logits, const float* table, int B,
                                   int D, int HW, int C, float tau, float thresh, int64_t* labels, float* soft,
                                   float* dist, float* sums, void* workspace, size_t workspace_bytes, int impl,
                                   const uint32_t* peer_done, const uint32_t* peer_epoch, int peer_world,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ONDA_REQUIRE((peer_done == nullptr) == (peer_epoch == nullptr) && peer_world >= 0 && peer_world <= kMaxPeers,
                 "onda_pseudolabel_fused_guarded: bad peer guard");
    ONDA_REQUIRE(feat && sums && workspace, "onda_pseudolabel_fused: null feat/sums/workspace");
    ONDA_REQUIRE(B > 0 && D > 0 && HW > 0, "onda_pseudolabel_fused: empty shape B=%d D=%d HW=%d", B, D, HW);
    ONDA_REQUIRE(C > 0 && C <= ONDA_MAX_CLASSES, "onda_pseudolabel_fused: %d classes unsupported (max %d)", C,
                 ONDA_MAX_CLASSES);
    const bool want_dist = labels || soft || dist;
    const bool want_sums = logits != nullptr;
    ONDA_REQUIRE(want_dist || want_sums, "onda_pseudolabel_fused: nothing to compute");
    ONDA_REQUIRE(!want_dist || table, "onda_pseudolabel_fused: distance outputs need a distance table");
    ONDA_REQUIRE(!(labels || soft) || prior, "onda_pseudolabel_fused: labels/soft need a prior");
    ONDA_REQUIRE(impl == ONDA_IMPL_AUTO || impl == ONDA_IMPL_SIMT || impl == ONDA_IMPL_TCGEN05,
                 "onda_pseudolabel_fused: unknown impl %d", impl);
    const long long n_tiles = ((long long)B * HW + kTilePixels - 1) / kTilePixels;
    const bool tc_ok = want_dist && tc_supported(B, D, HW, C);
    ONDA_REQUIRE(impl != ONDA_IMPL_TCGEN05 || tc_ok || !want_dist,
                 "onda_pseudolabel_fused: the tcgen05 kernel covers D % 64 == 0 and C <= 32 "
                 "(got D=%d C=%d)", D, C);
    // AUTO: tensor-core kernel when the shape allows and there are enough tiles to fill the machine
    const bool use_tc = tc_ok && (impl == ONDA_IMPL_TCGEN05 || (impl == ONDA_IMPL_AUTO && n_tiles >= 32));
    const Workspace ws = fused_workspace(B, D, HW, C);
    ONDA_REQUIRE(workspace_bytes >= ws.total, "onda_pseudolabel_fused: workspace too small (%zu < %zu)", workspace_bytes,
                 ws.total);
    const int sms = cached_sm_count();
    const SimtPlan pl = plan_simt(B, D, HW, C, sms, want_dist, want_sums);

    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.feat = feat; p.prior = prior; p.logits = logits; p.table = table;
    p.B = B; p.D = D; p.HW = HW; p.C = C; p.N = (long long)B * HW;
    p.tau = tau; p.inv_tau = 1.0f / tau; p.thresh = thresh;
    p.labels = (long long*)labels; p.soft = soft; p.dist = dist;
    char* base = (char*)workspace;
    p.cta_partials = (float*)(base + ws.off_cta);
    p.stat_partials = (float*)(base + ws.off_stat);
    p.dots_scratch = (float*)(base + ws.off_dots);
    p.nslices = pl.nslices; p.slice_channels = pl.DS; p.tiles = pl.tiles;
    p.debug = g_debug;

    int n_cta, n_stat;
    if (use_tc) {
        p.tiles_per_img = (HW + kTilePixels - 1) / kTilePixels;
        p.sched = reinterpret_cast<unsigned*>(const_cast<float*>(table) + table_layout(C, D).off_sched);
        p.tiles = tc_tiles(B, HW);
        const int slices = tc_slices(p.tiles, D, sms);
        p.nslices = slices; p.slice_channels = D / slices; p.cluster = 1;
        n_cta = tc_grid(p.tiles, sms, slices);
        n_stat = n_cta;
        int rc = launch_fused_tc(p, n_cta, want_sums, stream);
        if (rc != ONDA_OK) return rc;
        if (slices > 1) {          // the per-pixel tail over the parked partial dot products
            FusedParams f = p;
            f.tiles = pl.tiles;    // tiles of the flattened pixel axis
            rc = launch_split_finish(f, sms, &n_stat, stream);
            if (rc != ONDA_OK) return rc;
        }
    } else {
        int rc = launch_fused_simt(p, pl, want_dist, want_sums, stream);
        if (rc != ONDA_OK) return rc;
        n_cta = pl.grid_x;
        n_stat = want_dist ? (pl.nslices > 1 ? pl.finish_grid : pl.grid_x) : 0;
    }
    const int class_elems = 2 * C * D + C;
    const int blocks = (class_elems + kStatSlots + 31) / 32;
    ONDA_CUDA_TRY(launch_chained(reduce_partials_kernel, dim3(blocks), dim3(256), 0, stream, p.cta_partials, n_cta, sums_floats(C, D),
                                 class_elems, want_sums ? 1 : 0, p.stat_partials, n_stat, want_dist ? 1 : 0, sums, peer_done,
                                 peer_epoch, peer_world));
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_class_sums_labelled(const float* feat, const int64_t* class_ids, int B, int D, int HW, int C, float* sums,
                             void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ONDA_REQUIRE(feat && class_ids && sums && workspace, "onda_class_sums_labelled: null pointer");
    ONDA_REQUIRE(B > 0 && D > 0 && HW > 0 && C > 0 && C <= ONDA_MAX_CLASSES, "onda_class_sums_labelled: bad shape");
    const Workspace ws = fused_workspace(B, D, HW, C);
    ONDA_REQUIRE(workspace_bytes >= ws.total, "onda_class_sums_labelled: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    const int sms = cached_sm_count();
    const SimtPlan pl = plan_simt(B, D, HW, C, sms, false, true);
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.feat = feat; p.class_ids = (const long long*)class_ids;
    p.B = B; p.D = D; p.HW = HW; p.C = C; p.N = (long long)B * HW;
    p.tau = 1.f; p.inv_tau = 1.f;
    char* base = (char*)workspace;
    p.cta_partials = (float*)(base + ws.off_cta);
    p.stat_partials = (float*)(base + ws.off_stat);
    p.dots_scratch = (float*)(base + ws.off_dots);
    p.nslices = pl.nslices; p.slice_channels = pl.DS; p.tiles = pl.tiles;
    int rc = launch_fused_simt(p, pl, false, true, stream);
    if (rc != ONDA_OK) return rc;
    const int class_elems = 2 * C * D + C;
    reduce_partials_kernel<<<(class_elems + kStatSlots + 31) / 32, 256, 0, stream>>>(p.cta_partials, pl.grid_x, sums_floats(C, D),
                                                                                   class_elems, 1, p.stat_partials, 0, 0, sums,
                                                                                   nullptr, nullptr, 0);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_ema_update(float* prototypes, float* squared_mean, const float* sums, int C, int D, float ma_lambda,
                    void* stream) {
    ONDA_REQUIRE(prototypes && squared_mean && sums && C > 0 && D > 0, "onda_ema_update: bad argument");
    const int n = C * D;
    ema_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(prototypes, squared_mean, sums, C, D, ma_lambda);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_ema_update_and_table(float* prototypes, float* squared_mean, const float* counter, const float* sums, int C,
                              int D, float ma_lambda, int metric, float* table, void* stream) {
    ONDA_REQUIRE(prototypes && squared_mean && sums && table, "onda_ema_update_and_table: null pointer");
    ONDA_REQUIRE(C > 0 && C <= ONDA_MAX_CLASSES && D > 0, "onda_ema_update_and_table: unsupported shape C=%d D=%d", C, D);
    ONDA_REQUIRE(metric == ONDA_METRIC_EUCLIDEAN || metric == ONDA_METRIC_MAHALANOBIS,
                 "onda_ema_update_and_table: unexpected value for attribute distance_metric (%d)", metric);
    if (metric == ONDA_METRIC_MAHALANOBIS) ONDA_REQUIRE(counter, "onda_ema_update_and_table: mahalanobis needs counter");
    ONDA_CUDA_TRY(launch_chained(table_kernel, dim3(round_up(D, 32) / 32), dim3(kTableThreads), 0, (cudaStream_t)stream, prototypes,
                                 squared_mean, counter, C, D, metric, table, sums, ma_lambda, PeerTable{}, 0, 0, 0u, (float*)nullptr,
                                 (uint32_t*)nullptr, g_debug));
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_ema_update_and_table_allreduce(float* prototypes, float* squared_mean, const float* counter, float* sums_out,
                                        int C, int D, float ma_lambda, int metric, float* table, int rank, int world,
                                        void* const* peer_bufs_host, void* const* peer_flags_host,
                                        void* const* peer_done_host, uint32_t epoch, uint32_t* epoch_counter,
                                        void* stream) {
    ONDA_REQUIRE(prototypes && squared_mean && sums_out && table && peer_bufs_host && peer_flags_host,
                 "onda_ema_update_and_table_allreduce: null pointer");
    ONDA_REQUIRE(C > 0 && C <= ONDA_MAX_CLASSES && D > 0, "onda_ema_update_and_table_allreduce: unsupported shape C=%d D=%d", C, D);
    ONDA_REQUIRE(metric == ONDA_METRIC_EUCLIDEAN || metric == ONDA_METRIC_MAHALANOBIS,
                 "onda_ema_update_and_table_allreduce: unexpected value for attribute distance_metric (%d)", metric);
    if (metric == ONDA_METRIC_MAHALANOBIS) ONDA_REQUIRE(counter, "onda_ema_update_and_table_allreduce: mahalanobis needs counter");
    ONDA_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "onda_ema_update_and_table_allreduce: bad rank %d / world %d", rank, world);
    ONDA_REQUIRE(epoch != 0 || epoch_counter, "onda_ema_update_and_table_allreduce: epoch 0 is the flags' initial value");
    for (int r = 0; r < world; ++r)
        ONDA_REQUIRE(peer_bufs_host[r] && peer_flags_host[r], "onda_ema_update_and_table_allreduce: null peer pointer for rank %d", r);
    const PeerTable peers = make_peer_table(rank, world, peer_bufs_host, peer_flags_host, peer_done_host);
    ONDA_CUDA_TRY(launch_chained(table_kernel, dim3(round_up(D, 32) / 32), dim3(kTableThreads), 0, (cudaStream_t)stream, prototypes,
                                 squared_mean, counter, C, D, metric, table, (const float*)nullptr, ma_lambda, peers, rank, world, epoch,
                                 sums_out, epoch_counter, g_debug));
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_append_update(float* prototypes, float* squared_mean, float* counter, const float* sums, int C, int D,
                       void* stream) {
    ONDA_REQUIRE(prototypes && squared_mean && counter && sums && C > 0 && C <= ONDA_MAX_CLASSES && D > 0,
                 "onda_append_update: bad argument");
    const int n = C * D;
    append_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(prototypes, squared_mean, counter, sums, C, D);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    counter_add_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter, sums, C, D);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

size_t onda_prior_workspace_bytes(int B, int C, int HW) {
    (void)B; (void)C; (void)HW;
    return 256 + (size_t)8 * cached_sm_count() * kStatSlots * sizeof(float);
}

int onda_prior_mix_stats(const float* logits0, const float* logits1, const float* logits2, float coef0, float coef1,
                         float coef2, float scale01, int B, int C, int HW, float* prior_out, float* stats_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    ONDA_REQUIRE(stats_out && workspace, "onda_prior_mix_stats: null stats/workspace");
    ONDA_REQUIRE(logits0 || logits1 || logits2, "onda_prior_mix_stats: no input");
    ONDA_REQUIRE(B > 0 && HW > 0 && C > 0 && C <= ONDA_MAX_CLASSES && (unsigned long long)B * C * HW < (1ull << 32),
                 "onda_prior_mix_stats: bad shape");
    ONDA_REQUIRE(workspace_bytes >= onda_prior_workspace_bytes(B, C, HW), "onda_prior_mix_stats: workspace too small");
    const long long N = (long long)B * HW;
    const int sms = cached_sm_count();
    long long want = (N + kPriorThreads - 1) / kPriorThreads;
    const int grid = (int)(want < 8LL * sms ? want : 8LL * sms);
    unsigned* ticket = (unsigned*)workspace;  // zero on first use; the kernel re-arms it
    float* partials = (float*)((char*)workspace + 256);
    if (padded_classes(C) == 20)
        prior_mix_kernel<20><<<grid, kPriorThreads, 0, (cudaStream_t)stream>>>(logits0, logits1, logits2, coef0, coef1, coef2, scale01, B,
                                                                               C, HW, prior_out, partials, ticket, stats_out);
    else
        prior_mix_kernel<32><<<grid, kPriorThreads, 0, (cudaStream_t)stream>>>(logits0, logits1, logits2, coef0, coef1, coef2, scale01, B,
                                                                               C, HW, prior_out, partials, ticket, stats_out);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

size_t onda_step_log_workspace_bytes(void) { return 256 + (size_t)4 * cached_sm_count() * kLogSlots * sizeof(double); }

int onda_step_log_stats(const int64_t* labels, const float* student_logits, const float* prototypes, int B, int C, int HW,
                        int D, float* out4, void* workspace, size_t workspace_bytes, void* stream) {
    ONDA_REQUIRE(labels && student_logits && prototypes && out4 && workspace, "onda_step_log_stats: null pointer");
    ONDA_REQUIRE(B > 0 && HW > 0 && D > 0 && C > 0 && C <= ONDA_MAX_CLASSES, "onda_step_log_stats: bad shape");
    ONDA_REQUIRE(workspace_bytes >= onda_step_log_workspace_bytes(), "onda_step_log_stats: workspace too small");
    const long long N = (long long)B * HW;
    const int sms = cached_sm_count();
    long long want = (N + kPriorThreads - 1) / kPriorThreads;
    const int grid = (int)(want < 4LL * sms ? want : 4LL * sms);
    unsigned* ticket = (unsigned*)workspace;  // zero on first use; the kernel re-arms it
    double* partials = (double*)((char*)workspace + 256);
    if (padded_classes(C) == 20)
        step_log_kernel<20><<<grid, kPriorThreads, 0, (cudaStream_t)stream>>>((const long long*)labels, student_logits, prototypes, B, C,
                                                                              HW, C * D, partials, ticket, out4);
    else
        step_log_kernel<32><<<grid, kPriorThreads, 0, (cudaStream_t)stream>>>((const long long*)labels, student_logits, prototypes, B, C,
                                                                              HW, C * D, partials, ticket, out4);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_weight_ema_update(const onda_ema_chunk* chunks_device, int n_chunks, float keep, float take, void* stream) {
    ONDA_REQUIRE(n_chunks >= 0 && (n_chunks == 0 || chunks_device), "onda_weight_ema_update: bad chunk table");
    if (n_chunks == 0) return ONDA_OK;
    const int sms = cached_sm_count();
    const int grid = n_chunks < 8 * sms ? n_chunks : 8 * sms;
    weight_ema_kernel<<<grid, kEmaThreads, 0, (cudaStream_t)stream>>>(chunks_device, n_chunks, keep, take);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_confusion_update(const float* logits, int B, int C, int h, int w, const int64_t* labels, int H, int W,
                          unsigned long long* hist, unsigned char* pred_out, void* stream) {
    ONDA_REQUIRE(logits && labels && hist, "onda_confusion_update: null pointer");
    ONDA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && C <= ONDA_MAX_CLASSES,
                 "onda_confusion_update: bad shape B=%d C=%d %dx%d -> %dx%d", B, C, h, w, H, W);
    const int sms = cached_sm_count();
    // rows per strip: long strips reuse the interpolated source rows longest; short ones when the image alone would
    // not fill the machine (at least two CTAs per SM wanted)
    int strip_rows = 32;
    long long want = 0;
    for (;; strip_rows >>= 1) {
        want = (long long)B * ((H + strip_rows - 1) / strip_rows) * ((W + kConfThreads - 1) / kConfThreads);
        if (want >= 2LL * sms || strip_rows <= 4) break;
    }
    const int grid = (int)(want < 8LL * sms ? want : 8LL * sms);
    if (padded_classes(C) == 20)
        confusion_kernel<20><<<grid, kConfThreads, 0, (cudaStream_t)stream>>>(logits, B, C, h, w, (const long long*)labels, H, W, strip_rows, hist, pred_out);
    else
        confusion_kernel<32><<<grid, kConfThreads, 0, (cudaStream_t)stream>>>(logits, B, C, h, w, (const long long*)labels, H, W, strip_rows, hist, pred_out);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

int onda_allreduce_oneshot(float* out, size_t n, int rank, int world, void* const* peer_bufs_host,
                           void* const* peer_flags_host, uint32_t epoch, void* stream) {
    ONDA_REQUIRE(out && peer_bufs_host && peer_flags_host, "onda_allreduce_oneshot: null pointer");
    ONDA_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "onda_allreduce_oneshot: bad rank %d / world %d", rank, world);
    ONDA_REQUIRE(epoch != 0, "onda_allreduce_oneshot: epoch 0 is the flags' initial value");
    for (int r = 0; r < world; ++r)
        ONDA_REQUIRE(peer_bufs_host[r] && peer_flags_host[r], "onda_allreduce_oneshot: null peer pointer for rank %d", r);
    return launch_allreduce_oneshot(out, n, rank, world, peer_bufs_host, peer_flags_host, epoch, (cudaStream_t)stream);
}

}  // extern "C"
