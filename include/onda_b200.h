/*
 * onda_b200.h -- C ABI of the B200-native prototype pseudo-labelling path.
 *
 * theo2021/OnDA has no FFI of its own: the boundary of this path in the reference
 * is the Python duck type `prototype_handler`
 * (framework/domain_adaptation/methods/prototype_handler.py:8-166) and the
 * `prototype_predictions()` methods that drive it
 * (framework/domain_adaptation/methods/prototypes_hybrid_switch.py:45-101 and
 * siblings).  The entry points below are what a ctypes binding for that class
 * calls; each one names the reference lines it replaces.  INTEGRATION.md shows
 * the reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all arithmetic is fp32, labels are int64, tensors are dense and laid out
 *     exactly as the reference holds them (feat/prior/logits NCHW, results
 *     pixel-major (N, C) with n = (b*H + y)*W + x, prototype_handler.py:105-109);
 *   - `stream` is a cudaStream_t passed as void*; nothing here allocates device
 *     memory or synchronises with the host unless stated;
 *   - return value: 0 on success, a negative ONDA_E* code otherwise;
 *     onda_last_error() returns a thread-local description of the last failure.
 */
#ifndef ONDA_B200_H
#define ONDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONDA_OK 0
#define ONDA_EINVAL (-1)    /* bad argument (shape, null pointer, unsupported class count) */
#define ONDA_ECUDA (-2)     /* a CUDA runtime call or launch failed */
#define ONDA_EUNSUPPORTED (-3)

#define ONDA_MAX_CLASSES 32 /* the reference uses 19 */
#define ONDA_IGNORE_LABEL 255 /* prototype_handler.py:165 */

#define ONDA_METRIC_EUCLIDEAN 0   /* prototype_handler.distance, :127-138 */
#define ONDA_METRIC_MAHALANOBIS 1 /* prototype_handler.mahalanobis_distance, :111-125 */

/* kernel selection for onda_pseudolabel_fused (ONDA_IMPL_AUTO picks tcgen05 when the shape allows) */
#define ONDA_IMPL_AUTO 0
#define ONDA_IMPL_SIMT 1    /* CUDA-core kernel, any shape */
#define ONDA_IMPL_TCGEN05 2 /* 3xTF32 tcgen05.mma kernel: D % 64 == 0, C <= 32 (see onda_impl_supported) */

/* Number of statistic slots at the tail of a `sums` buffer (see onda_sums_floats). */
#define ONDA_NUM_STATS 8
#define ONDA_STAT_PROTO_CONF 0  /* sum_n max_k softmax(-d/tau)      prototype_handler.py:150 */
#define ONDA_STAT_PRIOR_CONF 1  /* sum_n max_k prior[n,k]           prototypes_hybrid_switch.py:88 */
#define ONDA_STAT_PL_CONF 2     /* sum_n max_k r[n,k]               prototypes_hybrid_switch.py:94-96 */
#define ONDA_STAT_PL_PIXELS 3   /* #{n : label != 255}              prototypes.py:341-345 */
#define ONDA_STAT_PIXELS 4      /* N (so that an all-reduced buffer carries its own denominator) */
#define ONDA_STAT_ENTROPY 5     /* sum_n sum_k -r*log2(r+1e-30)/log2(C)   framework/utils/func.py:71-74 */

/* ---- library / device info ------------------------------------------------ */
int onda_abi_version(void);
const char* onda_last_error(void);
/* number of SMs of the current device (148 on B200); negative on error */
int onda_sm_count(void);
/* kernels launched by this library since it was loaded (process-wide, monotonically increasing) */
unsigned long long onda_launch_count(void);

/*
 * Per-kernel timing of the dominant kernel (the fused pass) for the roofline report: when enabled,
 * every onda_pseudolabel_fused call brackets its main kernel with CUDA events on the launch stream.
 * (`enable` = n > 0 brackets every n-th call, to keep the probe effect on the timed region small; 0 = off).
 * onda_kernel_timing_read synchronises those events and returns the summed duration and the count
 * since the last enable/reset.  Off by default; at most 4096 launches are recorded per window.
 */
int onda_kernel_timing_enable(int enable);
int onda_kernel_timing_read(float* total_ms_host, int* launches_host);

/* Launch chain.  onda_pseudolabel_fused[_guarded] (its tcgen05 kernel and the combine of the per-CTA partials) and
 * onda_ema_update_and_table[_allreduce] launch their kernels with the programmatic-stream-serialization attribute: on
 * one stream, a kernel of the chain may be scheduled while its predecessor is still running and waits
 * (griddepcontrol.wait) before its first global-memory access, so back-to-back steps do not pay the launch latencies.
 * Results are bit-identical to plain launches; ONDA_PDL=0 in the environment (read once) selects plain launches.
 * The kernel-timing diagnostics above are thread-safe; the tile schedule and the debug buffer are process-wide settings. */

/* Tile schedule of the tcgen05 kernel (process-wide; returns the previous value).  1 (default, or ONDA_TC_DYNAMIC_TILES
 * unset): tiles beyond the first two per SM are drawn from a device counter -- fastest, but which SM accumulates which
 * tile, and with it the last bits of the class sums, varies from run to run.  0: fixed round-robin schedule, class sums
 * bit-reproducible (what a test that compares two runs bit for bit wants).  Other values only query. */
int onda_set_tile_schedule(int dynamic);
/* Diagnostics: when a device buffer of gridDim*32*8 int64 is set, the tcgen05 kernel records per-warp cycle
 * counters (time spent in each pipeline wait, total) into it.  NULL (default) disables it. */
int onda_debug_set_buffer(void* device_buffer);
/* Measurement tool: runs a kernel that performs only the fused pass's loads of `feat` [B, D, HW] (same tile walk,
 * `workers` warps per SM in groups of four, `smem_bytes` of dynamic shared memory to set the L1 size) and writes
 * one float per warp to out[sm_count * 32].  Its bandwidth is what this access pattern can reach at best. */
int onda_debug_load_probe(const float* feat, int B, int D, int HW, int workers, int smem_bytes, float* out, void* stream);

/* ---- buffer sizing (host, no CUDA calls) ----------------------------------- */
/* floats in a distance table built by onda_build_distance_table */
size_t onda_table_floats(int C, int D);
/* floats in a `sums` buffer: [C*D sum | C*D sum of squares | C counts | ONDA_NUM_STATS stats] */
size_t onda_sums_floats(int C, int D);
/* bytes of scratch onda_pseudolabel_fused needs for this shape (per-CTA partials, split-D dots) */
size_t onda_fused_workspace_bytes(int B, int D, int HW, int C, int impl);

/* 1 if `impl` (ONDA_IMPL_SIMT / ONDA_IMPL_TCGEN05) has a kernel for this shape, else 0 (host only) */
int onda_impl_supported(int B, int D, int HW, int C, int impl);

/* ---- prototype statistics ------------------------------------------------- */
/*
 * Builds the per-step distance table from the handler state: pooled std
 * (global_var, prototype_handler.py:53-60), its inverse variance, the centring
 * point, the scaled prototypes Q = w*(P-mu) and the per-class bias
 * sum_j w_j (P_kj-mu_j)^2, so that d^2[n,k] = sum_j w_j (x_nj-mu_j)^2 - 2 (x_n-mu).Q_k + bias_k
 * equals the reference's sum_j ((x_nj-P_kj)/sigma_j)^2 (prototype_handler.py:117-120, :132-135).
 * `squared_mean` and `counter` may be NULL for ONDA_METRIC_EUCLIDEAN.  The table buffer (onda_table_floats floats)
 * must be zero-initialised before its first use (it holds a ticket counter that the kernel re-arms itself).
 */
int onda_build_distance_table(const float* prototypes, const float* squared_mean, const float* counter,
                              int C, int D, int metric, float* table, void* stream);
/* sigma[D]: pooled std of global_var(), prototype_handler.py:53-60 (reads it back out of a table) */
int onda_table_global_std(const float* table, int C, int D, float* sigma_out, void* stream);
/* out[C*D] = sqrt(squared_mean - prototypes^2): prototype_var(), prototype_handler.py:49-51 */
int onda_prototype_std(const float* prototypes, const float* squared_mean, int C, int D, float* out, void* stream);

/* ---- the fused pass --------------------------------------------------------- */
/*
 * One pass over feat (B, D, HW) that produces any subset of:
 *   labels[N]  int64  argmax_k r, 255 where max_k r < thresh   (pseudo_labels hard,  :163-166)
 *   soft[N*C]  f32    r = q*prior / sum_k q*prior, q = softmax_k(-(d-min d)/tau)  (:147, :159-160)
 *   dist[N*C]  f32    d - min_k d                              (distance / mahalanobis_distance, :111-138)
 *   sums       f32    class sum / sum of squares / count keyed by argmax_k logits[n,k]
 *                     (get_proto_array on feat and feat**2, :76-90) plus the ONDA_STAT_* sums
 * Outputs whose pointer is NULL are skipped; `prior` may be NULL only if labels and soft are;
 * `logits` NULL skips the class sums.  `sums` must hold onda_sums_floats(C, D) floats and is
 * fully overwritten (deterministically: fixed-order combine of per-CTA partials).
 * `workspace` must hold onda_fused_workspace_bytes(...) bytes.
 */
int onda_pseudolabel_fused(const float* feat, const float* prior, const float* logits, const float* table,
                           int B, int D, int HW, int C, float tau, float thresh,
                           int64_t* labels, float* soft, float* dist, float* sums,
                           void* workspace, size_t workspace_bytes, int impl, void* stream);
/*
 * Same, for a `sums` buffer that peers read over NVLink (onda_ema_update_and_table_allreduce): `sums` is written only
 * once every flag peer_done[r] (r < peer_world, this rank's own "done" words) has reached *peer_epoch - 1, i.e. once
 * every rank has finished reading what the buffer held for the previous exchange.  This makes a replayed CUDA graph
 * of the step safe with a single peer-visible buffer.  peer_done == NULL: no guard.
 */
int onda_pseudolabel_fused_guarded(const float* feat, const float* prior, const float* logits, const float* table,
                                   int B, int D, int HW, int C, float tau, float thresh,
                                   int64_t* labels, float* soft, float* dist, float* sums,
                                   void* workspace, size_t workspace_bytes, int impl,
                                   const uint32_t* peer_done, const uint32_t* peer_epoch, int peer_world, void* stream);

/* ---- prototype updates ------------------------------------------------------ */
/*
 * Class sums keyed by labels given directly, straight from the NCHW feature map: sums = { sum_{n: id_n = k} x_n,
 * sum of squares, count } for class_ids[N] (int64; values outside [0, C), e.g. the ignore label 255, are skipped;
 * the statistics tail is zero).  Replaces the mask-gather + transposition + one-hot matmul of
 * online_proDA.calculate_prototypes with STARTING_PROTO == "source"
 * (framework/domain_adaptation/methods/prototypes.py:142-154) in front of append().
 */
int onda_class_sums_labelled(const float* feat, const int64_t* class_ids, int B, int D, int HW, int C, float* sums,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ma(): P_k <- P_k*rho_k + (1-rho_k)*sum_k/max(cnt_k,1), rho_k = lambda if cnt_k>0 else 1; same for the
 * squared mean; counter untouched.  prototype_handler.py:88-99. */
int onda_ema_update(float* prototypes, float* squared_mean, const float* sums, int C, int D, float ma_lambda,
                    void* stream);
/* ma() and the rebuild of the distance table for the next step in ONE launch (same results as onda_ema_update
 * followed by onda_build_distance_table; `table` must have been zero-initialised once, like for the plain build). */
int onda_ema_update_and_table(float* prototypes, float* squared_mean, const float* counter, const float* sums, int C,
                              int D, float ma_lambda, int metric, float* table, void* stream);
/* append(): counter += cnt; P += (sum - P*cnt)/max(counter,1); likewise S.  prototype_handler.py:62-74. */
int onda_append_update(float* prototypes, float* squared_mean, float* counter, const float* sums, int C, int D,
                       void* stream);

/* ---- switch statistics / prior mix ------------------------------------------ */
/*
 * prior[b,k,p] = (coef0 * softmax_k(logits_0) + coef1 * softmax_k(logits_1)) * scale01 + coef2 * softmax_k(logits_2)
 * over the non-NULL inputs, every product and sum rounded to fp32 in this order (scale01 is the h-switch's
 * `prior *= percentage_static`, prototypes_hswitch.py:56; pass 1 otherwise), written to
 * `prior_out` (may be NULL: statistics only), and stats_out[0..2] = sum_n max_k softmax(logits_i)[n,k],
 * stats_out[3] = sum_n max_k prior[n,k], stats_out[4] = N.  Replaces the softmax / max / mean chains of
 * prototypes_hybrid_switch.py:52-88, prototypes_hswitch.py:30-68, prototypes_vswitch.py:40-70 and
 * prototypes.py:213-250.  stats_out holds 8 floats and is overwritten.
 */
int onda_prior_mix_stats(const float* logits0, const float* logits1, const float* logits2,
                         float coef0, float coef1, float coef2, float scale01, int B, int C, int HW,
                         float* prior_out, float* stats_out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Per-step log reductions of the method classes (framework/domain_adaptation/methods/prototypes.py:341-352), one
 * launch: out4 = { number of pixels whose pseudo-label equals the first argmax over classes of student_logits
 * [B, C, HW] (":346-347", labels [B*HW] int64, 255 never agrees), number of labels that are neither negative nor 255
 * (":342"), sum of squared prototype entries (":351", divide by C*D for the mean), B*HW }.  Counts are exact.
 * workspace: onda_step_log_workspace_bytes() bytes, zeroed once. */
size_t onda_step_log_workspace_bytes(void);
int onda_step_log_stats(const int64_t* labels, const float* student_logits, const float* prototypes, int B, int C, int HW,
                        int D, float* out4, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Loss-side consumers of the pseudo-labels, forward + gradient in one pass over the student logits [B, C, HW]
 * (online_proDA.pseudolabel_loss, prototypes.py:313-336, hard labels):
 *   out6[0] = ce    = cross_entropy_2d(logits, labels)          framework/utils/loss.py:16-45
 *   out6[1] = rce   = rce(logits, labels)                        framework/utils/loss.py:88-112
 *   out6[2] = reg   = regular_loss(regularizer, logits)          prototypes.py:29-39 (ONDA_REG_*)
 *   out6[3] = alpha * ce + beta * rce + reg_weight * reg         (RCE_ALPHA, RCE_BETA, REGULARIZER_WEIGHT)
 *   out6[4] = mean(labels == argmax_k logits)                    "output & prototype agreement", prototypes.py:346-347
 *   out6[5] = number of labels that are neither negative nor 255
 *   grad[B, C, HW] (may be NULL) = d out6[3] / d logits.
 * `labels` [B*HW] int64 in the pixel order of the fused pass, 255 = ignore.  `n_valid` (device, may be NULL): that
 * count as a float, e.g. &sums[2*C*D + C + ONDA_STAT_PL_PIXELS] of the fused pass; NULL adds a counting launch.
 * `workspace`: onda_target_loss_workspace_bytes() bytes, zero-filled before its first use.
 */
size_t onda_target_loss_workspace_bytes(void);
int onda_target_loss_fused(const float* student_logits, const int64_t* labels, int B, int C, int HW, const float* n_valid,
                           float alpha, float beta, float reg_weight, int regularizer, float* grad, float* out6,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- model-weight EMA ("next" row f2 of the scope table) --------------------------------------------------
 * update_ema (framework/domain_adaptation/methods/prototypes.py:407-416) walks every parameter in a Python loop with
 * two clones and three kernels each, then copies every buffer.  Here: ONE launch over a table of chunks that the
 * host builds once per model pair.  mode 0: dst[i] = dst[i]*keep + src[i]*take over `count` floats (both products
 * rounded, then the sum: bit-identical to `param_k.clone()*a + param_q.clone()*(1-a)` with keep = (float)a,
 * take = (float)(1.0 - a)); mode 1: copy `count` bytes (buffers of any dtype, ":414-416").  A chunk is at most
 * ONDA_EMA_CHUNK_BYTES long. */
#define ONDA_REG_NONE 0
#define ONDA_REG_MRKLD 1 /* regular_loss, framework/domain_adaptation/methods/prototypes.py:36-39 */
#define ONDA_REG_MRENT 2 /* :32-35 */

#define ONDA_EMA_CHUNK_BYTES 32768
typedef struct {
    const void* src; /* the trained model's tensor (chunk start) */
    void* dst;       /* the EMA model's tensor (chunk start), updated in place */
    uint32_t count;  /* floats (mode 0) or bytes (mode 1) */
    uint32_t mode;
} onda_ema_chunk;
int onda_weight_ema_update(const onda_ema_chunk* chunks_device, int n_chunks, float keep, float take, void* stream);

/* ---- evaluation ("next" row f3) ---------------------------------------------------------------------------------
 * da_model.evaluate (framework/domain_adaptation/methods/adaptation_model.py:143-160): interp(pred) (bilinear,
 * align_corners=True, ":94-98") -> softmax -> per-image argmax -> .cpu().numpy() -> fast_hist (framework/utils/func.py:
 * 77-79).  One launch instead: logits [B, C, h, w] at network resolution, labels [B, H, W] int64 at full resolution
 * (entries outside [0, C) are ignored like fast_hist's mask); hist [C*C] uint64 (row = label, column = prediction) is
 * ACCUMULATED (zero it before the first batch); pred_out [B*H*W] uint8 (nullable) receives the per-pixel prediction.
 * The prediction is the first argmax of the interpolated logits (softmax is monotone per pixel). */
int onda_confusion_update(const float* logits, int B, int C, int h, int w, const int64_t* labels, int H, int W,
                          unsigned long long* hist, unsigned char* pred_out, void* stream);

size_t onda_prior_workspace_bytes(int B, int C, int HW);

/* ---- multi-GPU ---------------------------------------------------------------- */
/*
 * One-shot sum all-reduce of `n` floats over `world` <= 8 ranks of one NVLink/NVSwitch node.
 * peer_bufs_host[r] / peer_flags_host[r] (HOST arrays of `world` DEVICE pointers) are rank r's input buffer
 * (n floats) and flag array (`world` uint32, zero-initialised) as mapped into this process (symmetric / peer
 * memory); this rank's own input is peer_bufs_host[rank].  The kernel stores `epoch` (non-zero, increasing,
 * the same on every rank for a given call) into flag[rank] of every peer, waits for all `world` flags of its
 * own array, then every rank reads all inputs and adds them in rank order 0..world-1 into `out` (local memory,
 * n floats): all ranks obtain bit-identical sums.  Inputs must be double-buffered by the caller (two slots
 * with separate flag arrays, alternating per call).  Replaces nothing in the reference (single GPU); it is the
 * exchange step of the batch-sharded path, DESIGN.md section (e).
 */
int onda_allreduce_oneshot(float* out, size_t n, int rank, int world, void* const* peer_bufs_host,
                           void* const* peer_flags_host, uint32_t epoch, void* stream);

/* ma() across ranks in one launch: onda_allreduce_oneshot + onda_ema_update_and_table fused.  Every rank's `sums`
 * (as written by onda_pseudolabel_fused) sits in its peer-mapped slot peer_bufs[rank]; the kernel does the flag
 * handshake of the one-shot all-reduce, reads the peers' slots over NVLink, adds them in rank order while it blends
 * (prototype_handler.py:88-99) and rebuilds the distance table, and writes the reduced buffer to sums_out
 * (onda_sums_floats(C, D) floats: the statistics tail is global afterwards).  Same slot / flag / epoch rules as
 * onda_allreduce_oneshot.  If epoch_counter (device, one uint32 per slot, initialised to the same non-zero value
 * on every rank) is given, the epoch is read from it and incremented by the kernel, and `epoch` is ignored: the
 * call can then be captured in a CUDA graph and replayed.  peer_done_host (or NULL): per rank r, the device pointer
 * of r's "done" words; once this rank has read every slot it stores the epoch into done[r][rank] on every peer, which
 * is what onda_pseudolabel_fused_guarded waits for before it overwrites a slot. */
int onda_ema_update_and_table_allreduce(float* prototypes, float* squared_mean, const float* counter, float* sums_out,
                                        int C, int D, float ma_lambda, int metric, float* table, int rank, int world,
                                        void* const* peer_bufs_host, void* const* peer_flags_host,
                                        void* const* peer_done_host, uint32_t epoch, uint32_t* epoch_counter,
                                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ONDA_B200_H */
