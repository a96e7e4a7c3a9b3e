// Loss-side consumers of the pseudo-labels: cross entropy + reverse cross entropy on hard labels and
// the MRKLD / MRENT regulariser of online_proDA.pseudolabel_loss (framework/domain_adaptation/methods/
// prototypes.py:313-336), forward AND gradient with respect to the student logits in ONE pass over (B, C, h, w):
//   ce   = mean over valid pixels of -log softmax(z)[label]                       (framework/utils/loss.py:16-45, hard)
//   rce  = sum over valid pixels of -sum_k p_k log(clamp(onehot_k, 1e-4, 1)) / (n_valid + 1e-6)      (loss.py:88-112)
//        = -log(1e-4) * sum over valid pixels of (1 - p_label) / (n_valid + 1e-6)
//   reg  = MRKLD: -sum log softmax(z) / (N * C);   MRENT: sum p log p / N                            (prototypes.py:29-39)
//   total = alpha * ce + beta * rce + reg_weight * reg
// The reference does this with ~10 tensor-wide passes plus a boolean-mask gather of the logits; here a thread owns a
// pixel (coalesced reads along the class planes), keeps its C probabilities in registers, and writes the gradient of
// `total` in the same pass.  n_valid comes from the fused pseudo-label pass (ONDA_STAT_PL_PIXELS, on the device) or
// from a counting pre-pass.  Block partial sums are folded in CTA order by the last CTA (deterministic).
#include "epilogue.cuh"

namespace onda {

constexpr int kLossThreads = 256;
constexpr int kLossSlots = 4;      // ce sum | rce sum | reg sum | agreeing pixels

__global__ void __launch_bounds__(kLossThreads) count_valid_kernel(const long long* __restrict__ labels, long long N, int C,
                                                                   unsigned long long* __restrict__ count) {
    unsigned local = 0;
    for (long long n = (long long)blockIdx.x * kLossThreads + threadIdx.x; n < N; n += (long long)gridDim.x * kLossThreads) {
        const long long l = labels[n];
        local += (l >= 0 && l != ONDA_IGNORE_LABEL) ? 1u : 0u;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);      // integer: order-independent
    (void)C;
}

template <int CP>
__global__ void __launch_bounds__(kLossThreads, CP == 20 ? 3 : 2) target_loss_kernel(const float* __restrict__ logits,
                                                                   const long long* __restrict__ labels, int B, int C, int HW,
                                                                   const float* __restrict__ n_valid_f,
                                                                   const unsigned long long* __restrict__ n_valid_u,
                                                                   float alpha, float beta, float reg_weight, int reg_kind,
                                                                   float* __restrict__ grad, double* __restrict__ partials,
                                                                   unsigned* __restrict__ ticket, float* __restrict__ out) {
    __shared__ double red[kLossThreads / 32][kLossSlots];
    __shared__ bool last;
    const unsigned N = (unsigned)B * (unsigned)HW, hw = (unsigned)HW;        // 32-bit indices (the host checks B*C*HW < 2^32)
    const float nv = n_valid_f != nullptr ? *n_valid_f : (float)*n_valid_u;
    const float log_floor = 9.210340371976182f;            // -log(1e-4): the clamp of the one-hot target (loss.py:104-106)
    const float g_ce = alpha / nv;                          // F.cross_entropy(..., size_average=True) over the selected pixels
    const float g_rce = beta * log_floor / (nv + 1e-6f);
    const float g_reg = reg_kind == ONDA_REG_MRKLD ? reg_weight / ((float)N * (float)C) : reg_weight / (float)N;
    double ce = 0.0, rce = 0.0, reg = 0.0, agree = 0.0;
    for (unsigned n = blockIdx.x * kLossThreads + threadIdx.x; n < N; n += gridDim.x * kLossThreads) {
        const unsigned b = n / hw, q = n - b * hw;
        const unsigned off = (b * (unsigned)C) * hw + q;
        float z[CP];
        float zmax = -__int_as_float(0x7f800000);
        int arg = 0;
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < C) {
                z[k] = __ldg(logits + off + (unsigned)k * hw);
                if (torch_greater(z[k], zmax)) { zmax = z[k]; arg = k; }
            }
        float esum = 0.f;
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < C) esum += __expf(z[k] - zmax);
        const float lse = zmax + __logf(esum);
        const long long lab = labels[n];
        const bool valid = lab >= 0 && lab != ONDA_IGNORE_LABEL && lab < C;
        agree += (lab == (long long)arg) ? 1.0 : 0.0;       // "output & prototype agreement", prototypes.py:346-347
        float p_l = 0.f, logp_sum = 0.f, plogp = 0.f;
        float p[CP];
#pragma unroll
        for (int k = 0; k < CP; ++k)
            if (k < C) {
                const float lp = z[k] - lse;
                p[k] = __expf(lp);
                logp_sum += lp;
                plogp = fmaf(p[k], lp, plogp);
                if (valid && k == (int)lab) { p_l = p[k]; ce -= (double)lp; }
            }
        if (valid) rce += (double)(1.f - p_l);
        reg += reg_kind == ONDA_REG_MRKLD ? -(double)logp_sum : (double)plogp;
        if (grad != nullptr) {
#pragma unroll
            for (int k = 0; k < CP; ++k)
                if (k < C) {
                    float gk = 0.f;
                    if (valid) {
                        const float d = p[k] - ((k == (int)lab) ? 1.f : 0.f);
                        gk = g_ce * d + g_rce * p_l * d;
                    }
                    if (reg_kind == ONDA_REG_MRKLD) gk += g_reg * ((float)C * p[k] - 1.f);
                    else if (reg_kind == ONDA_REG_MRENT) gk += g_reg * p[k] * ((z[k] - lse) - plogp);
                    grad[off + (unsigned)k * hw] = gk;
                }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v4[kLossSlots] = {ce, rce, reg, agree};
#pragma unroll
    for (int s = 0; s < kLossSlots; ++s) {
        double v = v4[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][s] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossSlots) {
        double v = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * kLossSlots + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        __threadfence();
        if (threadIdx.x < kLossSlots) {
            double v = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) v += __ldcg(partials + (size_t)c * kLossSlots + threadIdx.x);
            // out: ce | rce | reg | total | agreement | n_valid
            float r = 0.f;
            if (threadIdx.x == 0) r = (float)(v / (double)nv);
            if (threadIdx.x == 1) r = (float)(v * (double)log_floor / ((double)nv + 1e-6));
            if (threadIdx.x == 2) r = reg_kind == ONDA_REG_NONE ? 0.f : (float)(v / (reg_kind == ONDA_REG_MRKLD ? (double)N * C : (double)N));
            if (threadIdx.x == 3) r = (float)(v / (double)N);
            out[threadIdx.x == 3 ? 4 : threadIdx.x] = r;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            out[3] = alpha * out[0] + beta * out[1] + reg_weight * out[2];
            out[5] = nv;
            *ticket = 0u;
        }
    }
}

}  // namespace onda

using namespace onda;

extern "C" {

size_t onda_target_loss_workspace_bytes(void) { return 512 + (size_t)8 * cached_sm_count() * kLossSlots * sizeof(double); }

int onda_target_loss_fused(const float* student_logits, const int64_t* labels, int B, int C, int HW, const float* n_valid,
                           float alpha, float beta, float reg_weight, int regularizer, float* grad, float* out6,
                           void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ONDA_REQUIRE(student_logits && labels && out6 && workspace, "onda_target_loss_fused: null pointer");
    ONDA_REQUIRE(B > 0 && HW > 0 && C > 0 && C <= ONDA_MAX_CLASSES && (unsigned long long)B * C * HW < (1ull << 32),
                 "onda_target_loss_fused: bad shape");
    ONDA_REQUIRE(regularizer == ONDA_REG_NONE || regularizer == ONDA_REG_MRKLD || regularizer == ONDA_REG_MRENT,
                 "onda_target_loss_fused: unknown regularizer %d", regularizer);
    ONDA_REQUIRE(workspace_bytes >= onda_target_loss_workspace_bytes(), "onda_target_loss_fused: workspace too small");
    const long long N = (long long)B * HW;
    const int sms = cached_sm_count();
    long long want = (N + kLossThreads - 1) / kLossThreads;
    const int grid = (int)(want < 8LL * sms ? want : 8LL * sms);
    unsigned* ticket = (unsigned*)workspace;                                   // zero on first use; the kernel re-arms it
    unsigned long long* count = (unsigned long long*)((char*)workspace + 256);
    double* partials = (double*)((char*)workspace + 512);
    if (n_valid == nullptr) {                                                  // no count from the fused pass: count here
        ONDA_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(unsigned long long), stream));
        count_valid_kernel<<<grid, kLossThreads, 0, stream>>>((const long long*)labels, N, C, count);
        ONDA_CUDA_TRY(cudaGetLastError());
        count_launch(1);
    }
    if (padded_classes(C) == 20)
        target_loss_kernel<20><<<grid, kLossThreads, 0, stream>>>(student_logits, (const long long*)labels, B, C, HW, n_valid, count, alpha,
                                                                 beta, reg_weight, regularizer, grad, partials, ticket, out6);
    else
        target_loss_kernel<32><<<grid, kLossThreads, 0, stream>>>(student_logits, (const long long*)labels, B, C, HW, n_valid, count, alpha,
                                                                 beta, reg_weight, regularizer, grad, partials, ticket, out6);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

}  // extern "C"
