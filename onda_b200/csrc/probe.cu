// Measurement tool, not part of the product path: a kernel that only performs the fused pass's feature loads
// (same persistent tile walk, same warp-to-(pixel quarter, channel chunk) mapping, same 4-byte strided loads
// along the NCHW channel planes) and folds them into one float per warp.  Its bandwidth is the ceiling this
// access pattern can reach on the machine -- what bench.py's roofline compares the real kernel with besides the
// streaming-copy peak.  `smem_bytes` of dynamic shared memory are requested to reproduce the real kernel's
// L1 size (L1 = 256 KB minus the shared-memory carve-out).
#include "common.cuh"

namespace onda {

template <int kLoads>
__global__ void __launch_bounds__(1024, 1) load_probe_kernel(const float* __restrict__ feat, int D, int HW, int N, int tiles,
                                                           int workers, float* __restrict__ out) {
    extern __shared__ unsigned char probe_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= workers) return;
    const int quarter = warp & 3, group = warp >> 2, groups = workers >> 2;
    const int NB = D / 32;
    const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * NB;
    float acc = 0.f;
    float x[2][kLoads];
    auto issue = [&](int q, float (&dst)[kLoads]) {
        const int t = q / NB, b = q - t * NB;
        const unsigned tile = blockIdx.x + (unsigned)t * gridDim.x;
        unsigned n = tile * 128u + 32u * quarter + lane;
        n = n < (unsigned)N ? n : (unsigned)N - 1;
        const unsigned bimg = n / (unsigned)HW, pix = n - bimg * (unsigned)HW;
        const float* src = feat + ((size_t)bimg * D + (size_t)b * 32) * HW + pix;
#pragma unroll
        for (int j = 0; j < kLoads; ++j) dst[j] = ldg_stream(src + (size_t)j * HW);
    };
    if (group < total) issue(group, x[0]);
    int cur = 0;
    for (int q = group; q < total; q += groups) {
        if (q + groups < total) {
            if (cur == 0) issue(q + groups, x[1]); else issue(q + groups, x[0]);
        }
        if (cur == 0) {
#pragma unroll
            for (int j = 0; j < kLoads; ++j) acc += x[0][j];
        } else {
#pragma unroll
            for (int j = 0; j < kLoads; ++j) acc += x[1][j];
        }
        cur ^= 1;
    }
    if (lane == 0) out[blockIdx.x * 32 + warp] = acc + (float)probe_smem[0] * 0.f;
}

}  // namespace onda

extern "C" int onda_debug_load_probe(const float* feat, int B, int D, int HW, int workers, int smem_bytes, float* out,
                                     void* stream) {
    using namespace onda;
    ONDA_REQUIRE(D % 32 == 0 && workers >= 4 && workers <= 32 && workers % 4 == 0, "load probe: bad shape (D %d, workers %d)", D, workers);
    const int N = B * HW, tiles = (N + 127) / 128;
    int sms = 0;
    ONDA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = tiles < sms ? tiles : sms;
    auto kern = load_probe_kernel<32>;
    ONDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<grid, 1024, smem_bytes, (cudaStream_t)stream>>>(feat, D, HW, N, tiles, workers, out);
    ONDA_CUDA_TRY(cudaGetLastError());
    return ONDA_OK;
}
