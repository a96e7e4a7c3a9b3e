// One-shot sum all-reduce of the small `sums` buffer over NVLink peer memory.
//
// Every rank's input lives in a symmetric (peer-mapped) buffer.  The kernel (1) publishes "my input is
// ready" by storing the call's epoch into a flag word on every peer, (2) waits until all peers have
// published theirs, then (3) reads all inputs over NVLink and adds them in rank order 0..world-1, so every
// rank computes bit-identical sums without a second exchange.  Payload 39-311 KB: latency-bound, one
// launch, no NCCL ring/tree.  Inputs are double-buffered by the caller (slot = call parity), so a slot is
// only rewritten after every peer has passed the next call's handshake, i.e. finished reading it.
#include "common.cuh"

namespace onda {

__global__ void __launch_bounds__(256) allreduce_oneshot_kernel(float* __restrict__ out, size_t n, int rank, int world,
                                                                PeerTable peers, uint32_t epoch) {
    peer_handshake(peers, rank, world, epoch);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[i] = peer_sum(peers, world, i);
    }
}

int launch_allreduce_oneshot(float* out, size_t n, int rank, int world, void* const* bufs, void* const* flags,
                             uint32_t epoch, cudaStream_t stream) {
    const PeerTable t = make_peer_table(rank, world, bufs, flags);
    const int threads = 256;
    int blocks = (int)((n + threads - 1) / threads);
    if (blocks > 64) blocks = 64;
    if (blocks < 1) blocks = 1;
    allreduce_oneshot_kernel<<<blocks, threads, 0, stream>>>(out, n, rank, world, t, epoch);
    ONDA_CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return ONDA_OK;
}

}  // namespace onda
