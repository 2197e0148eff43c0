// peer.cu -- the one exchange of the data-parallel path: a sum of per-rank scalars (the training loss) over
// NVLink peer memory.
//
// The reference has no multi-GPU code; SURVEY section 8(e) / BASELINE's north star give the path one collective,
// an all-reduce of the scalar loss.  Issued through torch.distributed/NCCL that 4-byte all-reduce cost +75 us per
// step at N = 2 on a step whose GPU work is ~265 us (bench, same box, with and without the collective).  Here
// every rank owns a small MAILBOX in its own HBM, exported with CUDA IPC and mapped by its peers:
//
//   post:  one thread block stores (value, tag) as ONE 8-byte word into the slot [turn][my_rank] of EVERY mailbox
//          -- plain st.global over NVLink for the peers, no handshake;
//   sum:   a later launch (the host reads the loss a step or more late, as a logger does) checks that all `world`
//          words of the turn carry the tag, waits for stragglers with a bounded spin, and writes the sum.
//
// An 8-byte naturally aligned store is single-copy atomic, so a word is never seen half-written; the tag (the step
// number) tells a fresh word from the one left by the previous lap around the ring of turns.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nerfacc_b200.h"

namespace nfa {

__global__ void mailbox_post_kernel(const float* __restrict__ value, unsigned long long* const* __restrict__ boxes,
                                    int32_t world, int32_t rank, int32_t turn, uint32_t tag)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    const unsigned long long word = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(*value);
    volatile unsigned long long* dst = boxes[r] + ((size_t)turn * world + rank);
    *dst = word;
    __threadfence_system();
}

__global__ void mailbox_sum_kernel(const unsigned long long* __restrict__ box, int32_t world, int32_t turn, uint32_t tag,
                                   float scale, float* __restrict__ out, int32_t* __restrict__ status)
{
    __shared__ float s_val[64];
    __shared__ int s_bad;
    const int r = threadIdx.x;
    if (r == 0) s_bad = 0;
    __syncthreads();
    if (r < world) {
        const volatile unsigned long long* src = box + ((size_t)turn * world + r);
        unsigned long long word = *src;
        // a straggler's word arrives within microseconds; ~0.5 s of polling means a rank died
        for (long spin = 0; (uint32_t)(word >> 32) != tag && spin < (1L << 24); ++spin) {
            __nanosleep(32);
            word = *src;
        }
        if ((uint32_t)(word >> 32) != tag) atomicExch(&s_bad, 1);
        s_val[r] = __uint_as_float((uint32_t)word);
    }
    __syncthreads();
    if (r == 0) {
        float acc = 0.f;
        for (int i = 0; i < world; ++i) acc += s_val[i];  // rank order: the same sum on every rank
        *out = s_bad ? __int_as_float(0x7fc00000) : acc * scale;
        if (status && s_bad) *status = 1;
    }
}

}  // namespace nfa

using namespace nfa;

extern "C" {

int32_t nfa_mailbox_create(int32_t world, int32_t turns, void** box, unsigned char* handle64)
{
    if (world <= 0 || world > 64 || turns <= 0 || !box || !handle64) return NFA_ERR_ARG;
    void* p = nullptr;
    const size_t bytes = (size_t)world * turns * sizeof(unsigned long long);
    cudaError_t e = cudaMalloc(&p, bytes);  // a plain allocation: pool / VMM memory cannot be exported with CUDA IPC
    if (e != cudaSuccess) return (int32_t)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) {
            static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
            for (int i = 0; i < 64; ++i) handle64[i] = ((const unsigned char*)&h)[i];
        }
    }
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int32_t)e;
    }
    *box = p;
    return NFA_OK;
}

int32_t nfa_mailbox_open(const unsigned char* handle64, void** peer_box)
{
    if (!handle64 || !peer_box) return NFA_ERR_ARG;
    cudaIpcMemHandle_t h;
    for (int i = 0; i < 64; ++i) ((unsigned char*)&h)[i] = handle64[i];
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int32_t)e;
    *peer_box = p;
    return NFA_OK;
}

int32_t nfa_mailbox_close(void* peer_box) { return peer_box ? (int32_t)cudaIpcCloseMemHandle(peer_box) : NFA_OK; }

int32_t nfa_mailbox_destroy(void* box) { return box ? (int32_t)cudaFree(box) : NFA_OK; }

int32_t nfa_mailbox_post(const float* value, const void* boxes, int32_t world, int32_t rank, int32_t turn, uint32_t tag,
                         nfa_stream_t stream)
{
    if (!value || !boxes || world <= 0 || world > 64 || rank < 0 || rank >= world || turn < 0) return NFA_ERR_ARG;
    mailbox_post_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(value, (unsigned long long* const*)boxes, world, rank, turn,
                                                            tag);
    return (int32_t)cudaGetLastError();
}

int32_t nfa_mailbox_sum(const void* box, int32_t world, int32_t turn, uint32_t tag, float scale, float* out,
                        int32_t* status, nfa_stream_t stream)
{
    if (!box || !out || world <= 0 || world > 64 || turn < 0) return NFA_ERR_ARG;
    mailbox_sum_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const unsigned long long*)box, world, turn, tag, scale, out,
                                                           status);
    return (int32_t)cudaGetLastError();
}

}  // extern "C"
