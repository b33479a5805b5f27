// Micro-benchmark: how fast can one CTA per SM push shared memory to global memory?
//   mode 0: cp.async.bulk (TMA engine) of `piece` bytes, `inflight` groups outstanding, issued by one thread
//   mode 1: st.global.v4 by 256 threads, coalesced 16 B per thread
// usage: bulk_store_bw <mode> <piece_bytes> <inflight> <MB_per_cta>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(unsigned char *dst, size_t per_cta, int mode, int piece, int inflight)
{
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < 131072 / 4; i += 256) reinterpret_cast<uint32_t *>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned char *base = dst + (size_t)blockIdx.x * per_cta;
    if (mode == 0) {
        if (threadIdx.x == 0) {
            int pending = 0;
            for (size_t off = 0; off < per_cta; off += piece) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(smem_u32(smem + (off % 131072))), "r"(piece) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (++pending >= inflight) {
                    if (inflight == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else if (inflight == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else if (inflight == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
                    --pending;
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        for (size_t off = (size_t)threadIdx.x * 16; off < per_cta; off += 256 * 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(smem + (off % 131072));
            *reinterpret_cast<uint4 *>(base + off) = v;
        }
    }
}
int main(int argc, char **argv)
{
    const int mode = atoi(argv[1]), piece = atoi(argv[2]), inflight = atoi(argv[3]);
    const size_t per_cta = (size_t)atoi(argv[4]) << 20;
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    unsigned char *d; cudaMalloc(&d, per_cta * nsm);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k<<<nsm, 256, 131072>>>(d, per_cta, mode, piece, inflight);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep == 2) printf("mode %d piece %6d inflight %d: %.1f GB/s (%.2f B/clk/SM at 1.965 GHz) err=%s\n", mode, piece, inflight,
                             per_cta * nsm / ms / 1e6, per_cta / (ms * 1e-3) / 1.965e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
