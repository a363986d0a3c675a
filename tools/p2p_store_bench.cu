// p2p_store_bench.cu — how fast can SMs push data into a peer GPU's memory over NVLink?  (design probe for the fused
// partition + exchange; not part of the product)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/p2p_store_bench tools/p2p_store_bench.cu && ./tools/p2p_store_bench
// Variants: plain coalesced 8-byte / 16-byte stores from registers, shared-memory staged runs of RUN bytes written like the
// partition kernel does (one run = one contiguous piece, warps stride over it), cp.async.bulk shared -> peer global, and the
// copy engine (cudaMemcpyPeerAsync) as the yardstick.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void st8(const uint64_t* __restrict__ in, uint64_t* out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void st16(const uint4* __restrict__ in, uint4* out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
// tile of 64 KB staged in shared memory, then written out in runs of `run` bytes to scattered destinations (run-granular permutation)
template <bool BULK>
__global__ void __launch_bounds__(512, 2) staged(const uint64_t* __restrict__ in, uint64_t* out, size_t n_tiles, uint32_t run_elems) {
    extern __shared__ __align__(128) uint64_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t T = 8192;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < T; i += 512) sm[i] = in[t * T + i];
        __syncthreads();
        const uint32_t runs = T / run_elems;
        if (BULK) {
            // one elected lane per run issues a bulk copy shared -> global
            for (uint32_t r = threadIdx.x; r < runs; r += 512) {
                const size_t dst_run = (t * runs + (r * 7919u) % runs);      // permute runs inside the tile's output window
                uint64_t* dst = out + dst_run * run_elems;
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm + r * run_elems);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(run_elems * 8) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else {
            for (uint32_t i = threadIdx.x; i < T; i += 512) {
                const uint32_t r = i / run_elems, o = i % run_elems;
                const size_t dst_run = (t * runs + (r * 7919u) % runs);
                out[dst_run * run_elems + o] = sm[i];
            }
        }
    }
    (void)bar;
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t bytes = 2ull << 30, n = bytes / 8;
    uint64_t *src, *dst_peer, *dst_local;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&dst_peer, bytes));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst_local, bytes)); CK(cudaMemset(src, 1, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto report = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9); };
    float ms;
    for (int peer = 0; peer < 2; peer++) {
        uint64_t* dst = peer ? dst_peer : dst_local;
        const char* where = peer ? "peer " : "local";
        char nm[128];
#define TIME(label, launch) do { launch; CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); for (int r = 0; r < 3; r++) { launch; } CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms, e0, e1)); snprintf(nm, sizeof nm, "%s %s", where, label); report(nm, ms / 3); } while (0)
        TIME("st.u64 coalesced, 148x8 CTAs x 512", (st8<<<148 * 8, 512>>>(src, dst, n)));
        TIME("st.v4.u32 coalesced, 148x8 CTAs x 512", (st16<<<148 * 8, 512>>>((const uint4*)src, (uint4*)dst, n / 2)));
        TIME("st.v4.u32 coalesced, 148x2 CTAs x 1024", (st16<<<148 * 2, 1024>>>((const uint4*)src, (uint4*)dst, n / 2)));
        CK(cudaFuncSetAttribute(staged<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(staged<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        const size_t n_tiles = n / 8192;
        for (uint32_t run : {16u, 32u, 256u, 1024u}) {
            snprintf(nm, sizeof nm, "staged tile, plain stores, runs of %u B", run * 8);
            char lab[96]; snprintf(lab, sizeof lab, "%s", nm);
            TIME(lab, (staged<false><<<(unsigned)n_tiles, 512, 65536>>>(src, dst, n_tiles, run)));
            snprintf(lab, sizeof lab, "staged tile, cp.async.bulk, runs of %u B", run * 8);
            TIME(lab, (staged<true><<<(unsigned)n_tiles, 512, 65536>>>(src, dst, n_tiles, run)));
        }
        if (peer) TIME("cudaMemcpyPeerAsync (copy engine)", (cudaMemcpyPeerAsync(dst, 1, src, 0, bytes)));
    }
    return 0;
}
