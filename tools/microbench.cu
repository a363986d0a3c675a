// microbench.cu — design probes for the bucketed path (not part of the product):
//  T1: one-pass scatter of n 8-byte records into NB fixed-capacity buckets, slot claimed by a global atomic per record
//  T3: per-bucket shared-memory radix sort (2 x 8-bit passes, atomic ranking) of those buckets
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31; return x; }

template <int MODE>
__global__ void scatter_kernel(uint64_t n, uint32_t NB, uint32_t cap, uint32_t* cur, uint64_t* out, unsigned long long* overflow) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // 40 copies of each key (like 40 genomes sharing a flank), consecutive threads = different keys
        const uint64_t key = mix64(i % (n / 40) + 12345) ;
        const uint64_t rec = (key & ~0xFFULL) | (i / (n / 40));
        const uint32_t b = __umulhi((uint32_t)(key >> 32), NB);
        if (MODE == 0) {
            const uint32_t slot = atomicAdd(&cur[b], 1u);
            if (slot < cap) out[(uint64_t)b * cap + slot] = rec; else atomicAdd(overflow, 1ULL);
        } else if (MODE == 1) {        // store only (no atomic): pseudo slot
            const uint32_t slot = (uint32_t)(i / NB) % cap;
            out[(uint64_t)b * cap + slot] = rec;
        } else {                        // atomic only
            const uint32_t slot = atomicAdd(&cur[b], 1u);
            if (slot == 0xFFFFFFFFu) out[0] = rec;
        }
    }
}

// one CTA per bucket (grid-strided): load, 2 smem radix passes with atomic ranking, checksum out
template <int THREADS, int CAP>
__global__ void __launch_bounds__(THREADS) bucket_sort_kernel(uint32_t NB, const uint32_t* cnt, const uint64_t* in, uint64_t* sums, int passes) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* A = reinterpret_cast<uint64_t*>(smem);
    uint64_t* B = A + CAP;
    uint32_t* hist = reinterpret_cast<uint32_t*>(B + CAP);   // 256
    __shared__ uint32_t wsum[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t b = blockIdx.x; b < NB; b += gridDim.x) {
        const uint32_t n = min(cnt[b], (uint32_t)CAP);
        const uint64_t* src = in + (uint64_t)b * CAP;
        for (uint32_t i = tid; i < n; i += THREADS) A[i] = src[i];
        uint64_t* cur = A; uint64_t* alt = B;
        for (int p = 0; p < passes; p++) {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            const uint32_t shift = 24 + 8 * p;
            // rank = atomicAdd return (unstable), kept in registers: up to CAP/THREADS items per thread
            uint32_t rk[CAP / THREADS];
#pragma unroll
            for (int j = 0; j < CAP / THREADS; j++) {
                const uint32_t i = j * THREADS + tid;
                if (i < n) rk[j] = atomicAdd(&hist[(uint32_t)(cur[i] >> shift) & 255u], 1u);
            }
            __syncthreads();
            uint32_t c = 0, x = 0;
            if (tid < 256) { c = hist[tid]; x = c; }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x += y; }
            if (tid < 256 && lane == 31) wsum[warp] = x;
            __syncthreads();
            if (tid < 256) { uint32_t add = 0; for (uint32_t w = 0; w < warp; w++) add += wsum[w]; hist[tid] = add + x - c; }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < CAP / THREADS; j++) {
                const uint32_t i = j * THREADS + tid;
                if (i < n) { const uint64_t k = cur[i]; alt[hist[(uint32_t)(k >> shift) & 255u] + rk[j]] = k; }
            }
            __syncthreads();
            uint64_t* t = cur; cur = alt; alt = t;
        }
        uint64_t s = 0;
        for (uint32_t i = tid; i < n; i += THREADS) s += cur[i] * (i + 1);
        if (s == 0x1234567) sums[b] = s;
        __syncthreads();
    }
}

int main() {
    const uint64_t n = 400000000ULL;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t NB : {65536u, 16384u, 4096u}) {
        const uint32_t cap = (uint32_t)(n / NB * 1.25) + 64;
        uint32_t* cur; uint64_t* out; unsigned long long* ov;
        CK(cudaMalloc(&cur, NB * 4)); CK(cudaMalloc(&out, (uint64_t)NB * cap * 8)); CK(cudaMalloc(&ov, 8));
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e9;
            for (int it = 0; it < 3; it++) {
                CK(cudaMemset(cur, 0, NB * 4)); CK(cudaMemset(ov, 0, 8));
                cudaEventRecord(e0);
                if (mode == 0) scatter_kernel<0><<<148 * 8, 512>>>(n, NB, cap, cur, out, ov);
                else if (mode == 1) scatter_kernel<1><<<148 * 8, 512>>>(n, NB, cap, cur, out, ov);
                else scatter_kernel<2><<<148 * 8, 512>>>(n, NB, cap, cur, out, ov);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
            }
            unsigned long long h_ov; CK(cudaMemcpy(&h_ov, ov, 8, cudaMemcpyDeviceToHost));
            printf("T1 NB=%u cap=%u mode=%s: %.3f ms  (%.1f Grec/s) overflow=%llu\n", NB, cap,
                   mode == 0 ? "atomic+store" : (mode == 1 ? "store-only" : "atomic-only"), best, n / best / 1e6, h_ov);
        }
        if (NB == 65536u) {
            // T3 on the buckets just written (mode 0 result was overwritten by mode 1/2; redo mode 0)
            CK(cudaMemset(cur, 0, NB * 4));
            scatter_kernel<0><<<148 * 8, 512>>>(n, NB, cap, cur, out, ov);
            CK(cudaDeviceSynchronize());
            constexpr int CAP = 8192;
            if (cap <= CAP) {
                uint64_t* sums; CK(cudaMalloc(&sums, NB * 8));
                // note: layout stride is `cap`, kernel assumes CAP: re-scatter with cap = CAP
                CK(cudaFree(out)); CK(cudaMalloc(&out, (uint64_t)NB * CAP * 8));
                CK(cudaMemset(cur, 0, NB * 4));
                scatter_kernel<0><<<148 * 8, 512>>>(n, NB, CAP, cur, out, ov);
                CK(cudaDeviceSynchronize());
                for (int passes : {0, 1, 2, 3}) {
                    auto run = [&](auto kern, int threads, int grid, const char* name) -> int {
                        size_t smem = (size_t)CAP * 16 + 1024;
                        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        float best = 1e9;
                        for (int it = 0; it < 3; it++) {
                            cudaEventRecord(e0);
                            kern<<<grid, threads, smem>>>(NB, cur, out, sums, passes);
                            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
                        }
                        printf("T3 %s passes=%d: %.3f ms (%.1f Grec/s)\n", name, passes, best, n / best / 1e6);
                        return 0;
                    };
                    if (run(bucket_sort_kernel<1024, CAP>, 1024, 148, "1024thr x1/SM")) return 1;
                    if (run(bucket_sort_kernel<512, CAP>, 512, 148, "512thr x1/SM")) return 1;
                }
                cudaFree(sums);
            }
        }
        cudaFree(cur); cudaFree(out); cudaFree(ov);
    }
    return 0;
}
