// smem_bench.cu — shared-memory pipe probes on B200 (design tool, not product): what one rank atomic / presence RED / scattered
// store / table look-up costs per 32 records when 2 x 512 threads per SM hammer random addresses, and what a tagged
// load-store-verify update (no atomic, warp-private table) costs instead.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/smem_bench tools/smem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

#define THREADS 512
#define ITERS 4096

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// MODE 0: ALU only (the random-number overhead)        1: atomicAdd with return on 256 counters
//      2: red.or (no return) on 1024 words             3: LDS.32 random in 256 words
//      4: STS.64 random scatter into 8192 slots        5: LDS.64 random in 512 slots
//      6: tagged load / store / verify on a warp-private 256-counter histogram (no atomic)
//      7: 1 + 3 + 4 together (one partition record)    8: atomicAdd with return on 65536 packed 16-bit counters (128 KB)
//      9: red.or on 64-bit presence (two 32-bit halves, like K3) + LDS.64 key
template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) probe(uint32_t* out, int iters) {
    extern __shared__ __align__(16) uint32_t sm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 16384 + 2048; i += THREADS) sm[i] = 0;
    __syncthreads();
    uint32_t s = tid * 2654435761u + blockIdx.x * 97u + 1u, acc = 0;
    uint64_t* sm64 = reinterpret_cast<uint64_t*>(sm);
    uint32_t* wh = sm + 16384 + warp * 0;   // (mode 6 uses a per-warp region below)
    (void)wh;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = lcg(s);
        if (MODE == 0) acc += r;
        if (MODE == 1 || MODE == 7) acc += atomicAdd(&sm[r & 255u], 1u);
        if (MODE == 2) asm volatile("red.shared.or.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&sm[r & 1023u])), "r"(1u << (r >> 27)) : "memory");
        if (MODE == 3 || MODE == 7) acc += *reinterpret_cast<volatile uint32_t*>(&sm[256 + ((r >> 8) & 255u)]);
        if (MODE == 4 || MODE == 7) *reinterpret_cast<volatile uint64_t*>(&sm64[512 + ((r >> 4) & 4095u)]) = ((uint64_t)r << 32) | acc;
        if (MODE == 5) acc += (uint32_t)*reinterpret_cast<volatile uint64_t*>(&sm64[(r >> 3) & 511u]);
        if (MODE == 6) {
            uint32_t* h = sm + warp * 256;                      // warp-private 256 counters
            const uint32_t d = r & 255u;
            bool done = false;
            uint32_t rank = 0;
            while (true) {
                uint32_t old = 0;
                if (!done) {
                    old = *reinterpret_cast<volatile uint32_t*>(&h[d]);
                    *reinterpret_cast<volatile uint32_t*>(&h[d]) = ((old + 1u) & 0xFFFFu) | (lane << 16);
                }
                __syncwarp();
                if (!done && (*reinterpret_cast<volatile uint32_t*>(&h[d]) >> 16) == lane) { done = true; rank = old & 0xFFFFu; }
                __syncwarp();
                if (__all_sync(0xFFFFFFFFu, done)) break;
            }
            acc += rank;
        }
        if (MODE == 8) { const uint32_t b = r & 65535u; acc += atomicAdd(&sm[(b >> 1) & 16383u], 1u << ((b & 1u) << 4)); }
        if (MODE == 9) {
            const uint32_t slot = (r >> 3) & 511u;
            acc += (uint32_t)*reinterpret_cast<volatile uint64_t*>(&sm64[slot]);
            asm volatile("red.shared.or.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&sm[8192 + slot * 2 + ((r >> 20) & 1u)])), "r"(1u << (r >> 27)) : "memory");
        }
    }
    if (acc == 0x12345678u) out[blockIdx.x * THREADS + tid] = acc;
}

template <int MODE>
int run(const char* name, uint32_t* out, int n_sm, double ghz_guess) {
    const size_t smem = (16384 + 2048) * 4 + (MODE == 8 ? 65536 : 0);
    CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<MODE><<<n_sm * 2, THREADS, smem>>>(out, ITERS);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warp_ops = (double)2 * (THREADS / 32) * ITERS;               // per SM
    const double cyc = best * 1e-3 * ghz_guess * 1e9;
    printf("%-58s %8.3f ms  %7.2f cycles per warp-op per SM (at %.2f GHz)  = %6.2f ms per 4e8 records over 148 SMs\n", name, best, cyc / warp_ops, ghz_guess,
           4e8 / 32.0 / 148.0 * (cyc / warp_ops) / (ghz_guess * 1e9) * 1e3);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int n_sm = p.multiProcessorCount;
    const double ghz = p.clockRate / 1e6;
    uint32_t* out; CK(cudaMalloc(&out, (size_t)n_sm * 2 * THREADS * 4));
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, n_sm, ghz);
    if (run<0>("0 ALU only (LCG)", out, n_sm, ghz)) return 1;
    if (run<1>("1 atomicAdd+return, 256 counters", out, n_sm, ghz)) return 1;
    if (run<8>("8 atomicAdd+return, 64K packed 16-bit counters", out, n_sm, ghz)) return 1;
    if (run<2>("2 red.or, 1024 words", out, n_sm, ghz)) return 1;
    if (run<3>("3 LDS.32 random, 256 words", out, n_sm, ghz)) return 1;
    if (run<4>("4 STS.64 random scatter, 4096 slots", out, n_sm, ghz)) return 1;
    if (run<5>("5 LDS.64 random, 512 slots", out, n_sm, ghz)) return 1;
    if (run<6>("6 tagged load/store/verify, warp-private 256 counters", out, n_sm, ghz)) return 1;
    if (run<7>("7 atomicAdd + LDS.32 + STS.64 (one partition record)", out, n_sm, ghz)) return 1;
    if (run<9>("9 LDS.64 key + red.or presence (one K3 record)", out, n_sm, ghz)) return 1;
    return 0;
}
