// cub_compare.cu — comparator only (not part of the product): cub::DeviceRadixSort::SortKeys on the same
// element count and bit range as K2 (64-bit elements, top 32 bits), for the "perf vs CUB" line in DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cub/cub.cuh>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void fill(uint64_t* p, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = i + 0x9e3779b97f4a7c15ULL; x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31; p[i] = x;
    }
}
int main() {
    const uint64_t n = 399889760ULL;
    uint64_t *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int begin_bit : {32, 24, 0}) {
        size_t tmp_bytes = 0; void* tmp = nullptr;
        CK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, a, b, n, begin_bit, 64));
        CK(cudaMalloc(&tmp, tmp_bytes));
        float best = 1e9;
        for (int it = 0; it < 4; it++) {
            fill<<<148 * 8, 256>>>(a, n);
            cudaEventRecord(e0);
            CK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, a, b, n, begin_bit, 64));
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) best = ms < best ? ms : best;
        }
        const int passes = (64 - begin_bit + 7) / 8;
        printf("CUB SortKeys u64 n=%llu bits[%d,64): %.3f ms total, %.3f ms per 8-bit pass (incl. histogram)\n",
               (unsigned long long)n, begin_bit, best, best / passes);
        cudaFree(tmp);
    }
    return 0;
}
