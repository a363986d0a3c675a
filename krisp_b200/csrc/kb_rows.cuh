// kb_rows.cuh — CSV rows of the survivors, rendered and ordered on the device.
//
// Replaces ConservedEndAmplicons.render_csv (krisp_fasta/Amplicon.py:663-671) with consensus (:550-558) and collapse_to_iupac
// (:42-66) for every survivor, and the row order of render_output with --cores 1 (outputAlignments.py:101-162: ascending
// (left, right), the order of the sorted k-mer file).  One row = left ',' consensus ',' right '\n' (fixed width L + D + R + 3):
// bases come from the survivor's flank words (MSB-first 2-bit codes), the consensus letter of column c from its 4-bit base set
// (bit 0 A .. bit 3 T; the inverse of Bio.Data.IUPACData.ambiguous_dna_values as Amplicon.py:10-12 builds it, 4 bases -> N):
// the ingroup set when an outgroup was given, every occurrence otherwise (krisp_fasta.py:282-283).
// Order: the survivors' indices are sorted by their flank words with the chunked LSD sort of kb_sort.cuh (kb_chunk_key_kernel).
#pragma once
#include "kb_common.cuh"

struct KbRowsArgs {
    const uint32_t* rank;        // != null: [n] row of survivor g (kb_rank_kernel); else `order` is used
    const uint64_t* order;       // [n] sorted elements, low 32 bits = survivor index
    uint64_t n;
    const uint64_t* flank;       // [n][FW]
    const uint32_t* in_mask;     // [n][MW]
    const uint32_t* out_mask;    // [n][MW]
    int L, D, R, FW, MW;
    int all_occurrences;         // 1: consensus over ingroup and outgroup (no --outgroup given)
    char* out;                   // [n][L + D + R + 3]
};

__global__ void __launch_bounds__(256) kb_rows_kernel(const KbRowsArgs a) {
    const uint32_t width = (uint32_t)(a.L + a.D + a.R + 3);
    const uint64_t total = a.n * width;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const uint64_t r = t / width;
        const uint32_t p = (uint32_t)(t % width);
        const uint64_t g = a.rank ? r : (a.order[r] & 0xFFFFFFFFULL);           // ranks: walk the survivors, write row rank[g]
        const uint64_t dst = a.rank ? (uint64_t)a.rank[r] * width + p : t;
        char ch;
        if (p == (uint32_t)a.L || p == (uint32_t)(a.L + 1 + a.D)) ch = ',';
        else if (p == width - 1) ch = '\n';
        else if (p > (uint32_t)a.L && p < (uint32_t)(a.L + 1 + a.D)) {
            const uint32_t c = p - (uint32_t)a.L - 1;
            uint32_t m = a.in_mask[g * a.MW + (c >> 3)];
            if (a.all_occurrences) m |= a.out_mask[g * a.MW + (c >> 3)];
            ch = "?ACMGRSVTWYHKDBN"[(m >> (28 - 4 * (c & 7))) & 0xFu];
        } else {
            const uint32_t b = p < (uint32_t)a.L ? p : p - 2 - (uint32_t)a.D;      // base index inside the flank bits
            const uint64_t w = a.flank[g * a.FW + (b >> 5)];
            ch = "ACGT"[(w >> (62 - 2 * (b & 31))) & 3ULL];
        }
        a.out[dst] = ch;
    }
}

__global__ void __launch_bounds__(256) kb_iota_kernel(uint64_t* ent, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) ent[i] = i;
}

// Few survivors (the usual case: thousands of rows out of 1e8 records): their order by flank words comes from counting, one launch
// instead of the chunked LSD sort's twenty — rank(i) = #{j : flank[j] < flank[i]} (+ ties by index; flank keys of one search are
// distinct anyway).  Every thread owns one survivor and walks a slice of all of them through shared memory (all lanes read the same
// element: broadcast; the slices of the CTA rows add up with one atomic each).  n <= KB_RANK_MAX keeps the n^2 walk in the microseconds.
#define KB_RANK_MAX 8192
#define KB_RANK_THREADS 128

struct KbRankArgs {
    const uint64_t* flank;       // [n][FW]
    uint32_t n;
    uint32_t slice;              // survivors j every CTA row (blockIdx.y) compares against: [y * slice, (y + 1) * slice)
    uint32_t* rank;              // [n] zeroed; every CTA row adds its share
};

template <int FW>
__global__ void __launch_bounds__(KB_RANK_THREADS) kb_rank_kernel(const KbRankArgs a) {
    __shared__ uint64_t tile[FW][KB_RANK_THREADS];
    const uint32_t tid = threadIdx.x, i = blockIdx.x * KB_RANK_THREADS + tid;
    uint64_t mine[FW];
#pragma unroll
    for (int w = 0; w < FW; w++) mine[w] = i < a.n ? a.flank[(uint64_t)i * FW + w] : 0ULL;
    uint32_t rank = 0;
    const uint32_t j0 = blockIdx.y * a.slice, j1 = min(a.n, j0 + a.slice);
    for (uint32_t base = j0; base < j1; base += KB_RANK_THREADS) {
        __syncthreads();
        const uint32_t j = base + tid;
#pragma unroll
        for (int w = 0; w < FW; w++) tile[w][tid] = j < j1 ? a.flank[(uint64_t)j * FW + w] : 0ULL;
        __syncthreads();
        const uint32_t cnt = min((uint32_t)KB_RANK_THREADS, j1 - base);
#pragma unroll 8
        for (uint32_t jj = 0; jj < cnt; jj++) {
            bool less = false, eq = true;
#pragma unroll
            for (int w = 0; w < FW; w++) {
                const uint64_t t = tile[w][jj];
                if (eq && t != mine[w]) { less = t < mine[w]; eq = false; }
            }
            rank += (less || (eq && base + jj < i)) ? 1u : 0u;
        }
    }
    if (i < a.n && rank) atomicAdd(a.rank + i, rank);
}

// The survivor table's columns (flank words, base sets, sizes, runs) copied into ONE staging image laid out like the host's result
// arena, next to the rows text: the whole result then leaves the device as a single copy.
#define KB_PACK_SEGS 5
struct KbPackArgs {
    const uint32_t* src[KB_PACK_SEGS];
    uint64_t dst_off[KB_PACK_SEGS];      // byte offset in the image (multiple of 4)
    uint64_t words[KB_PACK_SEGS];        // 32-bit words to copy (0: segment unused)
    uint8_t* image;
};

__global__ void __launch_bounds__(256) kb_pack_kernel(const KbPackArgs a) {
#pragma unroll
    for (int s = 0; s < KB_PACK_SEGS; s++) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.image + a.dst_off[s]);
        for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < a.words[s]; t += (uint64_t)gridDim.x * 256) dst[t] = a.src[s][t];
    }
}
