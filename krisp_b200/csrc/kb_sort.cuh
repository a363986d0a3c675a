// kb_sort.cuh — K2: onesweep LSD radix sort of 64-bit sort elements.
//
// Replaces the reference's external GNU sort (kstream/kstream.py:83-119,
// `LC_ALL=C sort -t, -k1,1 -k3,3`).  The reference needs the sort only to bring equal
// (left,right) keys together for the merge (shared.py:442-475); here the elements are ordered by
// the top 8*P bits of the sort element (mixed flank key, or flank hash in INDIRECT mode) with P
// stable 8-bit-digit passes, least significant digit first.
//
//   kb_hist_kernel  one read of the elements: all P digit histograms at once (shared-memory
//                   histograms, one flush of P*256 global atomics per CTA)
//   kb_scan_kernel  exclusive scan of each 256-bin histogram -> global digit offsets
//   kb_onesweep_kernel  one pass = ONE read + ONE write of every element: per-warp digit counters
//                   and ballot/match ranking (no atomics), per-tile digit counts chained between
//                   tiles by decoupled look-back on a status word per (tile, digit), elements
//                   staged in shared memory in ranked order so that each digit's run leaves the
//                   CTA as a contiguous, coalesced store.
#pragma once
#include "kb_common.cuh"

#define KB_RADIX_BITS 8
#define KB_RADIX 256

// digit of a sort element: radix digit (shard_n == 0) or destination shard (multi-GPU partition:
// the low 16 bits of the mixed flank key, scaled to [0, shard_n))
__device__ __forceinline__ uint32_t kb_digit(uint64_t key, uint32_t shift, uint32_t shard_n) {
    if (shard_n == 0) return (uint32_t)(key >> shift) & 255u;
    return __umulhi(((uint32_t)(key >> shift) & 0xFFFFu) << 16, shard_n);
}

// ---- histogram ---------------------------------------------------------------------------------
#define KB_HIST_THREADS 512
struct KbHistArgs {
    const uint64_t* in;
    uint64_t n;
    int P;                 // passes; pass j uses bits [shift0 + 8j, shift0 + 8j + 8)
    uint32_t shift0;
    uint32_t shard_n;      // != 0: P = 1 and the digit is the destination shard (see kb_digit)
    unsigned long long* hist;   // [P][256], zeroed
};

__global__ void __launch_bounds__(KB_HIST_THREADS) kb_hist_kernel(const KbHistArgs a) {
    __shared__ uint32_t sh[8 * KB_RADIX];
    for (uint32_t i = threadIdx.x; i < (uint32_t)a.P * KB_RADIX; i += KB_HIST_THREADS) sh[i] = 0;
    __syncthreads();
    const uint64_t npair = a.n >> 1;
    const uint64_t stride = (uint64_t)gridDim.x * KB_HIST_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * KB_HIST_THREADS + threadIdx.x; i < npair; i += stride) {
        const uint4 v = kb_ld_stream128(a.in + 2 * i);
        const uint64_t k0 = (uint64_t)v.x | ((uint64_t)v.y << 32), k1 = (uint64_t)v.z | ((uint64_t)v.w << 32);
        for (int p = 0; p < a.P; p++) {
            atomicAdd(&sh[p * KB_RADIX + kb_digit(k0, a.shift0 + 8 * p, a.shard_n)], 1u);
            atomicAdd(&sh[p * KB_RADIX + kb_digit(k1, a.shift0 + 8 * p, a.shard_n)], 1u);
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t k0 = a.in[a.n - 1];
        for (int p = 0; p < a.P; p++) atomicAdd(&sh[p * KB_RADIX + kb_digit(k0, a.shift0 + 8 * p, a.shard_n)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (uint32_t)a.P * KB_RADIX; i += KB_HIST_THREADS) {
        const uint32_t c = sh[i];
        if (c) atomicAdd(a.hist + i, (unsigned long long)c);
    }
}

// one CTA of 256 threads per pass: hist[p][d] -> exclusive prefix over d
__global__ void __launch_bounds__(KB_RADIX) kb_scan_kernel(unsigned long long* hist) {
    __shared__ unsigned long long ws[8];
    unsigned long long* h = hist + (size_t)blockIdx.x * KB_RADIX;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long v = h[threadIdx.x];
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned long long add = 0;
    for (uint32_t w = 0; w < warp; w++) add += ws[w];
    h[threadIdx.x] = add + x - v;
}

// ---- onesweep pass -----------------------------------------------------------------------------
// status word: value | flag in the top two bits
template <typename T> struct KbStatus;
template <> struct KbStatus<uint32_t> {
    static constexpr uint32_t LOCAL = 1u << 30, GLOBAL = 2u << 30, FLAGS = 3u << 30, VALUE = ~(3u << 30);
};
template <> struct KbStatus<unsigned long long> {
    static constexpr unsigned long long LOCAL = 1ULL << 62, GLOBAL = 2ULL << 62, FLAGS = 3ULL << 62, VALUE = ~(3ULL << 62);
};

template <typename ST>
struct KbSortArgs {
    const uint64_t* in;
    uint64_t* out;
    uint64_t n;
    uint32_t shift;                    // digit = kb_digit(key, shift, shard_n)
    uint32_t shard_n;
    const unsigned long long* base;    // [256] exclusive global digit offsets of this pass
    ST* status;                        // [n_tiles][256], zeroed
    uint32_t* ticket;                  // zeroed; dynamic tile ids keep look-back deadlock-free
};

__device__ __forceinline__ uint32_t kb_ld_relaxed(const uint32_t* p) {
    uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long kb_ld_relaxed(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void kb_st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void kb_st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// lanes of the warp holding the same 8-bit digit
template <bool HWMATCH>
__device__ __forceinline__ uint32_t kb_match_digit(uint32_t d) {
    if constexpr (HWMATCH) return __match_any_sync(0xFFFFFFFFu, d);
    // one ballot per digit bit; a lane keeps the ballot if its own bit is set, the complement otherwise
    uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < KB_RADIX_BITS; b++) {
        uint32_t m;
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                     "and.b32 t, %1, %2;\n\t"
                     "setp.ne.u32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                     "@!p not.b32 %0, %0;\n\t}"
                     : "=r"(m) : "r"(d), "r"(1u << b));
        peers &= m;
    }
    return peers;
}

template <typename ST, int THREADS, int ITEMS, int MINB, bool HWMATCH>
__global__ void __launch_bounds__(THREADS, MINB) kb_onesweep_kernel(const KbSortArgs<ST> a) {
    using S = KbStatus<ST>;
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    static_assert(THREADS >= KB_RADIX && TILE < 65536, "one thread per digit; 16-bit ranks");
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(kb_smem_raw);                         // TILE
    uint64_t* dbase = skeys + TILE;                                                     // 256: global offset - local start
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(dbase + KB_RADIX);                     // WARPS * 256
    uint32_t* wsum = wcnt + WARPS * KB_RADIX;                                           // 8 (scan scratch)
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
#pragma unroll
    for (int j = 0; j < KB_RADIX / 32; j++) wcnt[warp * KB_RADIX + j * 32 + lane] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_start = (uint64_t)tile * TILE;
    if (tile_start >= a.n) return;
    const uint32_t n_tile = (uint32_t)min((uint64_t)TILE, a.n - tile_start);

    // ---- load (warp-striped: stable order = warp, item, lane) ---------------------------------
    uint64_t key[ITEMS];
    const uint32_t wbase = warp * (ITEMS * 32) + lane;
    if (n_tile == TILE) {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) key[i] = kb_ld_stream(a.in + tile_start + wbase + i * 32);
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t idx = wbase + i * 32;
            key[i] = idx < n_tile ? kb_ld_stream(a.in + tile_start + idx) : ~0ULL;
        }
    }

    // ---- rank inside the warp -------------------------------------------------------------------
    uint32_t* mycnt = wcnt + warp * KB_RADIX;
    uint16_t rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t d = kb_digit(key[i], a.shift, a.shard_n);
        const uint32_t peers = kb_match_digit<HWMATCH>(d);
        const uint32_t before = mycnt[d];
        __syncwarp();
        if ((peers >> lane) <= 1u) mycnt[d] = before + __popc(peers);     // highest lane of the peer set
        __syncwarp();
        rank[i] = (uint16_t)(before + __popc(peers & kb_lanemask_lt()));
    }
    __syncthreads();

    // ---- per-digit: tile count (published at once), warp prefixes, local digit starts -----------
    uint32_t cnt = 0;
    uint32_t wc[WARPS];
    if (tid < KB_RADIX) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) { wc[w] = wcnt[w * KB_RADIX + tid]; cnt += wc[w]; }
        kb_st_relaxed(a.status + (size_t)tile * KB_RADIX + tid, (ST)cnt | (tile == 0 ? S::GLOBAL : S::LOCAL));
    }
    uint32_t lstart = 0;
    {
        uint32_t x = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x += y; }
        if (tid < KB_RADIX && lane == 31) wsum[warp] = x;
        __syncthreads();
        if (tid < KB_RADIX) {
            uint32_t add = 0;
            for (uint32_t w = 0; w < warp; w++) add += wsum[w];
            lstart = add + x - cnt;
            uint32_t run = lstart;                  // wcnt[w][d] := local start of (warp w, digit d)
#pragma unroll
            for (int w = 0; w < WARPS; w++) { wcnt[w * KB_RADIX + tid] = run; run += wc[w]; }
        }
    }
    __syncthreads();

    // ---- stage in ranked order --------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t d = kb_digit(key[i], a.shift, a.shard_n);
        skeys[mycnt[d] + rank[i]] = key[i];
    }

    // ---- decoupled look-back (after the staging, so predecessors had time to publish); the status
    //      words of 4 predecessor tiles are fetched per round trip -------------------------------------
    if (tid < KB_RADIX) {
        unsigned long long excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            bool done = false;
            while (!done) {
                ST v[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    v[j] = (t - j >= 0) ? kb_ld_relaxed(a.status + (size_t)(t - j) * KB_RADIX + tid) : (ST)S::GLOBAL;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (!done) {
                        while ((v[j] & S::FLAGS) == 0) v[j] = kb_ld_relaxed(a.status + (size_t)(t - j) * KB_RADIX + tid);
                        excl += (unsigned long long)(v[j] & S::VALUE);
                        if (v[j] & S::GLOBAL) done = true;
                    }
                }
                t -= 4;
            }
            kb_st_relaxed(a.status + (size_t)tile * KB_RADIX + tid, (ST)(excl + cnt) | S::GLOBAL);
        }
        dbase[tid] = __ldg(a.base + tid) + excl - (unsigned long long)lstart;
    }
    __syncthreads();

    // ---- coalesced store: each digit's run is contiguous in shared memory and in the output -----------
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t pos = i * THREADS + tid;
        if (pos < n_tile) {
            const uint64_t kv = skeys[pos];
            const uint32_t d = kb_digit(kv, a.shift, a.shard_n);
            a.out[dbase[d] + pos] = kv;
        }
    }
}

template <int THREADS, int ITEMS>
static inline size_t kb_onesweep_smem() {
    return (size_t)THREADS * ITEMS * 8 + KB_RADIX * 8 + (size_t)(THREADS / 32) * KB_RADIX * 4 + 8 * 4 + 16;
}

// ---- sorted tables of multi-word records (k > 28): LSD over 32-bit chunks of the record's base bits -----------------------
// The element is [chunk : 32][record index : 32]; one round = re-key every element with the next more significant chunk of its
// record (kb_chunk_key_kernel), then a stable 4-pass radix sort on the chunk (the onesweep above).  After the most significant
// chunk the indices are in the reference's LC_ALL=C order (left, right, middle: kstream.py:83-119); kb_table_gather_kernel
// writes the records out in that order.
struct KbChunkKeyArgs {
    uint64_t* ent;
    uint64_t n;
    const uint64_t* recs;
    uint32_t W, bit_pos, nbits;      // chunk = bits [bit_pos, bit_pos + nbits) of the record, nbits <= 32, inside one or two words
};

__global__ void __launch_bounds__(256) kb_chunk_key_kernel(const KbChunkKeyArgs a) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * 256) {
        const uint64_t idx = a.ent[i] & 0xFFFFFFFFULL;
        const uint64_t* r = a.recs + idx * a.W;
        const uint32_t w = a.bit_pos >> 6, o = a.bit_pos & 63;
        uint64_t v = r[w] << o;
        if (o && w + 1 < a.W) v |= r[w + 1] >> (64 - o);
        const uint64_t chunk = (v >> (64 - a.nbits)) << (32 - a.nbits);          // left-aligned in 32 bits
        a.ent[i] = (chunk << 32) | idx;
    }
}

struct KbTableGatherArgs {
    const uint64_t* ent;
    uint64_t n;
    const uint64_t* recs;
    uint32_t W;
    uint64_t* out;
};

__global__ void __launch_bounds__(256) kb_table_gather_kernel(const KbTableGatherArgs a) {
    const uint64_t total = a.n * a.W;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const uint64_t i = t / a.W, j = t % a.W;
        a.out[t] = a.recs[(a.ent[i] & 0xFFFFFFFFULL) * a.W + j];
    }
}

// kstream tables without --complements / with --canonicals: K1 wrote every record twice (KbExtractArgs.strand_mode), the sort left
// the copies next to each other: keep one of each pair.
__global__ void __launch_bounds__(256) kb_every_second_kernel(const uint64_t* in, uint64_t* out, uint64_t n_out) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * 256) out[i] = in[2 * i];
}
