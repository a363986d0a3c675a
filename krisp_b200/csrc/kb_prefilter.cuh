// kb_prefilter.cuh — K3 for multi-word records (k > 28, e.g. 32/60/32 primer mode): filter by flank hash, then verify.
//
// Same contract as kb_hash.cuh (reference: intersectSortedStreams shared.py:321-347 folded by mergeFiles
// intersectAmplicons.py:232-310 = rule S6, then filterAlignments.py:4-28 = rule S7), split in three steps so that the wide
// records are only ever built for the few occurrences that can still matter:
//
//   K1 (kb_extract.cuh, `lazy`)  writes NO records, only the 8-byte element [ flank hash : 31 ][ strand : 1 ][ window start : 32 ];
//   K2 (kb_part.cuh)             partitions the elements by the top hash bits, as always;
//   K3a kb_prefilter_kernel      per bucket: file-presence bitmaps in a direct-mapped shared-memory table indexed by the hash bits
//                                below the bucket bits.  Several keys may share a table entry, which can only ADD files: an entry
//                                that misses a file proves that every key mapped to it misses that file (S6 fails), so its
//                                occurrences are dropped; the others are compacted (bucket by bucket, contiguous);
//   K3b kb_materialize_kernel    builds the W-word record of every kept occurrence from the sequence bytes (window start + strand)
//                                in compacted order, so that record i belongs to element i;
//   K3c kb_hash_kernel           (kb_hash.cuh) exact grouping by COMPLETE flank key, S6 / S7, survivors — on the compacted buckets,
//                                whose record reads are now sequential.
//
// Exactness never depends on the hash: K3a only discards what is provably dead, K3c compares whole keys.
#pragma once
#include "kb_hash.cuh"

#define KB_PF_THREADS 512
#define KB_PF_ITEMS 8
#define KB_PF_CAP (KB_PF_THREADS * KB_PF_ITEMS)      // buckets up to this size stay in registers between the three steps
#define KB_PF_BLOCKS 1024                            // coarse position -> file table

struct KbPrefilterArgs {
    const uint64_t* ent;                 // partitioned lazy elements
    const unsigned long long* bstart;    // [n_buckets + 1]
    uint32_t n_buckets, bb, tb;          // bucket bits, table bits (2^tb entries of PWN presence words)
    const uint64_t* file_starts;         // [n_local_files + 1] positions in the sequence buffer (< 2^32)
    const uint32_t* file_gid;            // local file -> global file id
    int n_local_files;
    uint32_t blk_shift;                  // (last position >> blk_shift) < KB_PF_BLOCKS
    uint32_t full[8];
    uint64_t* out;                       // kept elements, bucket by bucket
    unsigned long long* n_out;           // cursor into `out` (one atomic per bucket)
    unsigned long long* brun;            // [n_buckets][2] start, length of the bucket's kept elements in `out`
};

// local file that holds sequence position `pos` (binary search over the shared-memory copy of the file table)
__device__ __forceinline__ uint32_t kb_pf_file(const uint32_t* s_fs, int nf, uint32_t pos) {
    int l = 0, h = nf;
    while (h - l > 1) { const int m = (l + h) >> 1; if (s_fs[m] <= pos) l = m; else h = m; }
    return (uint32_t)l;
}

// PWN = presence words per table entry (2, 4 or 8: up to 64 / 128 / 256 files)
template <int PWN>
__global__ void __launch_bounds__(KB_PF_THREADS) kb_prefilter_kernel(const KbPrefilterArgs a) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    uint32_t* tab = reinterpret_cast<uint32_t*>(kb_smem_raw);
    __shared__ uint32_t s_fs[KB_MAX_FILES + 1];
    __shared__ uint8_t s_gid[KB_MAX_FILES];
    // coarse block of positions -> { first file start after the block's first position, ids of the files before / after it, bit 16:
    // more than one file starts inside the block }: ONE random shared-memory read per record
    __shared__ uint2 s_blk[KB_PF_BLOCKS];
    __shared__ uint32_t s_count, s_cursor;
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const int nf = a.n_local_files;
    const uint32_t S = 1u << a.tb;
    for (int i = (int)tid; i <= nf; i += KB_PF_THREADS) s_fs[i] = (uint32_t)min((unsigned long long)a.file_starts[i], 0xFFFFFFFFull);
    for (int i = (int)tid; i < nf; i += KB_PF_THREADS) s_gid[i] = (uint8_t)a.file_gid[i];
    __syncthreads();
    for (uint32_t i = tid; i < KB_PF_BLOCKS; i += KB_PF_THREADS) {
        const uint32_t p0 = i << a.blk_shift, p1 = p0 + ((1u << a.blk_shift) - 1u);
        const uint32_t f = kb_pf_file(s_fs, nf, p0);                      // file of the block's first position
        const uint32_t next = s_fs[min(f + 1, (uint32_t)nf)];             // where the following file starts (or the data end)
        const uint32_t f1 = min(f + 1, (uint32_t)max(nf - 1, 0));
        const bool many = f + 2 < (uint32_t)nf && s_fs[f + 2] <= p1;      // a third file starts inside the block
        s_blk[i] = make_uint2(next, (uint32_t)s_gid[f] | ((uint32_t)s_gid[f1] << 8) | (many ? 0x10000u : 0u));
    }

    auto slot_of = [&](uint64_t e) -> uint32_t* { return tab + (size_t)((uint32_t)((e << a.bb) >> (64 - a.tb))) * PWN; };
    auto alive = [&](uint64_t e) -> bool {
        const uint32_t* p = slot_of(e);
        bool ok = true;
#pragma unroll
        for (int j = 0; j < PWN; j += 2) {
            const uint2 v = *reinterpret_cast<const uint2*>(p + j);
            ok = ok && v.x == a.full[j] && v.y == a.full[j + 1];
        }
        return ok;
    };
    auto mark = [&](uint64_t e) {
        const uint32_t pos = (uint32_t)e;
        const uint2 blk = s_blk[pos >> a.blk_shift];
        uint32_t id = pos < blk.x ? (blk.y & 0xFFu) : ((blk.y >> 8) & 0xFFu);
        if ((blk.y & 0x10000u) && pos >= blk.x) id = s_gid[kb_pf_file(s_fs, nf, pos)];   // several files start in this block (tiny files)
        atomicOr(slot_of(e) + (id >> 5), 1u << (id & 31));
    };

    for (uint32_t b = blockIdx.x; b < a.n_buckets; b += gridDim.x) {
        const uint64_t bs = a.bstart[b], be = a.bstart[b + 1];
        __syncthreads();                                   // the previous bucket's readers are done with the table
        if (be == bs) { if (tid == 0) { a.brun[2 * (size_t)b] = 0; a.brun[2 * (size_t)b + 1] = 0; } continue; }
        const uint32_t nb = (uint32_t)min(be - bs, (uint64_t)KB_PF_CAP + 1);
        const bool in_regs = nb <= KB_PF_CAP;
        uint64_t r[KB_PF_ITEMS];
        if (in_regs) {
#pragma unroll
            for (int u = 0; u < KB_PF_ITEMS; u++) { const uint32_t i = u * KB_PF_THREADS + tid; r[u] = i < nb ? kb_ld_stream(a.ent + bs + i) : 0ULL; }
        }
        {
            uint4* t4 = reinterpret_cast<uint4*>(tab);
            const uint32_t n4 = S * PWN / 4;
            for (uint32_t i = tid; i < n4; i += KB_PF_THREADS) t4[i] = make_uint4(0, 0, 0, 0);
            if (tid == 0) { s_count = 0; s_cursor = 0; }
        }
        __syncthreads();
        // ---- 1. presence bitmaps ----------------------------------------------------------------------------------------
        if (in_regs) {
#pragma unroll
            for (int u = 0; u < KB_PF_ITEMS; u++) if (u * KB_PF_THREADS + tid < nb) mark(r[u]);
        } else {
            for (uint64_t i = bs + tid; i < be; i += KB_PF_THREADS) mark(a.ent[i]);
        }
        __syncthreads();
        // ---- 2. count what stays -----------------------------------------------------------------------------------------
        uint32_t keep = 0, mine = 0;
        if (in_regs) {
#pragma unroll
            for (int u = 0; u < KB_PF_ITEMS; u++) if (u * KB_PF_THREADS + tid < nb && alive(r[u])) keep |= 1u << u;
            mine = __popc(keep);
        } else {
            for (uint64_t i = bs + tid; i < be; i += KB_PF_THREADS) mine += alive(a.ent[i]) ? 1u : 0u;
        }
        uint32_t inc = mine;                                // inclusive prefix over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += y; }
        uint32_t woff = 0;
        if (lane == 31 && inc) woff = atomicAdd(&s_count, inc);
        woff = __shfl_sync(0xFFFFFFFFu, woff, 31);
        __syncthreads();
        const uint32_t total = s_count;
        if (tid == 0) {
            const unsigned long long base = total ? atomicAdd(a.n_out, (unsigned long long)total) : 0ULL;
            s_base = base;
            a.brun[2 * (size_t)b] = base; a.brun[2 * (size_t)b + 1] = total;
        }
        if (total == 0) continue;                          // (uniform: every thread read the same s_count)
        __syncthreads();
        // ---- 3. compact ----------------------------------------------------------------------------------------------------
        const unsigned long long base = s_base;
        if (in_regs) {
            uint64_t* dst = a.out + base + woff + (inc - mine);
#pragma unroll
            for (int u = 0; u < KB_PF_ITEMS; u++) if ((keep >> u) & 1u) *dst++ = r[u];
        } else {
            // (the bucket is re-read from L2) one shared-memory atomic per warp and round, coalesced stores
            for (uint64_t i0 = bs; i0 < be; i0 += KB_PF_THREADS) {
                const uint64_t i = i0 + tid;
                uint64_t e = 0; bool k = false;
                if (i < be) { e = a.ent[i]; k = alive(e); }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
                uint32_t off = 0;
                if (lane == 0 && m) off = atomicAdd(&s_cursor, (uint32_t)__popc(m));
                off = __shfl_sync(0xFFFFFFFFu, off, 0);
                if (k) a.out[base + off + __popc(m & kb_lanemask_lt())] = e;
            }
        }
    }
}

static inline int kb_prefilter_pwn(int PW) { return PW <= 2 ? 2 : (PW <= 4 ? 4 : 8); }
static inline uint32_t kb_prefilter_tb(int PW, uint32_t bb) {
    uint32_t tb = 13;
    while (tb > 6 && ((size_t)1 << tb) * (size_t)kb_prefilter_pwn(PW) * 4 > 64 * 1024) tb--;
    return std::min<uint32_t>(tb, 31u - bb);
}
static inline size_t kb_prefilter_smem(int PW, uint32_t tb) { return ((size_t)1 << tb) * (size_t)kb_prefilter_pwn(PW) * 4 + 16; }

// ---- K3b: records of the kept occurrences, from the sequence bytes ------------------------------------------------------------
struct KbMatArgs {
    uint64_t* ent;               // in: lazy elements (compacted); out: [ flank hash : 32 ][ index : 32 ]
    uint64_t n;
    const uint8_t* bases;
    const uint64_t* file_starts;
    const uint32_t* file_gid;
    int n_local_files;
    KbLayout lo;
    uint64_t* recs;              // [n][W]
};

// `n` (1..64) bits of the MSB-first bit string s (SN words, in registers) starting at bit `pos`
template <int SN>
__device__ __forceinline__ uint64_t kb_mat_bits(const uint64_t (&s)[SN], uint32_t pos, uint32_t n) {
    const uint32_t w = pos >> 6, o = pos & 63;
    uint64_t x = 0, y = 0;
#pragma unroll
    for (int j = 0; j < SN; j++) {
        if (j == (int)w) x = s[j];
        if (j == (int)w + 1) y = s[j];
    }
    uint64_t v = x << o;
    if (o) v |= y >> (64 - o);
    return v >> (64 - n);
}
template <int WN, int SN>
__device__ __forceinline__ void kb_mat_copy(uint64_t (&rec)[WN], uint32_t dst, const uint64_t (&s)[SN], uint32_t src, uint32_t n) {
    while (n > 0) {
        const uint32_t c = n < 64 ? n : 64;
        kb_put_bits<WN>(rec, dst, c, kb_mat_bits<SN>(s, src, c));
        dst += c; src += c; n -= c;
    }
}

template <int WN>
__global__ void __launch_bounds__(256) kb_materialize_kernel(const KbMatArgs a) {
    __shared__ uint32_t s_fs[KB_MAX_FILES + 1];
    __shared__ uint32_t s_gid[KB_MAX_FILES];
    const KbLayout& lo = a.lo;
    const int nf = a.n_local_files;
    for (int i = (int)threadIdx.x; i <= nf; i += 256) s_fs[i] = (uint32_t)min((unsigned long long)a.file_starts[i], 0xFFFFFFFFull);
    for (int i = (int)threadIdx.x; i < nf; i += 256) s_gid[i] = a.file_gid[i];
    __syncthreads();
    const uint32_t k = (uint32_t)lo.k;
    constexpr int NV = 2 * WN + 1;          // 16-byte pieces that hold a window of <= 32 WN - 4 bases at any alignment
    constexpr int SN = WN + 2;              // 2-bit stream: 32 bases per word, + one readable pad word
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * 256) {
        const uint64_t e = a.ent[i];
        const uint32_t pos = (uint32_t)e;
        const uint32_t strand = (uint32_t)(e >> 32) & 1u;
        // the aligned 16-byte pieces around the window as a 2-bit stream (first base in the top bits of s[0]); the window
        // starts `off` bases into it
        const uint4* v16 = reinterpret_cast<const uint4*>(a.bases + (pos & ~15u));
        const uint32_t off = pos & 15u;
        const uint32_t nld = (off + k + 15u) >> 4;
        uint64_t s[SN];
#pragma unroll
        for (int j = 0; j < SN; j++) s[j] = 0;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            if (j < (int)nld) {
                const uint4 v = __ldg(v16 + j);
                const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
                uint32_t pk = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t u = wd[t] & 0xDFDFDFDFu;
                    const uint32_t c = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
                    pk = (pk << 8) | ((c * 0x40100401u) >> 24);
                }
                s[j >> 1] |= (uint64_t)pk << ((j & 1) ? 0 : 32);
            }
        }
        uint32_t o0 = off;                                        // first base of the occurrence in the stream
        if (strand) {                                             // reverse complement of the whole stream (32 (WN + 1) bases)
            uint64_t t[SN];
#pragma unroll
            for (int j = 0; j <= WN; j++) t[j] = kb_rc64(s[WN - j]);
            t[WN + 1] = 0;
#pragma unroll
            for (int j = 0; j < SN; j++) s[j] = t[j];
            o0 = 32u * (WN + 1) - (off + k);
        }
        uint64_t rec[WN];
#pragma unroll
        for (int j = 0; j < WN; j++) rec[j] = 0;
        if (lo.L) kb_mat_copy<WN, SN>(rec, 0, s, 2 * o0, 2 * lo.L);
        if (lo.R) kb_mat_copy<WN, SN>(rec, 2 * lo.L, s, 2 * (o0 + lo.L + lo.D), 2 * lo.R);
        const uint64_t h = kb_flank_hash([&](uint32_t p, uint32_t n) { return kb_mat_bits<SN>(s, p, n); },      // = K1's (lazy)
                                         2 * o0, 2 * (o0 + lo.L + lo.D), 2 * lo.L, 2 * lo.R);
        if (lo.D) kb_mat_copy<WN, SN>(rec, lo.FB, s, 2 * (o0 + lo.L), 2 * lo.D);
        rec[WN - 1] |= (uint64_t)s_gid[kb_pf_file(s_fs, nf, pos)];
        uint64_t* dst = a.recs + i * WN;
#pragma unroll
        for (int j = 0; j < WN; j += 2) kb_st_stream128(dst + j, rec[j], rec[j + 1]);
        a.ent[i] = (h & 0xFFFFFFFF00000000ULL) | (uint64_t)(uint32_t)i;
    }
}
