// kb_extract.cuh — K1: k-mer extraction, 2-bit packing, validity masking, both strands.
//
// Replaces the reference's per-record Python generator chain (kstream/kstream.py):
//   _kmers :617-642, _mapsoft :751-766 | _omitsoft :734-749, _complements :644-677,
//   _disallow("Nn") :715-732, _split([L,-R]) :805-832, and the text k-mer file of write :250-325.
//
// Input  : all sequences of this rank's files as ASCII bytes in one buffer, at least one separator
//          byte (any non-ACGT byte) between FASTA records and between files; `file_starts[f]` is the
//          first byte of local file f.  The buffer is padded with separators past n_bases.
// Output : packed records (kb_common.cuh), dense, in no particular order:
//          DIRECT   out_entries[i] = record word
//          INDIRECT out_recs[i*WN .. ] = record words, out_entries[i] = hash32(flank)<<32 | i
//          INDIRECT, lazy: no records; out_entries[i] = hash31(flank)<<33 | strand<<32 | window start (kb_prefilter.cuh)
//
// One CTA processes tiles of KB_K1_TB window-start positions (persistent, grid-strided).  Per tile:
//   1. the tile's bases (+ k-1 halo) are packed once into shared memory: forward 2-bit stream,
//      reverse-complement stream, and a "bad base" bitmap (N, IUPAC, separators, lowercase when
//      --omit-soft);
//   2. a window-validity bitmap is derived from the bad bitmap (window valid <=> no bad base in it);
//      its popcounts give every warp its dense output offset, the tile claims its output range with
//      ONE global atomic;
//   3. lane = one window start; consecutive lanes write consecutive (forward, reverse) record
//      pairs, so every warp store is a fully coalesced run of 16-byte (DIRECT) or 2*8*WN-byte pieces.
#pragma once
#include "kb_common.cuh"

#define KB_K1_THREADS 256
#define KB_K1_C 16                                   // window starts per thread
#define KB_K1_TB (KB_K1_THREADS * KB_K1_C)           // 4096 window starts per tile
#define KB_K1_MAXHALO 256                            // k - 1 <= 256 (k <= 252 by KB_MAX_W)
#define KB_K1_PAD (KB_K1_TB + KB_K1_MAXHALO + 64)    // separator bytes required past the last tile start

struct KbExtractArgs {
    const uint8_t* bases;
    uint64_t n_bases;            // bytes in use (windows starting at >= n_bases are never valid)
    const uint64_t* file_starts; // n_local_files + 1 entries (last = n_bases)
    const uint32_t* file_gid;    // local file index -> global file id
    int n_local_files;
    int soft_omit;               // 1 = --omit-soft: lowercase letters are "bad"
    KbLayout lo;
    uint64_t* out_entries;
    uint64_t* out_recs;          // INDIRECT only
    int lazy;                    // INDIRECT only: 1 = no records, element = [flank hash 31][strand 1][window start 32] (kb_prefilter.cuh)
    unsigned long long* n_out;   // global record counter (claimed per tile)
    uint32_t tile0, n_tiles;     // tiles [tile0, tile0 + n_tiles)
    uint64_t pos_lo, pos_hi;     // only windows starting in [pos_lo, pos_hi) are emitted (one-file tables)
    unsigned long long* hist;    // != null: histogram of digit (element >> hist_shift) & (2^hist_bits - 1), hist_bits <= 9
    uint32_t hist_shift, hist_bits;   // (first partition level of kb_part.cuh, fused here to save a read of the elements)
    int strand_mode;             // one-word records of the kstream table path: 0 = every window and its reverse complement (_complements,
                                 // kstream.py:644-677), 1 = the window only (no --complements), 2 = the alphabetically first of the two
                                 // (_canonicals :679-694).  Modes 1 / 2 still write TWO records per window — the same one twice —
                                 // so that the tile bookkeeping stays put; the table path keeps every second record after the sort
};

// 4 ASCII bytes (little-endian in x) -> 8 bits of 2-bit codes (first base in the top bits) and 4 "bad" bits
// (bit i = byte i).  A/C/G/T -> 0..3 via ((u>>1)^(u>>2))&3 on the upper-cased byte; a byte is good iff the letter that
// its code stands for ("ACGT"[code], one PRMT for all four bytes) is the byte itself.
__device__ __forceinline__ uint32_t kb_pack4(uint32_t x, int soft_omit, uint32_t& bad4) {
    const uint32_t u = x & 0xDFDFDFDFu;
    const uint32_t c = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
    const uint32_t packed = (c * 0x40100401u) >> 24;     // b0<<6 | b1<<4 | b2<<2 | b3 (no carries)
    const uint32_t t = (c | (c >> 4)) & 0x00FF00FFu;
    const uint32_t sel = (t | (t >> 8)) & 0xFFFFu;       // one selector nibble per byte = its code
    const uint32_t d = __byte_perm(0x54474341u, 0u, sel) ^ u;                    // zero byte <=> A, C, G or T (either case)
    uint32_t m = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;          // bit 7 of every non-zero byte
    if (soft_omit) m |= (x & 0x20202020u) << 2;                                  // --omit-soft: lower case is bad too
    bad4 = (m * 0x00204081u) >> 28;                                              // bits 7, 15, 23, 31 -> 0..3
    return packed;
}

// shared memory of one 256-thread group (bytes, multiple of 16)
__host__ __device__ __forceinline__ size_t kb_extract_smem_dev(int k) {
    const uint32_t halo = ((uint32_t)(k - 1) + 31u) & ~31u;
    const uint32_t NWORD = (KB_K1_TB + halo) / 32;
    const size_t s = (size_t)(NWORD + 1) * 8 * 2 + (size_t)((NWORD + 3) & ~1u) * 4 + (size_t)(KB_K1_TB / 32) * 4 * 2;
    return (s + 31) & ~(size_t)15;
}

// WN = 1: DIRECT; 2/4/8: INDIRECT with WN words per record.
// GROUPS = 1: one 256-thread group per CTA, digit histogram of <= 9 bits (512 shared-memory counters).
// GROUPS = 4: FOUR independent 256-thread groups per CTA (named barriers), ONE CTA per SM, sharing a shared-memory histogram
//             of up to 16 bits with 16-bit packed counters (128 KB): the child counts of the first TWO partition levels come out
//             of K1 and the partition never re-reads the records just to count them.
template <int GROUPS> __device__ __forceinline__ void kb_k1_sync(uint32_t group) {
    if constexpr (GROUPS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(KB_K1_THREADS) : "memory");
}

template <int WN, int GROUPS>
__global__ void __launch_bounds__(KB_K1_THREADS * GROUPS) kb_extract_kernel(const KbExtractArgs a) {
    constexpr bool DIRECT = (WN == 1);
    extern __shared__ __align__(16) unsigned char kb_smem_all[];
    const uint32_t group = GROUPS == 1 ? 0u : threadIdx.x / KB_K1_THREADS;
    unsigned char* kb_smem_raw = kb_smem_all + (size_t)group * kb_extract_smem_dev(a.lo.k);
    const KbLayout& lo = a.lo;
    const uint32_t k = (uint32_t)lo.k;
    const uint32_t halo = ((k - 1) + 31u) & ~31u;
    const uint32_t NB = KB_K1_TB + halo;            // bases staged per tile (multiple of 32)
    const uint32_t NWORD = NB / 32;
    constexpr uint32_t NOK = KB_K1_TB / 32;         // 128 validity words

    uint64_t* fwd = reinterpret_cast<uint64_t*>(kb_smem_raw);           // NWORD + 1
    uint64_t* rcs = fwd + NWORD + 1;                                     // NWORD + 1
    uint32_t* bad = reinterpret_cast<uint32_t*>(rcs + NWORD + 1);        // NWORD + 2
    uint32_t* okw = bad + ((NWORD + 3) & ~1u);                           // NOK
    uint32_t* pre = okw + NOK;                                           // NOK exclusive popcount prefix
    __shared__ unsigned long long s_base_g[GROUPS];
    __shared__ int s_flo_g[GROUPS], s_fhi_g[GROUPS];
    __shared__ uint32_t s_hist[GROUPS == 1 ? 512 : 1];
    unsigned long long& s_base = s_base_g[group];
    int& s_flo = s_flo_g[group];
    int& s_fhi = s_fhi_g[group];
    uint32_t* bh = reinterpret_cast<uint32_t*>(kb_smem_all + (size_t)GROUPS * kb_extract_smem_dev(a.lo.k));   // GROUPS > 1: 2^(hist_bits-1) packed words

    const uint32_t tid = GROUPS == 1 ? threadIdx.x : threadIdx.x % KB_K1_THREADS, lane = tid & 31, warp = tid >> 5;
    const uint32_t FB = (uint32_t)lo.FB;
    const bool do_hist = a.hist != nullptr;
    const uint32_t hmask = (1u << a.hist_bits) - 1u;
    if (do_hist) {
        if constexpr (GROUPS == 1) { for (uint32_t i = tid; i < 512; i += KB_K1_THREADS) s_hist[i] = 0; }
        else { for (uint32_t i = threadIdx.x; i < (1u << (a.hist_bits - 1)); i += KB_K1_THREADS * GROUPS) bh[i] = 0; __syncthreads(); }
    }
    auto hist_add = [&](uint64_t e) {
        const uint32_t bin = (uint32_t)(e >> a.hist_shift) & hmask;
        if constexpr (GROUPS == 1) atomicAdd(&s_hist[bin], 1u);
        else {
            const uint32_t sh = (bin & 1u) << 4;
            const uint32_t old = atomicAdd(&bh[bin >> 1], 1u << sh);
            if (((old >> sh) & 0xFFFFu) == 0x7FFFu) {            // half-way to the carry: move 2^15 counts to the global histogram
                atomicSub(&bh[bin >> 1], 0x8000u << sh);
                atomicAdd(a.hist + bin, 0x8000ULL);
            }
        }
    };

    for (uint32_t tile = a.tile0 + blockIdx.x * GROUPS + group; tile < a.tile0 + a.n_tiles; tile += gridDim.x * GROUPS) {
        const uint64_t tile_base = (uint64_t)tile * KB_K1_TB;
        kb_k1_sync<GROUPS>(group);   // previous tile's readers are done with shared memory

        // ---- 1. pack: 32 bases per thread-iteration ------------------------------------------
        for (uint32_t wv = tid; wv < NWORD; wv += KB_K1_THREADS) {
            const uint8_t* src = a.bases + tile_base + (uint64_t)wv * 32;
            const uint4 v0 = kb_ld_stream128(src), v1 = kb_ld_stream128(src + 16);
            const uint32_t wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint64_t f = 0; uint32_t bd = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                uint32_t b4;
                f = (f << 8) | kb_pack4(wd[t], a.soft_omit, b4);
                bd |= b4 << (4 * t);
            }
            fwd[wv] = f;
            bad[wv] = bd;                            // bit i = base 32*wv + i
            if (!DIRECT) rcs[NWORD - 1 - wv] = kb_rc64(f);
        }
        if (tid == 0) {
            fwd[NWORD] = 0; rcs[NWORD] = 0; bad[NWORD] = 0xFFFFFFFFu; bad[NWORD + 1] = 0xFFFFFFFFu;
            // files of the first and last window start of this tile
            const uint64_t g0 = tile_base, g1 = tile_base + KB_K1_TB - 1;
            int l0 = 0, h0 = a.n_local_files;
            while (h0 - l0 > 1) { int m = (l0 + h0) >> 1; if (__ldg(a.file_starts + m) <= g0) l0 = m; else h0 = m; }
            int l1 = l0, h1 = a.n_local_files;
            while (h1 - l1 > 1) { int m = (l1 + h1) >> 1; if (__ldg(a.file_starts + m) <= g1) l1 = m; else h1 = m; }
            s_flo = l0; s_fhi = l1;
        }
        kb_k1_sync<GROUPS>(group);

        // ---- 2. window validity bitmap + dense offsets -----------------------------------------
        if (tid < NOK) {
            uint32_t acc = 0;
            uint32_t w = tid, lo_w = bad[w], hi_w = bad[w + 1];
            for (uint32_t d = 0; d < k; d++) {       // OR of the bad bitmap shifted by 0..k-1
                const uint32_t o = d & 31;
                if (o == 0 && d) { w++; lo_w = hi_w; hi_w = bad[w + 1 <= NWORD + 1 ? w + 1 : NWORD + 1]; }
                acc |= __funnelshift_r(lo_w, hi_w, o);
            }
            uint32_t ok = ~acc;
            const uint64_t g = tile_base + 32ull * tid;              // restrict to [pos_lo, pos_hi)
            if (g < a.pos_lo) ok &= (a.pos_lo - g >= 32) ? 0u : (0xFFFFFFFFu << (uint32_t)(a.pos_lo - g));
            if (g + 32 > a.pos_hi) ok &= (g >= a.pos_hi) ? 0u : (0xFFFFFFFFu >> (uint32_t)(g + 32 - a.pos_hi));
            okw[tid] = ok;
        }
        kb_k1_sync<GROUPS>(group);
        if (warp == 0) {
            uint32_t c[4], s = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) { c[i] = __popc(okw[4 * lane + i]); s += c[i]; }
            uint32_t inc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += y; }
            uint32_t ex = inc - s;
#pragma unroll
            for (int i = 0; i < 4; i++) { pre[4 * lane + i] = ex; ex += c[i]; }
            if (lane == 31) s_base = inc ? atomicAdd(a.n_out, 2ULL * inc) : 0ULL;
        }
        kb_k1_sync<GROUPS>(group);
        const uint64_t obase = s_base;
        const int flo = s_flo, fhi = s_fhi;
        const uint32_t gid_uniform = __ldg(a.file_gid + flo);

        // ---- 3. records ------------------------------------------------------------------------
#pragma unroll 2
        for (uint32_t c = 0; c < KB_K1_C; c++) {
            const uint32_t wi = c * (KB_K1_THREADS / 32) + warp;
            const uint32_t okbits = okw[wi];
            if (!((okbits >> lane) & 1u)) continue;
            const uint32_t p = wi * 32 + lane;                      // tile-relative window start
            const uint64_t oidx = obase + 2ULL * (pre[wi] + __popc(okbits & kb_lanemask_lt()));
            uint32_t gid = gid_uniform;
            if (flo != fhi) {                                        // tile spans several files (rare)
                const uint64_t gp = tile_base + p;
                int l0 = flo, h0 = fhi + 1;
                while (h0 - l0 > 1) { int m = (l0 + h0) >> 1; if (__ldg(a.file_starts + m) <= gp) l0 = m; else h0 = m; }
                gid = __ldg(a.file_gid + l0);
            }
            if constexpr (DIRECT) {
                const uint32_t D2 = 2 * lo.D, R2 = 2 * lo.R, K2 = 2 * k;
                const uint64_t mD = kb_lowmask(D2), mR = kb_lowmask(R2), mK = kb_lowmask(K2);
                uint64_t win = kb_get_bits(fwd, 2 * p, K2);
                uint64_t rcw = kb_rc64(win << (64 - K2)) & mK;
                if (a.strand_mode == 1) rcw = win;
                else if (a.strand_mode == 2) { win = min(win, rcw); rcw = win; }   // (2-bit codes A < C < G < T: numeric = alphabetical order)
                uint64_t r[2];
#pragma unroll
                for (int st = 0; st < 2; st++) {
                    const uint64_t x = st ? rcw : win;
                    uint64_t key = ((x >> (D2 + R2)) << R2) | (x & mR);
                    const uint64_t mid = (x >> R2) & mD;
                    uint64_t v = (uint64_t)gid;
                    if (FB) {
                        if (lo.mix) key = kb_mix(key, FB, lo.shs);
                        v |= key << (64 - FB);
                    }
                    if (D2) v |= mid << (64 - FB - D2);
                    r[st] = v;
                }
                kb_st_stream128(a.out_entries + oidx, r[0], r[1]);
                if (do_hist) {
                    hist_add(r[0]);
                    hist_add(r[1]);
                }
            } else {
                uint64_t ent[2];
#pragma unroll
                for (int st = 0; st < 2; st++) {
                    const uint64_t* s = st ? rcs : fwd;
                    const uint32_t o = st ? (NB - (p + k)) : p;      // first base of the window in that stream
                    if (a.lazy) {                                     // the record is built later, and only if it can still matter
                        const uint64_t h = kb_flank_hash([&](uint32_t pos, uint32_t n) { return kb_get_bits(s, pos, n); },
                                                         2 * o, 2 * (o + lo.L + lo.D), 2 * lo.L, 2 * lo.R);
                        ent[st] = (h & 0xFFFFFFFE00000000ULL) | ((uint64_t)st << 32) | (uint64_t)(uint32_t)(tile_base + p);
                        continue;
                    }
                    uint64_t rec[WN];
#pragma unroll
                    for (int j = 0; j < WN; j++) rec[j] = 0;
                    if (lo.L) kb_copy_bits<WN>(rec, 0, s, 2 * o, 2 * lo.L);
                    if (lo.R) kb_copy_bits<WN>(rec, 2 * lo.L, s, 2 * (o + lo.L + lo.D), 2 * lo.R);
                    uint64_t h = 0x9e3779b97f4a7c15ULL;               // hash of the flank words (mid not yet in)
#pragma unroll
                    for (int j = 0; j < WN; j++) if (j < lo.FW) h = kb_mix64(h ^ rec[j]);
                    if (lo.D) kb_copy_bits<WN>(rec, FB, s, 2 * (o + lo.L), 2 * lo.D);
                    rec[WN - 1] |= (uint64_t)gid;
                    uint64_t* dst = a.out_recs + (oidx + st) * WN;
#pragma unroll
                    for (int j = 0; j < WN; j += 2) kb_st_stream128(dst + j, rec[j], rec[j + 1]);
                    ent[st] = (h & 0xFFFFFFFF00000000ULL) | (uint64_t)(uint32_t)(oidx + st);
                }
                kb_st_stream128(a.out_entries + oidx, ent[0], ent[1]);
                if (do_hist) {
                    hist_add(ent[0]);
                    hist_add(ent[1]);
                }
            }
        }
    }
    if (do_hist) {
        __syncthreads();
        if constexpr (GROUPS == 1) {
            for (uint32_t i = tid; i <= hmask; i += KB_K1_THREADS) {
                const uint32_t c = s_hist[i];
                if (c) atomicAdd(a.hist + i, (unsigned long long)c);
            }
        } else {
            for (uint32_t w = threadIdx.x; w < (1u << (a.hist_bits - 1)); w += KB_K1_THREADS * GROUPS) {
                const uint32_t v = bh[w];
                if (v & 0xFFFFu) atomicAdd(a.hist + 2 * w, (unsigned long long)(v & 0xFFFFu));
                if (v >> 16) atomicAdd(a.hist + 2 * w + 1, (unsigned long long)(v >> 16));
            }
        }
    }
}

static inline size_t kb_extract_smem(int k, int groups = 1, int hist_bits = 0) {
    return (size_t)groups * kb_extract_smem_dev(k) + (groups > 1 ? ((size_t)1 << (hist_bits - 1)) * 4 : 0) + 16;
}
