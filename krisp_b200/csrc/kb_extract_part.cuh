// kb_extract_part.cuh — K1 fused with level 0 of the radix partition ("slab" search path, one-word records).
//
// Same reference stages as kb_extract.cuh (kstream/kstream.py: _kmers :617-642, _mapsoft :751-766 | _omitsoft :734-749,
// _complements :644-677, _disallow("Nn") :715-732, _split([L,-R]) :805-832, the k-mer file of write :250-325) plus the
// first pass of the grouping that replaces GNU sort (kstream.py:83-119; see kb_part.cuh): the records of a tile never
// go to HBM in extraction order.  They are ranked by their level-0 digit (top bits of the mixed flank key) with one
// shared-memory atomic each, staged in shared memory in digit order, and every digit's run leaves the CTA as one
// contiguous store into that digit's SLAB: a fixed-capacity region whose fill level is a global cursor (one atomicAdd
// per tile and digit).  No histogram is needed before the records are written, so K1 and partition level 0 are ONE
// kernel: 16 B per base pair written instead of 16 written + 16 read + 16 written.
//
// Slab capacities come from the record bound (2 x bases) with a few percent of slack (the mixed key is uniform); a
// slab that overflows raises *ovf and the host repeats the search on the exact, histogram-based path (kb_extract.cuh +
// kb_part.cuh).  Multi-GPU: a digit's slab lives in the owner's receive buffer (out_elems[d] = a peer mapping), so the
// NVLink exchange is this kernel's store phase.
//
// Per tile of 4096 window starts (512 threads x 8 CONSECUTIVE windows — output order does not matter here, so the
// windows of a thread slide over one 128-bit register pair instead of being re-read from shared memory):
//   1. 16 bases per thread -> 2-bit forward stream + bad-base bitmap in shared memory;
//   2. every thread derives the validity of its 8 windows from 64 bits of the bad bitmap (no bad base in [p, p + k));
//   3. per window: forward record from the register pair, reverse-complement record from the pair's reverse
//      complement (computed once per thread), digit rank by shared-memory atomicAdd;
//   4. digit scan, slab claim, staging in digit order, coalesced stores.
#pragma once
#include "kb_common.cuh"
#include "kb_extract.cuh"

#define KB_XP_THREADS 512
#define KB_XP_WPT 8                                   // consecutive window starts per thread
#define KB_XP_TB (KB_XP_THREADS * KB_XP_WPT)          // 4096 = KB_K1_TB: both K1 variants share the tile grid / padding rules
#define KB_XP_MAXR 512                                // level-0 digits of at most 9 bits
#define KB_XP_DROP 0xFFFFFFFFFFFFFFFFULL
static_assert(KB_XP_TB == KB_K1_TB, "the fused kernel walks the same tiles as kb_extract_kernel");

struct KbXPartArgs {
    const uint8_t* bases;
    uint64_t n_bases;
    const uint64_t* file_starts;          // n_local_files + 1 entries
    const uint32_t* file_gid;
    int n_local_files;
    int soft_omit;
    KbLayout lo;                          // DIRECT layouts only (2k + 8 <= 64)
    uint32_t tile0, n_tiles;
    uint64_t pos_lo, pos_hi;              // only windows starting in [pos_lo, pos_hi)
    uint32_t bits;                        // level-0 digit = top `bits` bits of the record (1..9)
    unsigned long long* cursor;           // [2^bits] next free element of every digit's slab (advanced atomically)
    const unsigned long long* limit;      // [2^bits] end of every slab, same unit as cursor
    const unsigned long long* out_elems;  // [2^bits] element address (pointer / 8) cursor value 0 refers to: this GPU's buffer or a peer's
    unsigned long long* n_out;            // records produced
    unsigned long long* ovf;              // != 0 afterwards: some slab was too small (records were dropped)
};

static inline size_t kb_xpart_smem() {
    return (size_t)2 * KB_XP_TB * 8 + KB_XP_MAXR * 8 + 144 * 8 + KB_XP_MAXR * 4 + 16 * 4 + 136 * 4 + 16;
}

// SPACER: the layout is 25/1/2-like with mixing (FB = 54, D = 1, R = 2, k = 28): every shift is a compile-time constant.
template <bool SPACER>
__global__ void __launch_bounds__(KB_XP_THREADS, 2) kb_extract_part_kernel(const KbXPartArgs a) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(kb_smem_raw);                  // 2 * TB staged records
    uint64_t* dbase = skeys + 2 * KB_XP_TB;                                      // MAXR: element address of the digit's run - its local start
    uint64_t* fwd = dbase + KB_XP_MAXR;                                          // NWORD + 1 (<= 130)
    uint32_t* cnt = reinterpret_cast<uint32_t*>(fwd + 144);                      // MAXR digit counters, then local starts
    uint32_t* wsum = cnt + KB_XP_MAXR;                                           // 16
    uint32_t* bad = wsum + 16;                                                   // NWORD + 2 (<= 131)
    __shared__ int s_flo, s_fhi;
    __shared__ uint32_t s_total;

    const KbLayout& lo = a.lo;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = SPACER ? 28u : (uint32_t)lo.k;
    const uint32_t halo = ((k - 1) + 31u) & ~31u;                                // 0 (k = 1) or 32
    const uint32_t NB = KB_XP_TB + halo, NWORD = NB / 32;
    const uint32_t tile = a.tile0 + blockIdx.x;
    const uint64_t tile_base = (uint64_t)tile * KB_XP_TB;

    cnt[tid] = 0;
    // ---- 1. pack 16 bases per thread; lane pairs assemble the 64-bit stream words ------------------------------------
    if (warp * 32 < NB / 16) {                                                   // (warp-uniform: the shuffles need whole warps)
        uint32_t f = 0, bd = 0;
        if (tid < NB / 16) {
            const uint4 v = kb_ld_stream128(a.bases + tile_base + 16ull * tid);
            const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t b4;
                f = (f << 8) | kb_pack4(wd[t], a.soft_omit, b4);
                bd |= b4 << (4 * t);
            }
        }
        const uint32_t f2 = __shfl_down_sync(0xFFFFFFFFu, f, 1), bd2 = __shfl_down_sync(0xFFFFFFFFu, bd, 1);
        if (tid < NB / 16 && !(tid & 1u)) {
            fwd[tid >> 1] = ((uint64_t)f << 32) | f2;
            bad[tid >> 1] = bd | (bd2 << 16);                                    // bit i = base 32 * word + i
        }
    }
    if (tid == KB_XP_THREADS - 1) {
        fwd[NWORD] = 0; bad[NWORD] = 0xFFFFFFFFu; bad[NWORD + 1] = 0xFFFFFFFFu;
        const uint64_t g0 = tile_base, g1 = tile_base + KB_XP_TB - 1;            // files of the first and last window start
        int l0 = 0, h0 = a.n_local_files;
        while (h0 - l0 > 1) { const int m = (l0 + h0) >> 1; if (__ldg(a.file_starts + m) <= g0) l0 = m; else h0 = m; }
        int l1 = l0, h1 = a.n_local_files;
        while (h1 - l1 > 1) { const int m = (l1 + h1) >> 1; if (__ldg(a.file_starts + m) <= g1) l1 = m; else h1 = m; }
        s_flo = l0; s_fhi = l1;
    }
    __syncthreads();

    // ---- 2. validity of this thread's 8 windows: no bad base in [p, p + k) ------------------------------------------------
    uint32_t ok8 = 0;
    {
        const uint64_t B = ((((uint64_t)bad[(tid >> 2) + 1]) << 32) | bad[tid >> 2]) >> (8u * (tid & 3u));   // bit i = base 8 tid + i is bad
        const uint32_t mk = 0xFFFFFFFFu >> (32 - k);                             // (k <= 28 on this path)
#pragma unroll
        for (int j = 0; j < KB_XP_WPT; j++) ok8 |= (((uint32_t)(B >> j) & mk) == 0u ? 1u : 0u) << j;
        const uint64_t g = tile_base + (uint64_t)tid * KB_XP_WPT;                // restrict to [pos_lo, pos_hi)
        if (g < a.pos_lo) ok8 &= (a.pos_lo - g >= 8) ? 0u : (0xFFu << (uint32_t)(a.pos_lo - g));
        if (g + 8 > a.pos_hi) ok8 &= (g >= a.pos_hi) ? 0u : (0xFFu >> (uint32_t)(g + 8 - a.pos_hi));
    }

    // ---- 3. records of this thread's 8 windows + digit ranks -----------------------------------------------------------
    const uint32_t dshift = 64u - a.bits;
    uint64_t rec[2 * KB_XP_WPT];
    uint32_t rk[KB_XP_WPT];                                                      // ranks of (forward, reverse) as 16-bit halves
#pragma unroll
    for (int j = 0; j < KB_XP_WPT; j++) { rec[2 * j] = 0; rec[2 * j + 1] = 0; rk[j] = 0; }
    if (ok8) {
        const int flo = s_flo, fhi = s_fhi;
        uint64_t gids = 0x0101010101010101ULL * (uint64_t)__ldg(a.file_gid + flo);   // file id of every window (one byte each)
        if (flo != fhi) {                                                        // tile spans several files (rare)
#pragma unroll 1
            for (int j = 0; j < KB_XP_WPT; j++) {
                const uint64_t gp = tile_base + (uint64_t)tid * KB_XP_WPT + j;
                int l0 = flo, h0 = fhi + 1;
                while (h0 - l0 > 1) { const int m = (l0 + h0) >> 1; if (__ldg(a.file_starts + m) <= gp) l0 = m; else h0 = m; }
                gids = (gids & ~(0xFFULL << (8 * j))) | ((uint64_t)__ldg(a.file_gid + l0) << (8 * j));
            }
        }
        const uint32_t K2 = 2 * k;
        const uint32_t D2 = SPACER ? 2u : 2u * (uint32_t)lo.D, R2 = SPACER ? 4u : 2u * (uint32_t)lo.R, FB = SPACER ? 54u : (uint32_t)lo.FB;
        const uint64_t mD = kb_lowmask((int)D2), mR = kb_lowmask((int)R2), mK = kb_lowmask((int)K2);
        // the 64 bases starting at this thread's first window, as a 128-bit string (x = first 32 bases); only 8 + k - 1 <= 35 are used
        const uint64_t wa = fwd[tid >> 2], wb = fwd[(tid >> 2) + 1];
        const uint32_t s = 16u * (tid & 3u);
        const uint64_t x = s ? ((wa << s) | (wb >> (64 - s))) : wa;
        const uint64_t y = wb << s;
        // reverse complement of the 64-base string: base i <-> position 63 - i; window j = bits [2j, 2j + 2k) of (rx : ry)
        const uint64_t rx = kb_rc64(y), ry = kb_rc64(x);
#pragma unroll
        for (int j = 0; j < KB_XP_WPT; j++) {
            if (!((ok8 >> j) & 1u)) continue;
            const uint32_t gid = (uint32_t)(gids >> (8 * j)) & 0xFFu;
            const uint64_t top = j ? ((x << (2 * j)) | (y >> (64 - 2 * j))) : x;  // window j left-aligned
            const uint64_t win = top >> (64 - K2);
            const uint64_t rcw = (j ? ((ry >> (2 * j)) | (rx << (64 - 2 * j))) : ry) & mK;
            uint32_t rr = 0;
#pragma unroll
            for (int st = 0; st < 2; st++) {
                const uint64_t w = st ? rcw : win;
                uint64_t key = ((w >> (D2 + R2)) << R2) | (w & mR);
                const uint64_t mid = (w >> R2) & mD;
                uint64_t v = (uint64_t)gid;
                if (SPACER) {
                    v |= (key * KB_MIX_C1) << 10;                                // (the shift drops the bits above 2^54)
                    v |= mid << 8;
                } else {
                    if (FB) {
                        if (lo.mix) key = kb_mix(key, (int)FB, lo.shs);
                        v |= key << (64 - FB);
                    }
                    if (D2) v |= mid << (64 - FB - D2);
                }
                rec[2 * j + st] = v;
                const uint32_t r = atomicAdd(&cnt[(uint32_t)(v >> dshift)], 1u);
                rr |= r << (16 * st);
            }
            rk[j] = rr;
        }
    }
    __syncthreads();

    // ---- 4. per digit: claim the slab range, local start ----------------------------------------------------------------
    const uint32_t nd = 1u << a.bits;
    uint32_t c = 0, lstart = 0;
    unsigned long long g = 0;
    bool drop = false;
    if (tid < nd) {
        c = cnt[tid];
        if (c) {
            g = atomicAdd(a.cursor + tid, (unsigned long long)c);
            if (g + c > a.limit[tid]) { drop = true; *a.ovf = 1ULL; }
        }
    }
    {
        uint32_t xs = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t yv = __shfl_up_sync(0xFFFFFFFFu, xs, d); if (lane >= (uint32_t)d) xs += yv; }
        if (lane == 31) wsum[warp] = xs;
        __syncthreads();
        uint32_t add = 0;
        for (uint32_t w = 0; w < warp; w++) add += wsum[w];
        lstart = add + xs - c;
        cnt[tid] = lstart;
        if (tid == KB_XP_THREADS - 1) {
            s_total = lstart + c;
            if (lstart + c) atomicAdd(a.n_out, (unsigned long long)(lstart + c));
        }
        dbase[tid] = (drop || !c) ? KB_XP_DROP : (a.out_elems[tid] + g - (unsigned long long)lstart);
    }
    __syncthreads();

    // ---- stage in digit order ---------------------------------------------------------------------------------------------
    if (ok8) {
#pragma unroll
        for (int j = 0; j < KB_XP_WPT; j++) {
            if (!((ok8 >> j) & 1u)) continue;
            skeys[cnt[(uint32_t)(rec[2 * j] >> dshift)] + (rk[j] & 0xFFFFu)] = rec[2 * j];
            skeys[cnt[(uint32_t)(rec[2 * j + 1] >> dshift)] + (rk[j] >> 16)] = rec[2 * j + 1];
        }
    }
    __syncthreads();

    // ---- coalesced stores: every digit's run is contiguous in the staged tile and in its slab ------------------------------
    const uint32_t total = s_total;
#pragma unroll
    for (int i = 0; i < 2 * KB_XP_WPT; i++) {
        const uint32_t pos = i * KB_XP_THREADS + tid;
        if (pos < total) {
            const uint64_t kv = skeys[pos];
            const unsigned long long db = dbase[(uint32_t)(kv >> dshift)];
            if (db != KB_XP_DROP) *reinterpret_cast<uint64_t*>((db + pos) << 3) = kv;
        }
    }
}
