// kb_extract_sym.cuh — K1 fused with a STRAND-SYMMETRIC level 0 of the partition: window items instead of records.
//
// Same reference stages as kb_extract_part.cuh (kstream/kstream.py: _kmers :617-642, _mapsoft / _omitsoft :734-766,
// _complements :644-677, _disallow("Nn") :715-732, _split([L,-R]) :805-832; grouping in place of GNU sort :83-119).
//
// Every valid window yields TWO records (forward and reverse complement, kstream.py:661-677) with two different flank keys.
// Level 0 of the partition only has to be a function of the flank key — and there is one that both records of a window share:
// the window's CORE = the positions covered by the forward key AND by the mirrored (reverse-strand) key,
//     P = K ∩ mirror(K),  K = [0, L) ∪ [k - R, k),
// read in the canonical orientation min(core, rc(core)).  P is symmetric, so the core of rc(w) is rc(core(w)): both records of a
// window get the same level-0 digit, and P ⊆ K, so two records with the same flank key get it too.  Hence level 0 can move ONE
// 8-byte item per window — [2k bits of bases][file id] — instead of two records:
//   * K1 + level 0 (kb_extract_items_kernel) rank, stage and store half as many elements;
//   * level 1 (kb_part_expand_kernel) reads an item, forms both records and partitions them by the top bits of their own mixed
//     flank keys: records exist in HBM only from level 1 on;
//   * multi-GPU: the owner of a key range is chosen by the symmetric digit, so the exchange carries items — half the NVLink bytes.
// The bucket hash (kb_hash_warp.cuh) is unchanged: all records of a flank key still meet in one bucket.
#pragma once
#include "kb_extract_part.cuh"
#include "kb_part.cuh"

#define KB_XS_TPC 2                                   // K1 tiles (4096 window starts each) per CTA: 8192 items staged, 32-item runs at 256 digits
#define KB_XS_ITEMS (KB_XP_WPT * KB_XS_TPC)           // items per thread

struct KbXSymArgs {
    KbXPartArgs x;                       // as kb_extract_part_kernel; cursor / limit / out_elems count ITEMS
    uint64_t core_mask;                  // 2-bit-per-base mask of the core positions inside the 2k-bit window
};

// core mask of a layout: base i of the window occupies bits [2(k-1-i), 2(k-1-i)+1]
static inline uint64_t kb_core_mask(int L, int D, int R, int* core_bases) {
    const int k = L + D + R;
    uint64_t m = 0;
    int n = 0;
    for (int i = 0; i < k; i++) {
        const int j = k - 1 - i;                                     // mirrored position
        const bool in_i = i < L || i >= k - R, in_j = j < L || j >= k - R;
        if (in_i && in_j) { m |= 3ULL << (2 * (k - 1 - i)); n++; }
    }
    if (core_bases) *core_bases = n;
    return m;
}

static inline size_t kb_xsym_smem() {
    return (size_t)KB_XS_TPC * KB_XP_TB * 8 + KB_XP_MAXR * 8 + KB_XS_TPC * 144 * 8 + KB_XP_MAXR * 4 + 16 * 4 + KB_XS_TPC * 136 * 4 + 16;
}

__device__ __forceinline__ uint32_t kb_sym_digit(uint64_t win, uint64_t rcw, uint64_t core_mask, uint32_t dshift) {
    const uint64_t c0 = win & core_mask, c1 = rcw & core_mask;
    return (uint32_t)(((c0 < c1 ? c0 : c1) * KB_MIX_C1) >> dshift);
}

__global__ void __launch_bounds__(KB_XP_THREADS, 2) kb_extract_items_kernel(const KbXSymArgs as) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    const KbXPartArgs& a = as.x;
    uint64_t* skeys = reinterpret_cast<uint64_t*>(kb_smem_raw);                  // TPC * TB staged items
    uint64_t* dbase = skeys + KB_XS_TPC * KB_XP_TB;                              // MAXR
    uint64_t* fwd = dbase + KB_XP_MAXR;                                          // TPC x (NWORD + 1 <= 144)
    uint32_t* cnt = reinterpret_cast<uint32_t*>(fwd + KB_XS_TPC * 144);          // MAXR digit counters, then local starts
    uint32_t* wsum = cnt + KB_XP_MAXR;                                           // 16
    uint32_t* bad = wsum + 16;                                                   // TPC x (NWORD + 2 <= 136)
    __shared__ int s_flo[KB_XS_TPC], s_fhi[KB_XS_TPC];
    __shared__ uint32_t s_total;

    const KbLayout& lo = a.lo;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = (uint32_t)lo.k;
    const uint32_t halo = ((k - 1) + 31u) & ~31u;
    const uint32_t NB = KB_XP_TB + halo, NWORD = NB / 32;
    const uint32_t t_first = a.tile0 + blockIdx.x * KB_XS_TPC, t_end = a.tile0 + a.n_tiles;

    cnt[tid] = 0;
    // ---- 1. pack the tiles: 16 bases per thread, lane pairs assemble the 64-bit stream words ------------------------------
#pragma unroll
    for (int h = 0; h < KB_XS_TPC; h++) {
        const uint32_t tile = t_first + h;
        if (tile >= t_end) break;                                                // (uniform)
        const uint64_t tile_base = (uint64_t)tile * KB_XP_TB;
        uint64_t* fw = fwd + h * 144;
        uint32_t* bd_ = bad + h * 136;
        if (warp * 32 < NB / 16) {
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
                fw[tid >> 1] = ((uint64_t)f << 32) | f2;
                bd_[tid >> 1] = bd | (bd2 << 16);
            }
        }
        if (tid == KB_XP_THREADS - 1 - h) {
            fw[NWORD] = 0; bd_[NWORD] = 0xFFFFFFFFu; bd_[NWORD + 1] = 0xFFFFFFFFu;
            const uint64_t g0 = tile_base, g1 = tile_base + KB_XP_TB - 1;
            int l0 = 0, h0 = a.n_local_files;
            while (h0 - l0 > 1) { const int m = (l0 + h0) >> 1; if (__ldg(a.file_starts + m) <= g0) l0 = m; else h0 = m; }
            int l1 = l0, h1 = a.n_local_files;
            while (h1 - l1 > 1) { const int m = (l1 + h1) >> 1; if (__ldg(a.file_starts + m) <= g1) l1 = m; else h1 = m; }
            s_flo[h] = l0; s_fhi[h] = l1;
        }
    }
    __syncthreads();

    // ---- 2 + 3. validity, items and digit ranks of this thread's 8 windows per tile -------------------------------------------
    const uint32_t dshift = 64u - a.bits;
    const uint32_t K2 = 2 * k;
    const uint64_t mK = kb_lowmask((int)K2);
    uint64_t item[KB_XS_ITEMS];
    uint32_t rk[KB_XS_ITEMS / 2];                                                // ranks, two 16-bit halves per word
    uint32_t okm = 0;                                                            // validity of the 16 windows
#pragma unroll
    for (int i = 0; i < KB_XS_ITEMS; i++) item[i] = 0;
#pragma unroll
    for (int i = 0; i < KB_XS_ITEMS / 2; i++) rk[i] = 0;
#pragma unroll
    for (int h = 0; h < KB_XS_TPC; h++) {
        const uint32_t tile = t_first + h;
        if (tile >= t_end) break;
        const uint64_t tile_base = (uint64_t)tile * KB_XP_TB;
        const uint64_t* fw = fwd + h * 144;
        const uint32_t* bd_ = bad + h * 136;
        uint32_t ok8 = 0;
        {
            const uint64_t B = ((((uint64_t)bd_[(tid >> 2) + 1]) << 32) | bd_[tid >> 2]) >> (8u * (tid & 3u));
            const uint32_t mk = 0xFFFFFFFFu >> (32 - k);
#pragma unroll
            for (int j = 0; j < KB_XP_WPT; j++) ok8 |= (((uint32_t)(B >> j) & mk) == 0u ? 1u : 0u) << j;
            const uint64_t g = tile_base + (uint64_t)tid * KB_XP_WPT;
            if (g < a.pos_lo) ok8 &= (a.pos_lo - g >= 8) ? 0u : (0xFFu << (uint32_t)(a.pos_lo - g));
            if (g + 8 > a.pos_hi) ok8 &= (g >= a.pos_hi) ? 0u : (0xFFu >> (uint32_t)(g + 8 - a.pos_hi));
        }
        okm |= ok8 << (8 * h);
        if (!ok8) continue;
        const int flo = s_flo[h], fhi = s_fhi[h];
        uint64_t gids = 0x0101010101010101ULL * (uint64_t)__ldg(a.file_gid + flo);
        if (flo != fhi) {
#pragma unroll 1
            for (int j = 0; j < KB_XP_WPT; j++) {
                const uint64_t gp = tile_base + (uint64_t)tid * KB_XP_WPT + j;
                int l0 = flo, h0 = fhi + 1;
                while (h0 - l0 > 1) { const int m = (l0 + h0) >> 1; if (__ldg(a.file_starts + m) <= gp) l0 = m; else h0 = m; }
                gids = (gids & ~(0xFFULL << (8 * j))) | ((uint64_t)__ldg(a.file_gid + l0) << (8 * j));
            }
        }
        const uint64_t wa = fw[tid >> 2], wb = fw[(tid >> 2) + 1];
        const uint32_t s = 16u * (tid & 3u);
        const uint64_t x = s ? ((wa << s) | (wb >> (64 - s))) : wa;
        const uint64_t y = wb << s;
        const uint64_t rx = kb_rc64(y), ry = kb_rc64(x);
#pragma unroll
        for (int j = 0; j < KB_XP_WPT; j++) {
            if (!((ok8 >> j) & 1u)) continue;
            const uint64_t top = j ? ((x << (2 * j)) | (y >> (64 - 2 * j))) : x;
            const uint64_t win = top >> (64 - K2);
            const uint64_t rcw = (j ? ((ry >> (2 * j)) | (rx << (64 - 2 * j))) : ry) & mK;
            const uint32_t d = kb_sym_digit(win, rcw, as.core_mask, dshift);
            item[h * KB_XP_WPT + j] = (win << 8) | ((gids >> (8 * j)) & 0xFFu);
            const uint32_t r = atomicAdd(&cnt[d], 1u);
            rk[(h * KB_XP_WPT + j) >> 1] |= r << (16 * (j & 1));
        }
    }
    __syncthreads();

    // ---- 4. per digit: claim the slab range, local start --------------------------------------------------------------------
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
            if (lstart + c) atomicAdd(a.n_out, 2ULL * (unsigned long long)(lstart + c));     // (records = two per window)
        }
        dbase[tid] = (drop || !c) ? KB_XP_DROP : (a.out_elems[tid] + g - (unsigned long long)lstart);
    }
    __syncthreads();

    // ---- stage in digit order (the digit is recomputed from the item: window and its reverse complement) ----------------------
#pragma unroll
    for (int i = 0; i < KB_XS_ITEMS; i++) {
        if (!((okm >> i) & 1u)) continue;
        const uint64_t win = item[i] >> 8;
        const uint64_t rcw = kb_rc64(win << (64 - K2)) & mK;
        const uint32_t d = kb_sym_digit(win, rcw, as.core_mask, dshift);
        skeys[cnt[d] + ((rk[i >> 1] >> (16 * (i & 1))) & 0xFFFFu)] = item[i];
    }
    __syncthreads();

    // ---- coalesced stores ------------------------------------------------------------------------------------------------------
    const uint32_t total = s_total;
#pragma unroll
    for (int i = 0; i < KB_XS_ITEMS; i++) {
        const uint32_t pos = i * KB_XP_THREADS + tid;
        if (pos < total) {
            const uint64_t kv = skeys[pos];
            const uint64_t win = kv >> 8;
            const uint64_t rcw = kb_rc64(win << (64 - K2)) & mK;
            const unsigned long long db = dbase[kb_sym_digit(win, rcw, as.core_mask, dshift)];
            if (db != KB_XP_DROP) *reinterpret_cast<uint64_t*>((db + pos) << 3) = kv;
        }
    }
}

// ---- level 1 on items: expand every item into its two records, partition them by the top bits of their mixed flank keys ---------
// Same tile bookkeeping as kb_part_kernel<..., SLAB> (parents = level-0 slabs of items); an input tile of KB_PT_TILE items is
// processed in two halves, each yielding up to KB_PT_TILE records.
template <bool SPACER>
__global__ void __launch_bounds__(KB_PT_THREADS, 2) kb_part_expand_kernel(const KbPartArgs a, const KbLayout lo) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(kb_smem_raw);                   // TILE
    uint64_t* dbase = skeys + KB_PT_TILE;                                         // MAXR
    uint32_t* cnt = reinterpret_cast<uint32_t*>(dbase + KB_PT_MAXR);              // MAXR
    uint32_t* wsum = cnt + KB_PT_MAXR;                                            // MAXR / 32

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t parent, n_tile; uint64_t s;
    if (!kb_part_tile(a, blockIdx.x, parent, s, n_tile)) return;
    const uint32_t dmask = (1u << a.bits) - 1u;
    const uint32_t k = SPACER ? 28u : (uint32_t)lo.k, K2 = 2 * k;
    const uint32_t D2 = SPACER ? 2u : 2u * (uint32_t)lo.D, R2 = SPACER ? 4u : 2u * (uint32_t)lo.R, FB = SPACER ? 54u : (uint32_t)lo.FB;
    const uint64_t mD = kb_lowmask((int)D2), mR = kb_lowmask((int)R2), mK = kb_lowmask((int)K2);
    constexpr uint32_t HALF = KB_PT_TILE / 2, IPT = HALF / KB_PT_THREADS;         // 4096 items per half, 8 per thread

    // The items of a half are loaded one half ahead: the loads of half h + 1 are in flight while half h is scanned, staged and stored
    // (ncu of the unpipelined loop: 42 % of the stall cycles were long-scoreboard waits, issue slots 38 % busy).
    uint64_t it[IPT];
#pragma unroll
    for (int i = 0; i < (int)IPT; i++) {
        const uint32_t idx = i * KB_PT_THREADS + tid;
        it[i] = idx < min(HALF, n_tile) ? kb_ld_stream(a.in + s + idx) : 0ULL;
    }
    for (uint32_t base = 0; base < n_tile; base += HALF) {
        if (tid < KB_PT_MAXR) cnt[tid] = 0;
        __syncthreads();
        const uint32_t n_half = min(HALF, n_tile - base);
        uint64_t key[2 * IPT];
        uint32_t rank[IPT];
#pragma unroll
        for (int i = 0; i < (int)IPT; i++) {
            const uint32_t idx = i * KB_PT_THREADS + tid;
            const uint32_t gid = (uint32_t)it[i] & 0xFFu;
            const uint64_t win = it[i] >> 8;
            const uint64_t rcw = kb_rc64(win << (64 - K2)) & mK;
            uint32_t rr = 0;
#pragma unroll
            for (int st = 0; st < 2; st++) {
                const uint64_t w = st ? rcw : win;
                uint64_t fk = ((w >> (D2 + R2)) << R2) | (w & mR);
                const uint64_t mid = (w >> R2) & mD;
                uint64_t v = (uint64_t)gid;
                if (SPACER) {
                    v |= (fk * KB_MIX_C1) << 10;
                    v |= mid << 8;
                } else {
                    if (FB) {
                        if (lo.mix) fk = kb_mix(fk, (int)FB, lo.shs);
                        v |= fk << (64 - FB);
                    }
                    if (D2) v |= mid << (64 - FB - D2);
                }
                key[2 * i + st] = v;
                if (idx < n_half) rr |= atomicAdd(&cnt[(uint32_t)(v >> a.shift) & dmask], 1u) << (16 * st);
            }
            rank[i] = rr;
        }
        if (base + HALF < n_tile) {                                               // next half's items, consumed in the next iteration
            const uint32_t n_next = min(HALF, n_tile - base - HALF);
#pragma unroll
            for (int i = 0; i < (int)IPT; i++) {
                const uint32_t idx = i * KB_PT_THREADS + tid;
                it[i] = idx < n_next ? kb_ld_stream(a.in + s + base + HALF + idx) : 0ULL;
            }
        }
        __syncthreads();

        uint32_t c = 0, lstart = 0;
        unsigned long long g = 0;
        bool drop = false;
        if (tid < KB_PT_MAXR) {
            c = cnt[tid];
            if (c) {
                g = atomicAdd(a.cursor + (((size_t)parent << a.bits) | tid), (unsigned long long)c);
                if (g + c > ((((unsigned long long)parent << a.bits) | tid) + 1ULL) * a.ccap) { drop = true; *a.ovf = 1ULL; }
            }
        }
        {
            uint32_t x = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x += y; }
            if (tid < KB_PT_MAXR && lane == 31) wsum[warp] = x;
            __syncthreads();
            if (tid < KB_PT_MAXR) {
                uint32_t add = 0;
                for (uint32_t w = 0; w < warp; w++) add += wsum[w];
                lstart = add + x - c;
                cnt[tid] = lstart;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < (int)IPT; i++) {
            const uint32_t idx = i * KB_PT_THREADS + tid;
            if (idx < n_half) {
                skeys[cnt[(uint32_t)(key[2 * i] >> a.shift) & dmask] + (rank[i] & 0xFFFFu)] = key[2 * i];
                skeys[cnt[(uint32_t)(key[2 * i + 1] >> a.shift) & dmask] + (rank[i] >> 16)] = key[2 * i + 1];
            }
        }
        if (tid < KB_PT_MAXR) {
            const unsigned long long ob = (unsigned long long)(reinterpret_cast<uintptr_t>(a.out) >> 3);
            dbase[tid] = drop ? KB_PT_DROP : ob + g - (unsigned long long)lstart;
        }
        __syncthreads();
        const uint32_t n_rec = 2 * n_half;
#pragma unroll
        for (int i = 0; i < 2 * (int)IPT; i++) {
            const uint32_t pos = i * KB_PT_THREADS + tid;
            if (pos < n_rec) {
                const uint64_t kv = skeys[pos];
                const unsigned long long db = dbase[(uint32_t)(kv >> a.shift) & dmask];
                if (db != KB_PT_DROP) *reinterpret_cast<uint64_t*>((db + pos) << 3) = kv;
            }
        }
        __syncthreads();                                                          // (the next half reuses the counters and the staging area)
    }
}
