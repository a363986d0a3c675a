// kb_group_fast.cuh — K3 fast path: blocked reduce-by-key for one-word records.
//
// Same contract as kb_group.cuh (the reference stages it replaces are listed there), specialised for the
// headline shape: DIRECT records, <= 64 files, D <= 8 columns.  Instead of a warp per run, every thread
// owns 16 CONSECUTIVE sorted records and walks them serially with the group state in registers
// (file-presence bitmap 64 b, ingroup / outgroup one-hot column masks 32 b each, count); groups that
// span threads are stitched with one segmented warp-shuffle scan per tile, groups that leave the tile are
// finished by the owning tile reading ahead.  Cost per record drops from ~220 to ~40 instructions.
//
// Tile load: 128-bit coalesced global loads -> shared memory with an XOR swizzle on the 16-byte chunk
// index, so that both the coalesced store and the per-thread row read (stride 128 B) are conflict-free.
//
// Exactness under a prefix sort: if two adjacent records agree on the sort prefix but not on the full
// key, the run is "mixed" (records of one key need not be contiguous).  Every segment touching such a
// boundary is withheld, the boundary position is appended to a taint list, and kb_group_taint_kernel
// re-processes each mixed run exactly once with the ordered sweep of kb_group.cuh.
#pragma once
#include "kb_group.cuh"

#define KB_K3F_THREADS 256
#define KB_K3F_ITEMS 16
#define KB_K3F_TILE (KB_K3F_THREADS * KB_K3F_ITEMS)

struct KbFastArgs {
    KbGroupArgs g;
    uint64_t ingroup64, full64;
    unsigned long long* taint;      // positions of tainted boundaries
    unsigned long long* n_taint;
    uint64_t taint_cap;
};

struct KbFA {                        // aggregate of a (partial) group
    uint64_t pres;
    uint32_t in, out, cnt, start;    // start: tile-relative index of the group's first record
    uint32_t head, bad;              // head: the segment starts inside this item; bad: withheld
};

__device__ __forceinline__ KbFA kb_fa_combine(const KbFA& a, const KbFA& b) {   // b continues a
    KbFA r;
    r.pres = a.pres | b.pres; r.in = a.in | b.in; r.out = a.out | b.out; r.cnt = a.cnt + b.cnt;
    r.start = a.start; r.head = a.head; r.bad = a.bad | b.bad;
    return r;
}
__device__ __forceinline__ KbFA kb_fa_shfl_up(const KbFA& a, int d) {
    KbFA r;
    r.pres = __shfl_up_sync(0xFFFFFFFFu, a.pres, d);
    r.in = __shfl_up_sync(0xFFFFFFFFu, a.in, d); r.out = __shfl_up_sync(0xFFFFFFFFu, a.out, d);
    r.cnt = __shfl_up_sync(0xFFFFFFFFu, a.cnt, d); r.start = __shfl_up_sync(0xFFFFFFFFu, a.start, d);
    r.head = __shfl_up_sync(0xFFFFFFFFu, a.head, d); r.bad = __shfl_up_sync(0xFFFFFFFFu, a.bad, d);
    return r;
}
__device__ __forceinline__ KbFA kb_fa_shfl(const KbFA& a, int src) {
    KbFA r;
    r.pres = __shfl_sync(0xFFFFFFFFu, a.pres, src);
    r.in = __shfl_sync(0xFFFFFFFFu, a.in, src); r.out = __shfl_sync(0xFFFFFFFFu, a.out, src);
    r.cnt = __shfl_sync(0xFFFFFFFFu, a.cnt, src); r.start = __shfl_sync(0xFFFFFFFFu, a.start, src);
    r.head = __shfl_sync(0xFFFFFFFFu, a.head, src); r.bad = __shfl_sync(0xFFFFFFFFu, a.bad, src);
    return r;
}

// a closed group: S6 / S7 test and, for survivors, one row of the result table
__device__ __forceinline__ void kb_fa_finalize(const KbFastArgs& x, const KbFA& A, uint64_t key, uint64_t tile_start,
                                               uint32_t& n_closed, uint32_t& n_present) {
    const KbLayout& lo = x.g.lo;
    n_closed++;
    if (A.bad || A.pres != x.full64) return;
    n_present++;
    if (lo.D) {
        uint32_t y = A.in & A.out;
        y |= y >> 1; y |= y >> 2;
        if ((~y & 0x11111111u & (0xFFFFFFFFu << (4 * (8 - lo.D)))) == 0) return;
    }
    const unsigned long long slot = atomicAdd(x.g.n_res, 1ULL);
    if (slot >= x.g.cap) return;
    uint64_t kk = key;
    if (lo.mix) kk = kb_unmix(kk, lo.FB, lo.shs);
    x.g.res_flank[slot] = kk << (64 - lo.FB);
    if (lo.MW) { x.g.res_in[slot] = A.in; x.g.res_out[slot] = A.out; }
    x.g.res_size[slot] = A.cnt;
    x.g.res_run[2 * slot] = tile_start + A.start;
    x.g.res_run[2 * slot + 1] = A.cnt;
}

template <bool D1>
__global__ void __launch_bounds__(KB_K3F_THREADS) kb_group_fast_kernel(const KbFastArgs x) {
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    __shared__ uint4 s16[KB_K3F_TILE / 2];
    __shared__ KbFA wtot[KB_K3F_THREADS / 32];
    __shared__ uint32_t s_closed, s_present;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_start = (uint64_t)blockIdx.x * KB_K3F_TILE;
    if (tile_start >= a.n) return;
    const uint32_t n_tile = (uint32_t)min((uint64_t)KB_K3F_TILE, a.n - tile_start);
    if (tid == 0) { s_closed = 0; s_present = 0; }

    // ---- coalesced load, swizzled store ---------------------------------------------------------
#pragma unroll
    for (int i = 0; i < KB_K3F_ITEMS / 2; i++) {
        const uint32_t c = i * KB_K3F_THREADS + tid;            // 16-byte chunk of the tile
        uint4 v = make_uint4(0, 0, 0, 0);
        if (2 * c + 1 < n_tile) v = kb_ld_stream128(a.ent + tile_start + 2 * c);
        else if (2 * c < n_tile) { const uint64_t e = a.ent[tile_start + 2 * c]; v.x = (uint32_t)e; v.y = (uint32_t)(e >> 32); }
        const uint32_t row = c >> 3, pos = c & 7;
        s16[row * 8 + (pos ^ (row & 7))] = v;
    }
    __syncthreads();

    // ---- my 16 consecutive records + the neighbours on both sides --------------------------------
    uint64_t r[KB_K3F_ITEMS];
#pragma unroll
    for (int j = 0; j < KB_K3F_ITEMS / 2; j++) {
        const uint4 v = s16[tid * 8 + (j ^ (tid & 7))];
        r[2 * j] = (uint64_t)v.x | ((uint64_t)v.y << 32);
        r[2 * j + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
    }
    const uint32_t first = tid * KB_K3F_ITEMS;
    const uint32_t nv = first >= n_tile ? 0u : min((uint32_t)KB_K3F_ITEMS, n_tile - first);
    bool have_prev = false, have_next = false;
    uint64_t prev = 0, next = 0;
    if (nv) {
        if (tid > 0) {
            const uint4 v = s16[(tid - 1) * 8 + (7 ^ ((tid - 1) & 7))];
            prev = (uint64_t)v.z | ((uint64_t)v.w << 32); have_prev = true;
        } else if (tile_start > 0) { prev = a.ent[tile_start - 1]; have_prev = true; }
        if (nv == KB_K3F_ITEMS && first + KB_K3F_ITEMS < n_tile) {
            const uint4 v = s16[(tid + 1) * 8 + (0 ^ ((tid + 1) & 7))];
            next = (uint64_t)v.x | ((uint64_t)v.y << 32); have_next = true;
        } else {
            const uint64_t gi = tile_start + first + nv;
            if (gi < a.n) { next = a.ent[gi]; have_next = true; }
        }
    }

    const uint32_t kshift = 64 - lo.FB, pshift = 64 - lo.cmpbits;
    const bool check_taint = lo.cmpbits < lo.FB;
    const uint32_t D2 = 2 * lo.D;
    const uint32_t mshift = 64 - lo.FB - D2;
    const uint32_t colmask = lo.D ? (0xFFFFFFFFu << (4 * (8 - lo.D))) : 0u;
    uint32_t n_closed = 0, n_present = 0;

    auto add_taint = [&](uint64_t pos) {
        const unsigned long long s = atomicAdd(x.n_taint, 1ULL);
        if (s < x.taint_cap) x.taint[s] = pos;
    };

    // ---- serial walk (branch-free bookkeeping: the first closed segment is captured in H with selects,
    //      only a second boundary inside one thread's 16 records — groups shorter than 16 — takes a branch) ----
    KbFA A{}, H{};
    uint64_t cur = 0, hkey = 0;
    uint32_t nseg = 0;
    bool open_left = false, open_right = false;
    if (nv) {
        cur = r[0] >> kshift;
        const uint64_t pk = prev >> kshift;
        open_left = have_prev && pk == cur;
        const bool left_bad = check_taint && have_prev && pk != cur && (prev >> pshift) == (r[0] >> pshift);
        if (left_bad) add_taint(tile_start + first);
        A.start = first; A.bad = left_bad ? 1u : 0u;
        if (tid == 0 && open_left) A.bad = 1u;        // the group began in an earlier tile, which owns it
        nseg = 1;
        uint64_t last = r[0];
#pragma unroll
        for (int j = 0; j < KB_K3F_ITEMS; j++) {
            const bool valid = j < (int)nv;
            const uint64_t e = r[j];
            const uint64_t k = e >> kshift;
            const bool ns = valid && k != cur;                                   // a new segment starts at record j
            const bool taint = ns && check_taint && ((e ^ last) >> pshift) == 0;
            if (taint) add_taint(tile_start + first + j);                         // rare
            if (ns && nseg > 1) {                                                 // rare for groups >= 16 records
                KbFA C = A; C.bad |= taint ? 1u : 0u;
                kb_fa_finalize(x, C, cur, tile_start, n_closed, n_present);
            }
            const bool cap = ns && nseg == 1;
            H.pres = cap ? A.pres : H.pres; H.in = cap ? A.in : H.in; H.out = cap ? A.out : H.out;
            H.cnt = cap ? A.cnt : H.cnt; H.start = cap ? A.start : H.start; H.bad = cap ? (A.bad | (taint ? 1u : 0u)) : H.bad;
            hkey = cap ? cur : hkey;
            A.pres = ns ? 0ULL : A.pres; A.in = ns ? 0u : A.in; A.out = ns ? 0u : A.out; A.cnt = ns ? 0u : A.cnt;
            A.start = ns ? first + j : A.start; A.bad = ns ? (taint ? 1u : 0u) : A.bad;
            cur = ns ? k : cur;
            nseg += ns ? 1u : 0u;
            const uint32_t id = (uint32_t)e & 0xFFu;
            A.pres |= valid ? (1ULL << id) : 0ULL;
            if (D2) {
                uint32_t oh;
                if (D1) oh = 0x10000000u << ((uint32_t)(e >> mshift) & 3u);
                else oh = kb_onehot8(((uint32_t)(e >> mshift) & ((1u << D2) - 1u)) << (16 - D2)) & colmask;
                oh = valid ? oh : 0u;
                const bool isin = (x.ingroup64 >> id) & 1ULL;
                A.in |= isin ? oh : 0u;
                A.out |= isin ? 0u : oh;
            }
            A.cnt += valid ? 1u : 0u;
            last = valid ? e : last;
        }
        const uint64_t nk = next >> kshift;
        open_right = have_next && nk == cur;
        if (check_taint && have_next && nk != cur && (next >> pshift) == (last >> pshift)) A.bad = 1u;
    }

    // ---- segmented inclusive scan of the tail aggregates over the CTA -----------------------------
    KbFA S = A;
    S.head = (nv == 0 || nseg > 1 || !open_left) ? 1u : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const KbFA o = kb_fa_shfl_up(S, d);
        if (lane >= (uint32_t)d && !S.head) S = kb_fa_combine(o, S);
    }
    if (lane == 31) wtot[warp] = S;
    __syncthreads();
    KbFA W{}; bool have_w = false;                      // aggregate flowing in from the previous warps
    for (uint32_t w = 0; w < warp; w++) {
        const KbFA t = wtot[w];
        if (!have_w || t.head) { W = t; have_w = true; } else W = kb_fa_combine(W, t);
    }
    if (have_w && !S.head) S = kb_fa_combine(W, S);     // S is now CTA-inclusive
    KbFA carry = kb_fa_shfl_up(S, 1);
    if (lane == 0) carry = W;                           // (unused when warp == 0: thread 0 never combines)

    // ---- close the first and the last segment of every thread -------------------------------------
    if (nv) {
        if (nseg > 1) {
            const KbFA total = (open_left && tid > 0) ? kb_fa_combine(carry, H) : H;
            kb_fa_finalize(x, total, hkey, tile_start, n_closed, n_present);
        }
        if (!open_right) {
            const KbFA total = (nseg == 1 && open_left && tid > 0) ? kb_fa_combine(carry, A) : A;
            kb_fa_finalize(x, total, cur, tile_start, n_closed, n_present);
        }
    }

    // ---- the group that is still open at the tile end: the last warp reads ahead -------------------
    if (warp == KB_K3F_THREADS / 32 - 1 && n_tile == KB_K3F_TILE) {
        const bool open_tail = __shfl_sync(0xFFFFFFFFu, (int)open_right, 31);
        if (open_tail) {
            const uint64_t tkey = __shfl_sync(0xFFFFFFFFu, cur, 31);
            const uint64_t tpre = __shfl_sync(0xFFFFFFFFu, r[KB_K3F_ITEMS - 1] >> pshift, 31);
            KbFA T = kb_fa_shfl(S, 31);
            uint64_t pres = 0; uint32_t in = 0, out = 0, cnt = 0, bad = 0;
            for (uint64_t i0 = tile_start + KB_K3F_TILE;; i0 += 32) {
                const uint64_t i = i0 + lane;
                const bool valid = i < a.n;
                const uint64_t e = valid ? a.ent[i] : 0ULL;
                const bool same = valid && (e >> kshift) == tkey;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, same);
                const uint32_t stop = ~m ? (uint32_t)__ffs(~m) - 1u : 32u;      // first lane that does not continue the group
                if (lane < stop) {
                    const uint32_t id = (uint32_t)e & 0xFFu;
                    pres |= 1ULL << id;
                    if (D2) {
                        uint32_t oh;
                        if (D1) oh = 0x10000000u << ((uint32_t)(e >> mshift) & 3u);
                        else oh = kb_onehot8(((uint32_t)(e >> mshift) & ((1u << D2) - 1u)) << (16 - D2)) & colmask;
                        if ((x.ingroup64 >> id) & 1ULL) in |= oh; else out |= oh;
                    }
                    cnt++;
                }
                if (stop < 32) {
                    if (check_taint && lane == stop && valid && (e >> pshift) == tpre) bad = 1;
                    break;
                }
            }
            const uint32_t plo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)pres), phi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(pres >> 32));
            in = __reduce_or_sync(0xFFFFFFFFu, in); out = __reduce_or_sync(0xFFFFFFFFu, out);
            cnt = __reduce_add_sync(0xFFFFFFFFu, cnt); bad = __reduce_or_sync(0xFFFFFFFFu, bad);
            if (lane == 31) {
                T.pres |= (uint64_t)plo | ((uint64_t)phi << 32); T.in |= in; T.out |= out; T.cnt += cnt; T.bad |= bad;
                kb_fa_finalize(x, T, tkey, tile_start, n_closed, n_present);
            }
        }
    }

    // ---- statistics -----------------------------------------------------------------------------------
    n_closed = __reduce_add_sync(0xFFFFFFFFu, n_closed);
    n_present = __reduce_add_sync(0xFFFFFFFFu, n_present);
    if (lane == 0) { if (n_closed) atomicAdd(&s_closed, n_closed); if (n_present) atomicAdd(&s_present, n_present); }
    __syncthreads();
    if (tid == 0) {
        if (s_closed) atomicAdd(a.stats + 0, (unsigned long long)s_closed);
        if (s_present) atomicAdd(a.stats + 2, (unsigned long long)s_present);
    }
}

// One warp per tainted boundary: the warp that holds the FIRST tainted boundary of a mixed run
// re-processes the whole run with the ordered sweep.
__global__ void __launch_bounds__(256) kb_group_taint_kernel(const KbFastArgs x) {
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_t = min((uint64_t)*x.n_taint, x.taint_cap);
    const uint32_t kshift = 64 - lo.FB, pshift = 64 - lo.cmpbits;
    for (uint64_t w = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); w < n_t; w += (uint64_t)gridDim.x * 8) {
        const uint64_t p = x.taint[w];
        const uint64_t pre = a.ent[p] >> pshift;
        // run start: walk back while the prefix matches
        uint64_t start = p;
        while (start > 0) {
            const bool in_range = start >= (uint64_t)lane + 1;
            const bool same = in_range && (a.ent[start - 1 - lane] >> pshift) == pre;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, same);
            const uint32_t c = ~m ? (uint32_t)__ffs(~m) - 1u : 32u;
            start -= c;
            if (c < 32) break;
        }
        // is there an earlier key change inside (start, p)?  then another warp owns this run
        bool earlier = false;
        for (uint64_t q0 = start + 1; q0 < p; q0 += 32) {
            const uint64_t q = q0 + lane;
            const bool chg = q < p && (a.ent[q] >> kshift) != (a.ent[q - 1] >> kshift);
            if (__ballot_sync(0xFFFFFFFFu, chg)) { earlier = true; break; }
        }
        if (earlier) continue;
        // run end
        uint64_t end = p + 1;
        while (true) {
            const uint64_t i = end + lane;
            const bool same = i < a.n && (a.ent[i] >> pshift) == pre;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, same);
            const uint32_t c = ~m ? (uint32_t)__ffs(~m) - 1u : 32u;
            end += c;
            if (c < 32) break;
        }
        kb_process_mixed_run<1, 1>(a, start, end - start);
    }
}
