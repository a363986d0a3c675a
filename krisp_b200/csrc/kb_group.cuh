// kb_group.cuh — K3: segmented intersection + diagnostic filter + compaction over the sorted elements.
//
// Replaces, in one pass over the sorted sort-elements (one read of every record):
//   shared.py:210-240 simplifyStream, :442-475 alignmentStream      (grouping by (left,right))
//   shared.py:321-347 intersectSortedStreams folded over all files by
//   intersectAmplicons.py:232-310 mergeFiles                          (group present in EVERY file)
//   filterAlignments.py:4-28 + Amplicon.py:495-521 ingroupUniqueColumns (exists a column whose ingroup
//                                                                      and outgroup base sets are disjoint)
//   Amplicon.py:550-558 consensus / :42-66 collapse_to_iupac            (per-column base sets, as 4-bit masks)
//
// A "run" is a maximal stretch of elements whose top `cmpbits` bits agree (what the radix sort
// guarantees to be contiguous).  When cmpbits covers the whole flank key a run is exactly one group;
// otherwise (prefix / hash sort) a run may hold several groups and they are separated here by
// comparing complete flank keys — grouping is exact in every configuration.
//
// Per CTA tile: (A) run heads -> shared-memory bitmap (lane-neighbour compare + ballot);
// (B) runs too short to contain every file are dropped without touching their records, the rest are
// queued; (C) one warp per queued run: lanes stride over the run's records, accumulate the file
// presence bitmap and the per-column ingroup / outgroup base masks in registers, combine them with
// warp OR-reductions (__reduce_or_sync), test, and append survivors to the result table.
// A run is owned by the tile that holds its head and may extend past the tile end.
#pragma once
#include "kb_common.cuh"

#define KB_K3_THREADS 256
#define KB_K3_ITEMS 16
#define KB_K3_TILE (KB_K3_THREADS * KB_K3_ITEMS)
#define KB_K3_WARPS (KB_K3_THREADS / 32)

struct KbGroupArgs {
    const uint64_t* ent;       // sorted sort-elements
    uint64_t n;
    const uint64_t* recs;      // INDIRECT: records in extraction order
    KbLayout lo;
    uint32_t ingroup[8];       // bit f = file f carries an ingroup label
    uint32_t full[8];          // bit f = file f exists
    unsigned long long* n_res; // survivors found (may exceed cap: caller re-runs with a larger table)
    uint64_t cap;
    uint64_t* res_flank;       // [cap][FW]   MSB-first flank bits (left then right), unmixed
    uint32_t* res_in;          // [cap][MW]
    uint32_t* res_out;         // [cap][MW]
    uint32_t* res_size;        // [cap]       records in the group
    uint64_t* res_run;         // [cap][2]    run start, run length (elements that hold the group's records)
    unsigned long long* stats; // [0] runs  [1] queued runs  [2] groups present in every file  [3] mixed runs
};

template <int FWN> struct KbKey { uint64_t w[FWN]; };

template <int FWN>
__device__ __forceinline__ bool kb_key_eq(const KbKey<FWN>& a, const KbKey<FWN>& b, int FW) {
    bool e = true;
#pragma unroll
    for (int j = 0; j < FWN; j++) if (j < FW) e = e && (a.w[j] == b.w[j]);
    return e;
}
template <int FWN>
__device__ __forceinline__ bool kb_key_lt(const KbKey<FWN>& a, const KbKey<FWN>& b, int FW) {
    bool lt = false, decided = false;
#pragma unroll
    for (int j = 0; j < FWN; j++) if (j < FW && !decided && a.w[j] != b.w[j]) { lt = a.w[j] < b.w[j]; decided = true; }
    return lt;
}
template <int FWN>
__device__ __forceinline__ KbKey<FWN> kb_key_shfl(const KbKey<FWN>& a, int src, int FW) {
    KbKey<FWN> r;
#pragma unroll
    for (int j = 0; j < FWN; j++) r.w[j] = (j < FW) ? __shfl_sync(0xFFFFFFFFu, a.w[j], src) : 0ULL;
    return r;
}

template <int MWN>
struct KbAcc {
    uint32_t pres[8];
    uint32_t in[MWN], out[MWN];
    uint32_t count;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int j = 0; j < 8; j++) pres[j] = 0;
#pragma unroll
        for (int j = 0; j < MWN; j++) { in[j] = 0; out[j] = 0; }
        count = 0;
    }
};

// record of sort-element e -> rec words (the loads) ...
template <int WN>
__device__ __forceinline__ void kb_fetch_rec(const KbGroupArgs& a, uint64_t e, uint64_t (&rec)[WN]) {
    if constexpr (WN == 1) {
        rec[0] = e;
    } else {
        const uint64_t* p = a.recs + (e & 0xFFFFFFFFULL) * WN;
#pragma unroll
        for (int j = 0; j < WN; j += 2) {
            const uint4 v = *reinterpret_cast<const uint4*>(p + j);
            rec[j] = (uint64_t)v.x | ((uint64_t)v.y << 32);
            rec[j + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
        }
    }
}
// ... and its flank key
template <int WN>
__device__ __forceinline__ void kb_key_of(const KbLayout& lo, const uint64_t (&rec)[WN], KbKey<WN>& key) {
    if constexpr (WN == 1) {
        key.w[0] = lo.FB ? (rec[0] >> (64 - lo.FB)) : 0ULL;
    } else {
        const int nb = lo.FB - 64 * (lo.FW - 1);      // flank bits in the last flank word (1..64)
#pragma unroll
        for (int j = 0; j < WN; j++) {
            uint64_t w = (j < lo.FW) ? rec[j] : 0ULL;
            if (j == lo.FW - 1 && nb < 64) w &= ~0ULL << (64 - nb);
            key.w[j] = w;
        }
    }
}
template <int WN>
__device__ __forceinline__ void kb_fetch(const KbGroupArgs& a, uint64_t e, uint64_t (&rec)[WN], KbKey<WN>& key) {
    kb_fetch_rec<WN>(a, e, rec);
    kb_key_of<WN>(a.lo, rec, key);
}

template <int WN, int MWN>
__device__ __forceinline__ void kb_accumulate(const KbGroupArgs& a, const uint64_t (&rec)[WN], KbAcc<MWN>& acc) {
    const KbLayout& lo = a.lo;
    const uint32_t id = (uint32_t)rec[WN - 1] & 0xFFu;
    const uint32_t bit = 1u << (id & 31);
#pragma unroll
    for (int j = 0; j < 8; j++) if (j < lo.PW && (int)(id >> 5) == j) acc.pres[j] |= bit;
    const bool isin = (a.ingroup[id >> 5] >> (id & 31)) & 1u;
#pragma unroll
    for (int j = 0; j < MWN; j++) {
        if (j < lo.MW) {
            const int ncol = min(8, lo.D - 8 * j);
            const uint32_t v = (uint32_t)kb_rec_bits<WN>(rec, lo.FB + 16 * j, 2 * ncol) << (16 - 2 * ncol);
            const uint32_t oh = kb_onehot8(v) & (0xFFFFFFFFu << (4 * (8 - ncol)));
            if (isin) acc.in[j] |= oh; else acc.out[j] |= oh;
        }
    }
    acc.count++;
}

template <int MWN>
__device__ __forceinline__ void kb_acc_reduce(const KbLayout& lo, KbAcc<MWN>& acc) {
#pragma unroll
    for (int j = 0; j < 8; j++) if (j < lo.PW) acc.pres[j] = __reduce_or_sync(0xFFFFFFFFu, acc.pres[j]);
#pragma unroll
    for (int j = 0; j < MWN; j++) if (j < lo.MW) {
        acc.in[j] = __reduce_or_sync(0xFFFFFFFFu, acc.in[j]);
        acc.out[j] = __reduce_or_sync(0xFFFFFFFFu, acc.out[j]);
    }
    acc.count = __reduce_add_sync(0xFFFFFFFFu, acc.count);
}

// S6 (present in every file) and S7 (a column with disjoint ingroup / outgroup base sets; only when D > 0)
template <int MWN>
__device__ __forceinline__ void kb_evaluate(const KbGroupArgs& a, const KbAcc<MWN>& acc, bool& present, bool& diag) {
    const KbLayout& lo = a.lo;
    present = true;
#pragma unroll
    for (int j = 0; j < 8; j++) if (j < lo.PW) present = present && (acc.pres[j] == a.full[j]);
    diag = (lo.D == 0);
#pragma unroll
    for (int j = 0; j < MWN; j++) if (j < lo.MW) {
        const int ncol = min(8, lo.D - 8 * j);
        uint32_t y = acc.in[j] & acc.out[j];
        y |= y >> 1; y |= y >> 2;                                   // bit 0 of each nibble = nibble non-zero
        const uint32_t z = ~y & 0x11111111u & (0xFFFFFFFFu << (4 * (8 - ncol)));
        diag = diag || (z != 0);
    }
}

template <int WN, int MWN>
__device__ __forceinline__ void kb_emit(const KbGroupArgs& a, const KbKey<WN>& key, const KbAcc<MWN>& acc,
                                        uint64_t run_start, uint64_t run_len) {
    const KbLayout& lo = a.lo;
    const unsigned long long slot = atomicAdd(a.n_res, 1ULL);
    if (slot >= a.cap) return;
    if constexpr (WN == 1) {
        uint64_t kk = key.w[0];
        if (lo.FB) { if (lo.mix) kk = kb_unmix(kk, lo.FB, lo.shs); kk <<= (64 - lo.FB); }
        a.res_flank[slot] = kk;
    } else {
#pragma unroll
        for (int j = 0; j < WN; j++) if (j < lo.FW) a.res_flank[slot * lo.FW + j] = key.w[j];
    }
#pragma unroll
    for (int j = 0; j < MWN; j++) if (j < lo.MW) { a.res_in[slot * lo.MW + j] = acc.in[j]; a.res_out[slot * lo.MW + j] = acc.out[j]; }
    a.res_size[slot] = acc.count;
    a.res_run[2 * slot] = run_start;
    a.res_run[2 * slot + 1] = run_len;
}

// A run whose records carry several distinct flank keys (they share only the sort prefix): visit the
// keys in increasing order, one sweep over the run per key.  One warp; all lanes call it.
template <int WN, int MWN>
__device__ __noinline__ void kb_process_mixed_run(const KbGroupArgs& a, uint64_t start, uint64_t len) {
    const KbLayout& lo = a.lo;
    const uint32_t lane = threadIdx.x & 31;
    if (lane == 0) atomicAdd(a.stats + 3, 1ULL);
    KbAcc<MWN> acc;
    KbKey<WN> cur; bool have_cur = false;
#pragma unroll
    for (int j = 0; j < WN; j++) cur.w[j] = 0;
    for (int round = 0;; round++) {
        acc.reset();
        KbKey<WN> best; bool have_best = false;   // smallest key > cur (round 0: smallest key)
#pragma unroll
        for (int j = 0; j < WN; j++) best.w[j] = 0;
        for (uint64_t off = 0; off < len; off += 32) {
            const bool valid = off + lane < len;
            uint64_t rec[WN]; KbKey<WN> key;
#pragma unroll
            for (int j = 0; j < WN; j++) { rec[j] = 0; key.w[j] = 0; }
            if (valid) kb_fetch<WN>(a, a.ent[start + off + lane], rec, key);
            if (valid) {
                if (have_cur && kb_key_eq<WN>(key, cur, lo.FW)) kb_accumulate<WN, MWN>(a, rec, acc);
                else if ((!have_cur || kb_key_lt<WN>(cur, key, lo.FW)) && (!have_best || kb_key_lt<WN>(key, best, lo.FW))) { best = key; have_best = true; }
            }
        }
        if (have_cur) {
            kb_acc_reduce<MWN>(lo, acc);
            bool present, diag;
            kb_evaluate<MWN>(a, acc, present, diag);
            if (lane == 0 && present) {
                atomicAdd(a.stats + 2, 1ULL);
                if (diag) kb_emit<WN, MWN>(a, cur, acc, start, len);
            }
        }
        // warp minimum of `best`
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            KbKey<WN> o;
#pragma unroll
            for (int j = 0; j < WN; j++) o.w[j] = __shfl_xor_sync(0xFFFFFFFFu, best.w[j], d);
            const bool oh = __shfl_xor_sync(0xFFFFFFFFu, (int)have_best, d);
            if (oh && (!have_best || kb_key_lt<WN>(o, best, lo.FW))) { best = o; have_best = true; }
        }
        if (!have_best) break;
        cur = best; have_cur = true;
    }
}

template <int WN, int MWN>
__global__ void __launch_bounds__(KB_K3_THREADS) kb_group_kernel(const KbGroupArgs a) {
    const KbLayout& lo = a.lo;
    __shared__ uint32_t hb[KB_K3_TILE / 32 + 1];     // head bitmap (+ sentinel word)
    __shared__ uint32_t queue[KB_K3_TILE];           // start (12 bits) | length (13 bits, 0 = open-ended) << 12
    __shared__ uint32_t qn, s_runs;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_start = (uint64_t)blockIdx.x * KB_K3_TILE;
    if (tile_start >= a.n) return;
    const uint32_t n_tile = (uint32_t)min((uint64_t)KB_K3_TILE, a.n - tile_start);
    const uint32_t cshift = 64 - lo.cmpbits;         // cmpbits == 0 -> every prefix is 0
    auto prefix = [&](uint64_t e) -> uint64_t { return lo.cmpbits ? (e >> cshift) : 0ULL; };

    if (tid == 0) { qn = 0; s_runs = 0; hb[KB_K3_TILE / 32] = 0; }

    // ---- A. run heads ---------------------------------------------------------------------------
#pragma unroll 4
    for (int c = 0; c < KB_K3_ITEMS; c++) {
        const uint32_t q = c * KB_K3_THREADS + tid;
        const uint64_t i = tile_start + q;
        const bool valid = q < n_tile;
        const uint64_t e = valid ? a.ent[i] : 0ULL;
        uint64_t pe = __shfl_up_sync(0xFFFFFFFFu, e, 1);
        if (lane == 0 && valid && i > 0) pe = a.ent[i - 1];
        const bool head = valid && (i == 0 || prefix(e) != prefix(pe));
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, head);
        if (lane == 0) hb[c * KB_K3_WARPS + warp] = m;
    }
    __syncthreads();

    // ---- B. queue the runs that can contain every file -----------------------------------------
    if (tid < KB_K3_TILE / 32) {
        uint32_t w = hb[tid], nh = __popc(w);
        while (w) {
            const uint32_t b = __ffs(w) - 1;
            w &= w - 1;
            const uint32_t q = tid * 32 + b;
            uint32_t nxt = 0;                         // tile-relative position of the next head, 0 = none in tile
            if (w) nxt = tid * 32 + (__ffs(w) - 1);
            else {
                for (uint32_t t = tid + 1; t < KB_K3_TILE / 32; t++) { const uint32_t x = hb[t]; if (x) { nxt = t * 32 + (__ffs(x) - 1); break; } }
            }
            uint32_t len = 0;
            if (nxt) len = nxt - q;
            else if (tile_start + n_tile == a.n) len = n_tile - q;   // last tile: the run ends with the data
            if (len == 0 || len >= (uint32_t)lo.n_files) queue[atomicAdd(&qn, 1u)] = q | (len << 12);
        }
        if (nh) atomicAdd(&s_runs, nh);
    }
    __syncthreads();
    const uint32_t nq = qn;
    if (tid == 0) { atomicAdd(a.stats + 0, (unsigned long long)s_runs); atomicAdd(a.stats + 1, (unsigned long long)nq); }

    // ---- C. one warp per queued run --------------------------------------------------------------
    for (uint32_t qi = warp; qi < nq; qi += KB_K3_WARPS) {
        const uint32_t qe = queue[qi];
        const uint64_t start = tile_start + (qe & 0xFFFu);
        uint32_t klen = qe >> 12;                     // 0 = open-ended (continues past the tile)
        const uint64_t runpre = prefix(a.ent[start]);

        KbAcc<MWN> acc; acc.reset();
        KbKey<WN> pivot;
        bool mixed = false;
        uint64_t len = 0;
        for (uint64_t off = 0;; off += 32) {
            const uint64_t i = start + off + lane;
            bool valid = klen ? (off + lane < klen) : (i < a.n);
            uint64_t e = 0;
            if (valid) { e = a.ent[i]; if (!klen) valid = prefix(e) == runpre; }
            uint64_t rec[WN]; KbKey<WN> key;
#pragma unroll
            for (int j = 0; j < WN; j++) { rec[j] = 0; key.w[j] = 0; }
            if (valid) kb_fetch<WN>(a, e, rec, key);
            if (off == 0) pivot = kb_key_shfl<WN>(key, 0, lo.FW);
            const bool same = valid && kb_key_eq<WN>(key, pivot, lo.FW);
            if (same) kb_accumulate<WN, MWN>(a, rec, acc);
            const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
            if (__ballot_sync(0xFFFFFFFFu, valid && !same)) mixed = true;
            len += __popc(vm);
            if (vm != 0xFFFFFFFFu) break;             // an invalid lane: the run ended in this chunk
        }
        if (len < (uint64_t)lo.n_files) continue;     // (open-ended run that turned out short)

        if (!mixed) {
            kb_acc_reduce<MWN>(lo, acc);
            bool present, diag;
            kb_evaluate<MWN>(a, acc, present, diag);
            if (lane == 0 && present) {
                atomicAdd(a.stats + 2, 1ULL);
                if (diag) kb_emit<WN, MWN>(a, pivot, acc, start, len);
            }
            continue;
        }

        kb_process_mixed_run<WN, MWN>(a, start, len);
    }
}

// ---- records of the surviving groups' runs (for --out_align) -----------------------------------
// out[r] for r in [off[g], off[g] + runlen[g]) = records of run g, WN words each.
struct KbGatherArgs {
    const uint64_t* ent;
    const uint64_t* recs;
    const uint64_t* res_run;   // [n_groups][2]
    const uint64_t* off;       // [n_groups] exclusive prefix of run lengths
    uint64_t n_groups;
    uint64_t* out;
    KbLayout lo;
};

template <int WN>
__global__ void __launch_bounds__(256) kb_gather_kernel(const KbGatherArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= a.n_groups) return;
    const uint64_t start = a.res_run[2 * g], len = a.res_run[2 * g + 1], o = a.off[g];
    for (uint64_t i = lane; i < len; i += 32) {
        const uint64_t e = a.ent[start + i];
        if constexpr (WN == 1) {
            uint64_t v = e;
            if (a.lo.FB && a.lo.mix) {
                const uint64_t low = a.lo.FB < 64 ? (v & kb_lowmask(64 - a.lo.FB)) : 0ULL;
                v = (kb_unmix(v >> (64 - a.lo.FB), a.lo.FB, a.lo.shs) << (64 - a.lo.FB)) | low;
            }
            a.out[o + i] = v;
        } else {
            const uint64_t* p = a.recs + (e & 0xFFFFFFFFULL) * WN;
#pragma unroll
            for (int j = 0; j < WN; j++) a.out[(o + i) * WN + j] = p[j];
        }
    }
}
