// kb_hash_warp.cuh — K3: bucket hash aggregation with per-warp streaming (one-word records, D <= 8, up to 256 files).
//
// Same contract and per-bucket algorithm as kb_hash_stream_kernel (kb_hash_stream.cuh; reference stages: simplifyStream /
// alignmentStream shared.py:210-240,:442-475, intersectSortedStreams :321-347 folded by mergeFiles
// intersectAmplicons.py:232-310, filterAlignments.py:4-28 + ingroupUniqueColumns Amplicon.py:495-521).  What differs:
//   * every warp streams its share of the records through its own small ring (KB_HW_STAGES chunks of 128 records, filled
//     with 16-byte cp.async copies that run two chunks ahead, across bucket boundaries): no mbarrier, no producer warp, and
//     little shared memory per warp, so 24 warps fit an SM;
//   * a chunk never straddles two buckets, so the hot loop has no boundary logic: four records per lane, their four home
//     slots are loaded together, hit / miss is resolved for all four, misses go to the warp's queue and are inserted 32 at a
//     time;
//   * table in structure-of-arrays form (keys 8 B apart, presence words apart from them): key loads spread over all banks;
//   * buckets may be exact ranges (bstart[b], bstart[b + 1]) or slabs (b * bcap, filled up to bend[b]) — kb_extract_part.cuh.
// Two table arrangements:
//   SHARED = false  every WARP owns its buckets and a private table: no CTA barrier anywhere, a bucket end (scan + emit +
//                   clear) stalls one warp.  Best for small buckets (a 256-slot table is 4 KB).
//   SHARED = true   the 8 warps of a CTA share one table and one bucket at a time, warp w streams chunks w, w + 8, ... of it;
//                   three named barriers per bucket.  For buckets too big for a private table (fewer bucket bits: two
//                   partition levels still suffice when level 0 also has to separate the owner GPUs).
// Buckets whose distinct keys overflow the table or that hold more than KB_HW_MAX_INLINE survivors are deferred to
// kb_hash_fast_kernel / kb_hash_kernel exactly like in the stream kernel; group sizes come from kb_hsize_kernel.
#pragma once
#include <type_traits>
#include "kb_hash_stream.cuh"

#define KB_HW_PER 4                           // records per lane and chunk
#define KB_HW_CH (32 * KB_HW_PER)             // 128 records per chunk
#define KB_HW_STAGES 3
#define KB_HW_QCAP (32 + KB_HW_CH)            // the queue is drained to < 32 entries after every chunk
#define KB_HW_MAX_INLINE 8
#define KB_HW_WARPS 8                         // warps per CTA (per-warp tables; the shared-table arrangement also runs with KB_HW_WARPS_WIDE)
#define KB_HW_WARPS_WIDE 10                   // shared table: 10 warps x 3 CTAs = 30 warps per SM still fit the shared memory (2048-slot table)
#define KB_HW_THREADS (32 * KB_HW_WARPS)

struct KbHWarpArgs {
    KbHashArgs h;                        // g.ent = partitioned elements; buckets via h.bstart / h.bend / h.bcap
    uint32_t* deferred;                  // [n_buckets] bucket ids left to the fallback kernel
    unsigned long long* n_deferred;
    uint32_t wbytes;                     // shared memory per warp
    uint32_t tbytes;                     // SHARED: bytes of the CTA's table + control words (in front of the warp regions)
};

static inline uint32_t kb_hash_table_bytes(uint32_t slots_log2, int pwn, bool packed) {
    const uint32_t S = 1u << slots_log2;
    return 8u * S + 4u * S * (uint32_t)pwn + (packed ? 0u : 8u * S);
}
// bytes of one warp's region: ring | queue (| private table)
static inline uint32_t kb_hash_warp_wbytes(uint32_t slots_log2, int pwn, bool packed, bool shared) {
    const uint32_t b = 8u * (KB_HW_STAGES * KB_HW_CH + KB_HW_QCAP) + (shared ? 0u : kb_hash_table_bytes(slots_log2, pwn, packed));
    return (b + 15u) & ~15u;
}
static inline uint32_t kb_hash_cta_tbytes(uint32_t slots_log2, int pwn, bool packed) { return kb_hash_table_bytes(slots_log2, pwn, packed) + 64u; }

template <int NT>
__device__ __forceinline__ void kb_hw_bar() { asm volatile("bar.sync 1, %0;" :: "n"(NT) : "memory"); }

template <bool D1, bool SPACER, int PWN, bool SHARED, int NW = KB_HW_WARPS>
__global__ void __launch_bounds__(32 * NW, 3) kb_hash_warp_kernel(const KbHWarpArgs xs) {
    constexpr uint32_t NTHREADS = 32u * NW;
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    const KbHashArgs& x = xs.h;
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    const uint32_t S = 1u << x.slots_log2, smask = S - 1u;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // bucket / chunk ownership
    const uint32_t b_first = SHARED ? blockIdx.x : blockIdx.x * NW + warp;
    const uint32_t b_step = SHARED ? gridDim.x : gridDim.x * NW;
    const uint32_t c_first = SHARED ? warp : 0u;
    constexpr uint32_t c_step = SHARED ? (uint32_t)NW : 1u;
    const bool packed = SPACER || (D1 && lo.FB <= 54);       // both 4-bit base sets inside the key word (bits 56-63)

    uint32_t smem_a = kb_smem_u32(kb_smem_raw);
    asm volatile("" : "+r"(smem_a));                          // opaque: keep the base in a register
    const uint32_t warp_a = smem_a + (SHARED ? xs.tbytes : 0u) + warp * xs.wbytes;
    const uint32_t ring_a = warp_a;
    const uint32_t q_a = ring_a + KB_HW_STAGES * KB_HW_CH * 8;
    const uint32_t keys_a = SHARED ? smem_a : q_a + KB_HW_QCAP * 8;
    const uint32_t pres_a = keys_a + S * 8;
    const uint32_t msk_a = pres_a + S * 4 * PWN;              // [S][2] (unpacked sets only)
    const uint32_t ctl_a = msk_a + (packed ? 0u : S * 8);     // SHARED: per bucket parity p: [p] over, [2 + p] distinct keys, [4 + p] survivors

    // table slots this thread scans / clears
    const uint32_t t_first = SHARED ? tid : lane;
    constexpr uint32_t t_step = SHARED ? NTHREADS : 32u;
    for (uint32_t i = t_first; i < S; i += t_step) {
        kb_sts64(keys_a + i * 8, KB_KH_EMPTY);
#pragma unroll
        for (int j = 0; j < PWN; j++) kb_sts32(pres_a + (i * PWN + j) * 4, 0u);
        if (!packed) kb_sts64(msk_a + i * 8, 0ULL);
    }
    if (SHARED) { if (tid < 16) kb_sts32(ctl_a + 4 * tid, 0u); __syncthreads(); }
    else __syncwarp();

    const uint32_t kshift = SPACER ? 10u : 64 - lo.FB;
    const uint32_t D2 = SPACER ? 2u : 2 * lo.D;
    const uint32_t mshift = SPACER ? 8u : 64 - lo.FB - D2;
    const uint32_t colmask = lo.D ? (0xFFFFFFFFu << (4 * (8 - lo.D))) : 0u;
    const uint32_t limit = S - (S >> 2);
    const uint32_t hmask = SPACER ? 0xFFFFFFFFu : kb_kh_hmask((uint32_t)lo.FB, x.bb);
    const uint32_t ing_lo = (uint32_t)x.ingroup64, ing_hi = (uint32_t)(x.ingroup64 >> 32);
    const uint32_t sshift = 32 - x.slots_log2;
    const uint32_t lt_mask = kb_lanemask_lt();
    uint32_t qn = 0;                     // records waiting in the queue (warp-uniform)
    uint32_t nkeys = 0;                  // private table: distinct keys in it (warp-uniform)
    bool over = false;                   // the current bucket is given up (warp-uniform)
    uint32_t par = 0;                    // SHARED: parity of the current bucket (selects its control words)

    auto bucket_range = [&](uint32_t bl, uint64_t& s, uint64_t& e) {
        const uint32_t b = x.bucket0 + bl;
        if (x.bcap) { s = (uint64_t)b * x.bcap; e = min((uint64_t)x.bend[b], s + x.bcap); }
        else { s = x.bstart[b]; e = x.bend ? x.bend[b] : x.bstart[b + 1]; }
        if (e < s) e = s;
    };
    // first non-empty bucket at or after `from` (stride b_step); nb >= n_buckets: none
    auto next_bucket = [&](uint32_t from, uint32_t& nb, uint64_t& s, uint64_t& e) {
        nb = from; s = 0; e = 0;
        while (nb < x.n_buckets) {
            bucket_range(nb, s, e);
            if (e > s) break;
            nb += b_step;
        }
    };
    auto slot_of = [&](uint64_t e) -> uint32_t { return (kb_kh_bits(e, x.bb, hmask) * 0x9E3779B1u) >> sshift; };

    // record -> its bits in the table (slot index, khi = high half of the key word as last read)
    auto accumulate = [&](uint32_t slot, uint64_t e, uint32_t khi) {
        const uint32_t id = (uint32_t)e & 0xFFu;
        kb_reds_or(pres_a + (slot * PWN + (id >> 5)) * 4, 1u << (id & 31));
        if (D2) {
            const uint32_t isin = PWN == 2 ? ((((id & 32u) ? ing_hi : ing_lo) >> (id & 31)) & 1u) : ((a.ingroup[id >> 5] >> (id & 31)) & 1u);
            if (packed) {
                const uint32_t bit = (isin ? 0x01000000u : 0x10000000u) << ((uint32_t)(e >> mshift) & 3u);
                if (!(khi & bit)) kb_reds_or(keys_a + slot * 8 + 4, bit);
            } else {
                uint32_t oh;
                if (D1) oh = 0x10000000u << ((uint32_t)(e >> mshift) & 3u);
                else oh = kb_onehot8(((uint32_t)(e >> mshift) & ((1u << D2) - 1u)) << (16 - D2)) & colmask;
                const uint32_t ma = msk_a + slot * 8 + 4 - 4 * isin;           // [0] ingroup sets, [1] outgroup sets
                if ((kb_lds32(ma) & oh) != oh) kb_reds_or(ma, oh);
            }
        }
    };
    // full probe: 0 = key found, 1 = new key inserted, 2 = table full
    auto insert = [&](uint64_t e) -> uint32_t {
        const uint64_t key = e >> kshift;
        uint32_t slot = slot_of(e);
        for (uint32_t step = 0; step <= S; step++) {
            const uint32_t sa = keys_a + slot * 8;
            uint64_t k = kb_lds64(sa);
            uint32_t fresh = 0;
            if (k == KB_KH_EMPTY) {
                k = kb_atoms_cas64(sa, KB_KH_EMPTY, key);
                if (k == KB_KH_EMPTY) { k = key; fresh = 1; }
            }
            if ((packed ? (k & KB_HS_KEYMASK) : k) == key) { accumulate(slot, e, (uint32_t)(k >> 32)); return fresh; }
            slot = (slot + 1) & smask;
        }
        return 2u;
    };
    auto drain = [&]() {                 // up to 32 queued records, one per lane
        const uint32_t cnt = min(qn, 32u);
        qn -= cnt;
        uint32_t st = 0;
        if (lane < cnt) st = insert(kb_lds64(q_a + (qn + lane) * 8));
        __syncwarp();
        const uint32_t fresh = __popc(__ballot_sync(0xFFFFFFFFu, st == 1u));
        const bool full = __any_sync(0xFFFFFFFFu, st == 2u);
        if (SHARED) {
            uint32_t total = 0;
            if (lane == 0 && fresh) total = kb_atoms_add(ctl_a + 4 * (2 + par), fresh) + fresh;
            if (lane == 0 && (full || total > limit)) kb_sts32(ctl_a + 4 * par, 1u);
            __syncwarp();
            over = over || full || __any_sync(0xFFFFFFFFu, kb_lds32(ctl_a + 4 * par));   // (another warp may have given the bucket up)
        } else {
            nkeys += fresh;
            if (full || nkeys > limit) over = true;
        }
    };

    // ---- producer side: this warp's chunk sequence over its buckets, 16-byte cp.async pieces, one commit group per consumer
    //      iteration (empty once the chunks are exhausted, so that the group arithmetic stays put) -----------------------------
    const uint64_t* ent = a.ent;
    uint32_t pb, pk = 0, pnch = 0, pq = 0;
    uint64_t ps_al = 0, pn_al = 0;
    auto prod_enter = [&](uint32_t from) {      // first bucket at or after `from` that has a chunk for this warp
        pb = from; pk = c_first; pnch = 0;
        while (pb < x.n_buckets) {
            uint64_t s, e;
            bucket_range(pb, s, e);
            if (e > s) {
                ps_al = s & ~1ULL; pn_al = e - ps_al; pnch = (uint32_t)((pn_al + KB_HW_CH - 1) / KB_HW_CH);
                if (pnch > c_first) break;
            }
            pb += b_step;
        }
    };
    prod_enter(b_first);
    auto issue = [&]() {
        if (pb < x.n_buckets) {
            const uint32_t here = (uint32_t)min((uint64_t)KB_HW_CH, pn_al - (uint64_t)pk * KB_HW_CH);
            const uint64_t* src = ent + ps_al + (uint64_t)pk * KB_HW_CH;
            const uint32_t dst = ring_a + (pq % KB_HW_STAGES) * (KB_HW_CH * 8);
#pragma unroll
            for (int h = 0; h < KB_HW_CH / 64; h++) {
                const uint32_t piece = h * 32 + lane;                            // records 2 piece, 2 piece + 1
                if (2 * piece < here)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + piece * 16), "l"(src + 2 * piece) : "memory");
            }
            pq++;
            pk += c_step;
            if (pk >= pnch) prod_enter(pb + b_step);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int s = 0; s < KB_HW_STAGES - 1; s++) issue();

    uint32_t b;
    uint64_t bs, be;
    next_bucket(b_first, b, bs, be);
    uint32_t cq = 0;
    uint32_t n_closed_t = 0, n_present_t = 0, n_rounds = 0, n_defer = 0;
    while (b < x.n_buckets) {
        const uint64_t s_al = bs & ~1ULL;
        const uint32_t skip = (uint32_t)(bs - s_al);
        const uint64_t n_al = be - s_al;
        const uint32_t nch = (uint32_t)((n_al + KB_HW_CH - 1) / KB_HW_CH);
        over = false; nkeys = 0; qn = 0;
        for (uint32_t ck = c_first; ck < nch; ck += c_step) {
            issue();                                                             // chunk cq + STAGES - 1 (into the stage freed last iteration)
            asm volatile("cp.async.wait_group %0;" :: "n"(KB_HW_STAGES - 1) : "memory");
            __syncwarp();
            if (SHARED) over = over || __any_sync(0xFFFFFFFFu, kb_lds32(ctl_a + 4 * par));
            const uint32_t stage_a = ring_a + (cq % KB_HW_STAGES) * (KB_HW_CH * 8);
            const uint32_t cnt = (uint32_t)min((uint64_t)KB_HW_CH, n_al - (uint64_t)ck * KB_HW_CH);
            const uint32_t first = ck == 0 ? skip : 0u;
            if (!over) {
                // FULL: a whole chunk of 128 records (all but the first / last chunk of a bucket) — no per-record bounds tests
                auto chunk = [&](auto full_c) {
                    constexpr bool FULL = decltype(full_c)::value;
                    uint64_t e[KB_HW_PER], k[KB_HW_PER];
                    uint32_t sl[KB_HW_PER], m[KB_HW_PER];
                    bool hit[KB_HW_PER], miss[KB_HW_PER];
#pragma unroll
                    for (int j = 0; j < KB_HW_PER; j++) e[j] = kb_lds64(stage_a + (j * 32 + lane) * 8);
#pragma unroll
                    for (int j = 0; j < KB_HW_PER; j++) { sl[j] = slot_of(e[j]); k[j] = kb_lds64(keys_a + sl[j] * 8); }
#pragma unroll
                    for (int j = 0; j < KB_HW_PER; j++) {
                        const uint32_t idx = (uint32_t)(j * 32) + lane;
                        const bool act = FULL || (idx >= first && idx < cnt);
                        hit[j] = act && (packed ? (k[j] & KB_HS_KEYMASK) : k[j]) == (e[j] >> kshift);
                        miss[j] = act && !hit[j];
                        m[j] = __ballot_sync(0xFFFFFFFFu, miss[j]);
                    }
#pragma unroll
                    for (int j = 0; j < KB_HW_PER; j++) {
                        if (miss[j]) kb_sts64(q_a + (qn + __popc(m[j] & lt_mask)) * 8, e[j]);      // (no branch around the whole step: most chunks have a miss in every row)
                        qn += __popc(m[j]);
                        if (hit[j]) accumulate(sl[j], e[j], (uint32_t)(k[j] >> 32));
                    }
                };
                if (first == 0u && cnt == (uint32_t)KB_HW_CH) chunk(std::true_type{});
                else chunk(std::false_type{});
                __syncwarp();
                while (qn >= 32 && !over) drain();
            }
            __syncwarp();                                                        // every lane is done with this stage
            cq++;
        }
        // ---- this warp's share of bucket b is done: flush the queue; then evaluate, emit, clear --------------------------------
        __syncwarp();
        while (qn && !over) drain();
        __syncwarp();
        if (SHARED) { kb_hw_bar<NTHREADS>(); over = kb_lds32(ctl_a + 4 * par) != 0; }      // (uniform over the CTA from here on)
        n_rounds++;
        uint32_t flags = 0, n_surv = 0, n_closed = 0, n_present = 0;
        if (!over) {
            for (uint32_t q = 0, slot = t_first; slot < S; q++, slot += t_step) {
                const uint64_t kk0 = kb_lds64(keys_a + slot * 8);
                if (kk0 == KB_KH_EMPTY) continue;
                n_closed++;
                bool present = true;
#pragma unroll
                for (int j = 0; j < PWN; j++) present = present && (kb_lds32(pres_a + (slot * PWN + j) * 4) == a.full[j]);
                if (!present) continue;
                n_present++;
                bool ok = true;
                if (lo.D) {
                    const uint32_t m_in = packed ? ((uint32_t)(kk0 >> 56) & 0xFu) << 28 : kb_lds32(msk_a + slot * 8);
                    const uint32_t m_out = packed ? ((uint32_t)(kk0 >> 60) & 0xFu) << 28 : kb_lds32(msk_a + slot * 8 + 4);
                    uint32_t y = m_in & m_out;
                    y |= y >> 1; y |= y >> 2;
                    ok = (~y & 0x11111111u & colmask) != 0;
                }
                if (ok) flags |= 1u << q;
            }
            n_surv = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(flags));
            if (SHARED) {
                if (lane == 0 && n_surv) kb_atoms_add(ctl_a + 4 * (4 + par), n_surv);
                kb_hw_bar<NTHREADS>();
                n_surv = kb_lds32(ctl_a + 4 * (4 + par));
            }
        } else if (SHARED) kb_hw_bar<NTHREADS>();
        const bool defer = over || n_surv > KB_HW_MAX_INLINE;
        if (!defer) { n_closed_t += n_closed; n_present_t += n_present; }
        for (uint32_t q = 0, slot = t_first; slot < S; q++, slot += t_step) {
            const uint64_t kk0 = kb_lds64(keys_a + slot * 8);
            if (kk0 == KB_KH_EMPTY) continue;
            if (!defer && ((flags >> q) & 1u)) {
                const unsigned long long gs = atomicAdd(a.n_res, 1ULL);
                if (gs < a.cap) {
                    uint64_t kk = packed ? (kk0 & KB_HS_KEYMASK) : kk0;
                    if (lo.mix) kk = kb_unmix(kk, lo.FB, lo.shs);
                    a.res_flank[gs] = kk << kshift;
                    if (lo.MW) {
                        a.res_in[gs] = packed ? ((uint32_t)(kk0 >> 56) & 0xFu) << 28 : kb_lds32(msk_a + slot * 8);
                        a.res_out[gs] = packed ? ((uint32_t)(kk0 >> 60) & 0xFu) << 28 : kb_lds32(msk_a + slot * 8 + 4);
                    }
                    a.res_run[2 * gs] = bs;
                    a.res_run[2 * gs + 1] = be - bs;
                    a.res_size[gs] = 0xFFFFFFFFu;                    // filled by kb_hsize_kernel
                }
            }
            kb_sts64(keys_a + slot * 8, KB_KH_EMPTY);
#pragma unroll
            for (int j = 0; j < PWN; j++) kb_sts32(pres_a + (slot * PWN + j) * 4, 0u);
            if (!packed) kb_sts64(msk_a + slot * 8, 0ULL);
        }
        if (defer && (SHARED ? tid == 0 : lane == 0)) { const unsigned long long d = atomicAdd(xs.n_deferred, 1ULL); xs.deferred[d] = x.bucket0 + b; }
        if (defer && (SHARED ? warp == 0 : true)) n_defer++;
        if (SHARED) {
            if (tid == 0) { kb_sts32(ctl_a + 4 * (par ^ 1u), 0u); kb_sts32(ctl_a + 4 * (2 + (par ^ 1u)), 0u); kb_sts32(ctl_a + 4 * (4 + (par ^ 1u)), 0u); }
            kb_hw_bar<NTHREADS>();                                                         // table and the next bucket's control words are clean
            par ^= 1u;
            if (warp != 0) n_rounds--;                                           // (one count per CTA)
        } else __syncwarp();
        next_bucket(b + b_step, b, bs, be);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    n_closed_t = __reduce_add_sync(0xFFFFFFFFu, n_closed_t);
    n_present_t = __reduce_add_sync(0xFFFFFFFFu, n_present_t);
    if (lane == 0) {
        if (n_closed_t) atomicAdd(a.stats + 0, (unsigned long long)n_closed_t);
        if (n_rounds - n_defer) atomicAdd(a.stats + 1, (unsigned long long)(n_rounds - n_defer));
        if (n_present_t) atomicAdd(a.stats + 2, (unsigned long long)n_present_t);
    }
}
