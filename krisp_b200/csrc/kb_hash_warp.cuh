// kb_hash_warp.cuh — K3, warp-private version of the bucket hash aggregation (one-word records, D <= 8, up to 256 files).
//
// Same contract and per-bucket algorithm as kb_hash_stream_kernel (kb_hash_stream.cuh; reference stages: simplifyStream /
// alignmentStream shared.py:210-240,:442-475, intersectSortedStreams :321-347 folded by mergeFiles
// intersectAmplicons.py:232-310, filterAlignments.py:4-28 + ingroupUniqueColumns Amplicon.py:495-521), but every WARP owns
// its buckets, its hash table, its miss queue and its own ring of bulk-copy stages:
//   * no CTA barrier anywhere: a bucket end (scan + emit + clear) stalls one warp, not 256 threads;
//   * a stage is 256 records and never straddles two buckets, so the hot loop has no boundary logic: eight records per
//     lane are loaded, their eight home slots are loaded, then hit / miss is resolved (misses go to the warp's queue and
//     are inserted 32 at a time, as before);
//   * table in structure-of-arrays form (keys 8 B apart, presence words apart from them): key loads spread over all banks;
//   * buckets may be exact ranges (bstart[b], bstart[b + 1]) or slabs (b * bcap, filled up to bend[b]) — kb_extract_part.cuh.
// Buckets whose distinct keys overflow the table or that hold more than KB_HW_MAX_INLINE survivors are deferred to
// kb_hash_fast_kernel / kb_hash_kernel exactly like in the stream kernel; group sizes come from kb_hsize_kernel.
#pragma once
#include "kb_hash_stream.cuh"

#define KB_HW_CH 256                          // records per stage (2 KB)
#define KB_HW_PER (KB_HW_CH / 32)
#define KB_HW_STAGES 4
#define KB_HW_QCAP 64
#define KB_HW_MAX_INLINE 8
#define KB_HW_MAXWARPS 16

struct KbHWarpArgs {
    KbHashArgs h;                        // g.ent = partitioned elements; buckets via h.bstart / h.bend / h.bcap
    uint32_t* deferred;                  // [n_buckets] bucket ids left to the fallback kernel
    unsigned long long* n_deferred;
    uint32_t wbytes;                     // shared memory per warp (kb_hash_warp_wbytes)
};

// bytes of one warp's region: ring | queue | keys | presence words | column sets (only when they do not fit the key word) | mbarriers
static inline uint32_t kb_hash_warp_wbytes(uint32_t slots_log2, int pwn, bool packed) {
    const uint32_t S = 1u << slots_log2;
    const uint32_t b = 8u * (KB_HW_STAGES * KB_HW_CH + KB_HW_QCAP + S) + 4u * S * (uint32_t)pwn + (packed ? 0u : 8u * S) + 8u * KB_HW_STAGES;
    return (b + 15u) & ~15u;
}

template <bool D1, bool SPACER, int PWN>
__global__ void __launch_bounds__(32 * KB_HW_MAXWARPS, 1) kb_hash_warp_kernel(const KbHWarpArgs xs) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    const KbHashArgs& x = xs.h;
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    const uint32_t S = 1u << x.slots_log2, smask = S - 1u;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nw = blockDim.x >> 5;
    const uint32_t wg = blockIdx.x * nw + warp, TW = gridDim.x * nw;
    const bool packed = SPACER || (D1 && lo.FB <= 54);       // both 4-bit base sets inside the key word (bits 56-63)

    uint32_t smem_a = kb_smem_u32(kb_smem_raw) + warp * xs.wbytes;
    asm volatile("" : "+r"(smem_a));                          // opaque: keep the base in a register
    const uint32_t ring_a = smem_a;
    const uint32_t q_a = ring_a + KB_HW_STAGES * KB_HW_CH * 8;
    const uint32_t keys_a = q_a + KB_HW_QCAP * 8;
    const uint32_t pres_a = keys_a + S * 8;
    const uint32_t msk_a = pres_a + S * 4 * PWN;              // [S][2] (unpacked sets only)
    const uint32_t bars_a = msk_a + (packed ? 0u : S * 8);

    if (lane == 0) {
        for (int s = 0; s < KB_HW_STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bars_a + 8 * s), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = lane; i < S; i += 32) {
        kb_sts64(keys_a + i * 8, KB_KH_EMPTY);
#pragma unroll
        for (int j = 0; j < PWN; j++) kb_sts32(pres_a + (i * PWN + j) * 4, 0u);
        if (!packed) kb_sts64(msk_a + i * 8, 0ULL);
    }
    __syncwarp();

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
    uint32_t nkeys = 0;                  // distinct keys in the table (warp-uniform)
    bool over = false;                   // the current bucket is given up (warp-uniform)

    auto bucket_range = [&](uint32_t b, uint64_t& s, uint64_t& e) {
        if (x.bcap) { s = (uint64_t)b * x.bcap; e = min((uint64_t)x.bend[b], s + x.bcap); }
        else { s = x.bstart[b]; e = x.bend ? x.bend[b] : x.bstart[b + 1]; }
        if (e < s) e = s;
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
        nkeys += __popc(__ballot_sync(0xFFFFFFFFu, st == 1u));
        if (__any_sync(0xFFFFFFFFu, st == 2u) || nkeys > limit) over = true;
    };

    // ---- producer: the chunk sequence of this warp's buckets (wg, wg + TW, ...), bulk copies issued by lane 0 ----------------
    const uint64_t* ent = a.ent;
    uint32_t pb = wg, pk = 0, pnch = 0, pq = 0;
    uint64_t ps_al = 0, pn_al = 0;
    auto prod_enter = [&]() {            // geometry of bucket pb
        uint64_t s, e;
        bucket_range(pb, s, e);
        ps_al = s & ~1ULL;               // 16-byte aligned stream start
        pn_al = e > s ? e - ps_al : 0ULL;
        pnch = (uint32_t)((pn_al + KB_HW_CH - 1) / KB_HW_CH);
        pk = 0;
    };
    if (pb < x.n_buckets) prod_enter();
    auto issue = [&]() {
        while (pb < x.n_buckets && pk >= pnch) { pb += TW; if (pb < x.n_buckets) prod_enter(); }
        if (pb >= x.n_buckets) return;
        const uint32_t cnt = (uint32_t)min((uint64_t)KB_HW_CH, pn_al - (uint64_t)pk * KB_HW_CH);
        const uint32_t bytes = ((cnt + 1u) & ~1u) * 8u;
        const uint32_t st = pq % KB_HW_STAGES;
        if (lane == 0) {
            kb_mbar_expect_tx(bars_a + 8 * st, bytes);
            kb_bulk_g2s(ring_a + st * (KB_HW_CH * 8), ent + ps_al + (uint64_t)pk * KB_HW_CH, bytes, bars_a + 8 * st);
        }
        pk++; pq++;
    };
#pragma unroll 1
    for (int s = 0; s < KB_HW_STAGES; s++) issue();

    uint32_t cq = 0;
    uint32_t n_closed_t = 0, n_present_t = 0, n_rounds = 0, n_defer = 0;
    for (uint32_t b = wg; b < x.n_buckets; b += TW) {
        uint64_t bs, be;
        bucket_range(b, bs, be);
        if (be <= bs) continue;
        const uint64_t s_al = bs & ~1ULL;
        const uint32_t skip = (uint32_t)(bs - s_al);
        const uint64_t n_al = be - s_al;
        const uint32_t nch = (uint32_t)((n_al + KB_HW_CH - 1) / KB_HW_CH);
        over = false; nkeys = 0; qn = 0;
        for (uint32_t ck = 0; ck < nch; ck++) {
            const uint32_t st = cq % KB_HW_STAGES;
            kb_mbar_wait(bars_a + 8 * st, (cq / KB_HW_STAGES) & 1u);
            const uint32_t stage_a = ring_a + st * (KB_HW_CH * 8);
            const uint32_t cnt = (uint32_t)min((uint64_t)KB_HW_CH, n_al - (uint64_t)ck * KB_HW_CH);
            const uint32_t first = ck == 0 ? skip : 0u;
            if (!over) {
                uint64_t e[KB_HW_PER], k[KB_HW_PER];
                uint32_t sl[KB_HW_PER];
#pragma unroll
                for (int j = 0; j < KB_HW_PER; j++) e[j] = kb_lds64(stage_a + (j * 32 + lane) * 8);
#pragma unroll
                for (int j = 0; j < KB_HW_PER; j++) { sl[j] = slot_of(e[j]); k[j] = kb_lds64(keys_a + sl[j] * 8); }
                const bool whole = first == 0 && cnt == KB_HW_CH;
#pragma unroll
                for (int j = 0; j < KB_HW_PER; j++) {
                    const uint32_t idx = j * 32 + lane;
                    const bool act = whole || (idx >= first && idx < cnt);
                    const bool hit = act && (packed ? (k[j] & KB_HS_KEYMASK) : k[j]) == (e[j] >> kshift);
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, act && !hit);
                    if (m) {
                        if (act && !hit) kb_sts64(q_a + (qn + __popc(m & lt_mask)) * 8, e[j]);
                        qn += __popc(m);
                    }
                    if (hit) accumulate(sl[j], e[j], (uint32_t)(k[j] >> 32));
                    if (qn >= 32) { __syncwarp(); drain(); }
                }
            }
            __syncwarp();                // every lane is done with the stage: refill it with the chunk STAGES ahead
            cq++;
            issue();
        }
        // ---- bucket b is complete: flush the queue, evaluate, emit, clear ---------------------------------------------------
        __syncwarp();
        while (qn && !over) drain();
        __syncwarp();
        n_rounds++;
        uint32_t flags = 0, n_surv = 0;
        if (!over) {
            uint32_t n_closed = 0, n_present = 0;
            for (uint32_t q = 0, slot = lane; slot < S; q++, slot += 32) {
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
            if (n_surv <= KB_HW_MAX_INLINE) { n_closed_t += n_closed; n_present_t += n_present; }
        }
        const bool defer = over || n_surv > KB_HW_MAX_INLINE;
        for (uint32_t q = 0, slot = lane; slot < S; q++, slot += 32) {
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
        if (defer) {
            n_defer++;
            if (lane == 0) { const unsigned long long d = atomicAdd(xs.n_deferred, 1ULL); xs.deferred[d] = b; }
        }
        __syncwarp();
    }

    n_closed_t = __reduce_add_sync(0xFFFFFFFFu, n_closed_t);
    n_present_t = __reduce_add_sync(0xFFFFFFFFu, n_present_t);
    if (lane == 0) {
        if (n_closed_t) atomicAdd(a.stats + 0, (unsigned long long)n_closed_t);
        if (n_rounds - n_defer) atomicAdd(a.stats + 1, (unsigned long long)(n_rounds - n_defer));
        if (n_present_t) atomicAdd(a.stats + 2, (unsigned long long)n_present_t);
    }
}
