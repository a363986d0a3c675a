// kb_hash_stream.cuh — K3 fast path: persistent, TMA-fed version of the bucket hash aggregation.
//
// Same per-bucket algorithm and slot layout as kb_hash_fast_kernel (kb_hash.cuh; one-word records,
// <= 64 files, D <= 8), restructured so that the HBM stream never waits for the hash table:
//
//   * persistent CTAs: CTA c owns the buckets whose first record lies in [c n/G, (c+1) n/G) — a contiguous
//     slice of the partitioned array, found with two binary searches in the bucket table;
//   * one producer lane streams that slice into a ring of KB_HS_STAGES x 4 KB shared-memory stages with
//     bulk asynchronous copies (cp.async.bulk, completion on an mbarrier per stage), running ahead of the
//     consumers across bucket boundaries, so table scans and clears overlap with the next bucket's loads;
//   * 8 consumer warps take records from the ring (conflict-free LDS.64), probe / update the table, and at
//     every bucket end scan + clear the table and emit the survivors.
//
// Rare cases leave this kernel: a bucket whose distinct keys overflow the table, or with more than
// KB_HS_MAX_INLINE survivors, is appended to a deferred list and processed by kb_hash_fast_kernel (which can
// split buckets and counts group sizes in the table).  Survivors emitted here get their group size from
// kb_hsize_kernel (one warp per survivor re-reads its bucket, L2-resident on genome panels).
#pragma once
#include "kb_hash.cuh"

#define KB_HS_CONSUMERS 256
#define KB_HS_WARPS (KB_HS_CONSUMERS / 32)
#define KB_HS_THREADS KB_HS_CONSUMERS
#define KB_HS_CHUNK 1024                      // records per stage (8 KB)
#define KB_HS_PER (KB_HS_CHUNK / KB_HS_CONSUMERS)
#define KB_HS_STAGES 3
#define KB_HS_QCAP 64                         // per-warp queue of records that missed their home slot
#define KB_HS_MAX_INLINE 8
#define KB_HS_SLACK (KB_HS_CHUNK + 2)         // elements that must be readable past the end of the array

struct KbHStreamArgs {
    KbHashArgs h;
    const unsigned long long* n_ptr;     // elements in h.g.ent (device-resident: K1's record counter)
    uint32_t* deferred;                  // [n_buckets] bucket ids left to kb_hash_fast_kernel
    unsigned long long* n_deferred;
};

__device__ __forceinline__ uint32_t kb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(kb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kb_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool kb_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void kb_mbar_wait(uint32_t bar, uint32_t parity) { while (!kb_mbar_try(bar, parity)) {} }
__device__ __forceinline__ void kb_mbar_wait_backoff(uint32_t bar, uint32_t parity) { while (!kb_mbar_try(bar, parity)) __nanosleep(256); }
__device__ __forceinline__ void kb_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void kb_consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(KB_HS_CONSUMERS) : "memory"); }

// shared-memory accesses by 32-bit shared address (one instruction each, no generic-address arithmetic)
__device__ __forceinline__ uint64_t kb_lds64(uint32_t a) { uint64_t v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t kb_lds32(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void kb_sts64(uint32_t a, uint64_t v) { asm volatile("st.volatile.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void kb_sts32(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void kb_reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t kb_atoms_add(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
__device__ __forceinline__ uint64_t kb_atoms_cas64(uint32_t a, uint64_t cmp, uint64_t val) {
    uint64_t o; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(o) : "r"(a), "l"(cmp), "l"(val) : "memory"); return o;
}

// first index b in [0, nb] with bstart[b] >= v  (bstart has nb + 1 non-decreasing entries, bstart[nb] = n)
__device__ __forceinline__ uint32_t kb_lower_bound(const unsigned long long* bstart, uint32_t nb, unsigned long long v) {
    uint32_t lo = 0, hi = nb + 1;
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (bstart[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

// Slot = KbKhSlot (24 bytes: key | files 0-31 | files 32-63 | ingroup sets | outgroup sets).  With D == 1 and at most
// 54 flank bits (the spacer shape 25/1/2) the two 4-bit base sets live in bits 56-63 of the KEY word instead
// (ingroup 56-59, outgroup 60-63): the one LDS.64 that checks the key also tells whether the record's base is
// already recorded, so a record costs two table accesses (key load, presence RED) instead of three.
// PWN = 32-bit presence words (2 / 4 / 8: up to 64 / 128 / 256 files).  The slot size is kept at an odd number of 8-byte units
// (24 / 40 / 56 bytes) so that slot addresses spread over all shared-memory banks.
template <int PWN> struct KbHsSlotT { unsigned long long key; uint32_t pres[PWN]; uint32_t msk[2]; uint32_t pad[PWN == 2 ? 0 : 2]; };
template <> struct KbHsSlotT<2> { unsigned long long key; uint32_t pres[2]; uint32_t msk[2]; };
#define KB_HS_KEYMASK 0x00FFFFFFFFFFFFFFULL

// SPACER: the record layout is exactly 25/1/2-like (54 flank bits, D = 1): every shift is a compile-time constant.
template <bool D1, bool SPACER, int PWN>
__global__ void __launch_bounds__(KB_HS_THREADS) kb_hash_stream_kernel(const KbHStreamArgs xs) {
    typedef KbHsSlotT<PWN> KbHsSlot;
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    const KbHashArgs& x = xs.h;
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    const uint32_t S = 1u << x.slots_log2, smask = S - 1u;
    // dynamic shared memory: ring | per-warp queues | table | control words | mbarriers
    uint64_t* ring = reinterpret_cast<uint64_t*>(kb_smem_raw);                                  // STAGES * CHUNK
    uint64_t* queue = ring + KB_HS_STAGES * KB_HS_CHUNK;                                        // WARPS * QCAP
    KbHsSlot* tab = reinterpret_cast<KbHsSlot*>(queue + KB_HS_WARPS * KB_HS_QCAP);              // S
    uint32_t* s_ctl = reinterpret_cast<uint32_t*>(tab + S);   // per bucket parity p: [p] over, [2 + p] distinct keys, [4 + p] survivors; [8 + s] warps done with stage s
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ctl + 16);                                   // full[STAGES]
    __shared__ uint32_t s_b0, s_b1;
    __shared__ uint32_t s_closed, s_present, s_rounds, s_defer;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const unsigned long long n = *xs.n_ptr;
        const unsigned long long v0 = (unsigned long long)blockIdx.x * n / gridDim.x;
        const unsigned long long v1 = (blockIdx.x + 1 == gridDim.x) ? n : (unsigned long long)(blockIdx.x + 1) * n / gridDim.x;
        s_b0 = kb_lower_bound(x.bstart, x.n_buckets, v0);
        s_b1 = kb_lower_bound(x.bstart, x.n_buckets, v1);
        for (int s = 0; s < KB_HS_STAGES; s++) kb_mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 16; i++) s_ctl[i] = 0;
        s_closed = 0; s_present = 0; s_rounds = 0; s_defer = 0;
    }
    __syncthreads();
    const uint32_t b0 = s_b0, b1 = min(s_b1, x.n_buckets);
    if (b0 >= b1) return;
    const uint64_t r_begin = x.bstart[b0], r_end = x.bstart[b1];
    if (r_begin >= r_end) return;
    const uint64_t idx0 = r_begin & ~1ULL;                               // 16-byte aligned stream start
    const uint32_t nchunks = (uint32_t)((r_end - idx0 + KB_HS_CHUNK - 1) / KB_HS_CHUNK);
    uint32_t smem_a = kb_smem_u32(kb_smem_raw);
    asm volatile("" : "+r"(smem_a));                                     // opaque: keep the base in a register, never re-derive it
    const uint32_t ring_a = smem_a;
    const uint32_t tab_a = ring_a + (KB_HS_STAGES * KB_HS_CHUNK + KB_HS_WARPS * KB_HS_QCAP) * 8;
    const uint32_t ctl_a = tab_a + S * (uint32_t)sizeof(KbHsSlot);
    const uint32_t bars_a = ctl_a + 64;
    const uint64_t* src = a.ent + idx0;

    // the ring is primed by one thread; afterwards the LAST warp to finish a stage re-arms it (no producer warp, no polling)
    if (tid == 0) {
        for (uint32_t k = 0; k < min(nchunks, (uint32_t)KB_HS_STAGES); k++) {
            kb_mbar_expect_tx(bars_a + 8 * k, KB_HS_CHUNK * 8);
            kb_bulk_g2s(ring_a + k * (KB_HS_CHUNK * 8), src + (uint64_t)k * KB_HS_CHUNK, KB_HS_CHUNK * 8, bars_a + 8 * k);
        }
    }

    const uint32_t kshift = SPACER ? 10u : 64 - lo.FB;
    const uint32_t D2 = SPACER ? 2u : 2 * lo.D;
    const uint32_t mshift = SPACER ? 8u : 64 - lo.FB - D2;
    const uint32_t colmask = lo.D ? (0xFFFFFFFFu << (4 * (8 - lo.D))) : 0u;
    const uint32_t limit = S - (S >> 2);
    const uint32_t hmask = SPACER ? 0xFFFFFFFFu : kb_kh_hmask((uint32_t)lo.FB, x.bb);   // (SPACER: 54 - bb >= 32 key bits below the bucket bits)
    const uint32_t ing_lo = (uint32_t)x.ingroup64, ing_hi = (uint32_t)(x.ingroup64 >> 32);
    const uint32_t sshift = 32 - x.slots_log2;
    const uint32_t q_a = ring_a + (KB_HS_STAGES * KB_HS_CHUNK) * 8 + warp * (KB_HS_QCAP * 8);
    const uint32_t lt_mask = kb_lanemask_lt();
    uint32_t par = 0;                    // parity of the current bucket
    uint32_t qn = 0;                     // records waiting in this warp's queue (warp-uniform)

    const bool packed = SPACER || (D1 && lo.FB <= 54);       // base sets inside the key word
    // record -> its bits in the table slot (sa = shared address of the slot, khi = high half of the key word as last read)
    auto accumulate = [&](uint32_t sa, uint64_t e, uint32_t khi) {
        const uint32_t id = (uint32_t)e & 0xFFu;
        kb_reds_or(sa + 8 + ((id >> 5) << 2), 1u << (id & 31));
        if (D2) {
            const uint32_t isin = PWN == 2 ? ((((id & 32u) ? ing_hi : ing_lo) >> (id & 31)) & 1u) : ((a.ingroup[id >> 5] >> (id & 31)) & 1u);
            if (packed) {
                const uint32_t bit = (isin ? 0x01000000u : 0x10000000u) << ((uint32_t)(e >> mshift) & 3u);   // bit 24 + code (in) / 28 + code (out) of the high half
                if (!(khi & bit)) kb_reds_or(sa + 4, bit);
            } else {
                uint32_t oh;
                if (D1) oh = 0x10000000u << ((uint32_t)(e >> mshift) & 3u);
                else oh = kb_onehot8(((uint32_t)(e >> mshift) & ((1u << D2) - 1u)) << (16 - D2)) & colmask;
                const uint32_t ma = sa + 8 + 4 * PWN + 4 - 4 * isin;
                if ((kb_lds32(ma) & oh) != oh) kb_reds_or(ma, oh);
            }
        }
    };
    // full probe (dense: called with the queued records of a whole warp): find the key or claim an empty slot
    auto insert = [&](uint64_t e) {
        const uint64_t key = e >> kshift;
        uint32_t slot = (kb_kh_bits(e, x.bb, hmask) * 0x9E3779B1u) >> sshift;   // (one more multiply: the key bits below the bucket bits are only lightly mixed)
        for (uint32_t step = 0; step <= S; step++) {
            const uint32_t sa = tab_a + slot * (uint32_t)sizeof(KbHsSlot);
            uint64_t k = kb_lds64(sa);
            if (k == KB_KH_EMPTY) {
                k = kb_atoms_cas64(sa, KB_KH_EMPTY, key);
                if (k == KB_KH_EMPTY) {
                    if (kb_atoms_add(ctl_a + 4 * (2 + par), 1u) >= limit) kb_sts32(ctl_a + 4 * par, 1u);
                    k = key;
                }
            }
            if ((packed ? (k & KB_HS_KEYMASK) : k) == key) { accumulate(sa, e, (uint32_t)(k >> 32)); return; }
            slot = (slot + 1) & smask;
        }
        kb_sts32(ctl_a + 4 * par, 1u);
    };
    auto drain = [&]() {                 // up to 32 queued records, one per lane
        const uint32_t cnt = min(qn, 32u);
        qn -= cnt;
        if (lane < cnt) insert(kb_lds64(q_a + (qn + lane) * 8));
        __syncwarp();
    };
    // the common case inline: the key sits in its home slot; everything else goes through the queue
    auto handle = [&](uint64_t e, bool act) {
        const uint64_t key = e >> kshift;
        const uint32_t sa = tab_a + ((kb_kh_bits(e, x.bb, hmask) * 0x9E3779B1u) >> sshift) * (uint32_t)sizeof(KbHsSlot);
        const uint64_t k = kb_lds64(sa);
        const bool hit = act && (packed ? (k & KB_HS_KEYMASK) : k) == key;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, act && !hit);
        if (act && !hit) kb_sts64(q_a + (qn + __popc(m & lt_mask)) * 8, e);
        qn += __popc(m);
        if (hit) accumulate(sa, e, (uint32_t)(k >> 32));
        __syncwarp();
        if (qn >= 32) drain();
    };

    auto handle_full = [&](uint64_t e) {          // handle() for a record that is known to be in range
        const uint64_t key = e >> kshift;
        const uint32_t sa = tab_a + ((kb_kh_bits(e, x.bb, hmask) * 0x9E3779B1u) >> sshift) * (uint32_t)sizeof(KbHsSlot);
        const uint64_t k = kb_lds64(sa);
        const bool hit = (packed ? (k & KB_HS_KEYMASK) : k) == key;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, !hit);
        if (m) {
            if (!hit) kb_sts64(q_a + (qn + __popc(m & lt_mask)) * 8, e);
            qn += __popc(m);
        }
        if (hit) accumulate(sa, e, (uint32_t)(k >> 32));
        __syncwarp();
        if (qn >= 32) drain();
    };

    auto clear_slot = [&](uint32_t i) {
        tab[i].key = KB_KH_EMPTY; tab[i].msk[0] = 0; tab[i].msk[1] = 0;
#pragma unroll
        for (int j = 0; j < PWN; j++) tab[i].pres[j] = 0;
    };
    if (tid < KB_HS_CONSUMERS) for (uint32_t i = tid; i < S; i += KB_HS_CONSUMERS) clear_slot(i);
    __syncthreads();
    if (warp == KB_HS_WARPS) return;     // (spare warp of the launch shape; all work is done by the 8 consumer warps)

    uint32_t b = b0;
    uint64_t bs = r_begin, be = x.bstart[b + 1];
    while (be == bs && b + 1 < b1) { b++; be = x.bstart[b + 1]; }        // leading empty buckets
    uint64_t pos = r_begin;
    uint32_t n_closed_t = 0, n_present_t = 0;

    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t s = k % KB_HS_STAGES, u = k / KB_HS_STAGES;
        kb_mbar_wait(bars_a + 8 * s, u & 1u);
        const uint32_t stage_a = ring_a + s * (KB_HS_CHUNK * 8);
        const uint64_t c0 = idx0 + (uint64_t)k * KB_HS_CHUNK;
        const uint64_t c1 = min(c0 + KB_HS_CHUNK, r_end);
        while (pos < c1) {
            const uint64_t seg_end = min(c1, be);
            const uint32_t o0 = (uint32_t)(pos - c0), o1 = (uint32_t)(seg_end - c0);
            if (!__any_sync(0xFFFFFFFFu, kb_lds32(ctl_a + 4 * par))) {
                if (o0 == 0 && o1 == KB_HS_CHUNK) {                       // the whole stage belongs to the current bucket
                    uint64_t e[KB_HS_PER];
#pragma unroll
                    for (int j = 0; j < KB_HS_PER; j++) e[j] = kb_lds64(stage_a + (j * KB_HS_CONSUMERS + tid) * 8);
#pragma unroll
                    for (int j = 0; j < KB_HS_PER; j++) handle_full(e[j]);
                } else {
                    for (uint32_t j0 = o0 & ~31u; j0 < o1; j0 += KB_HS_CONSUMERS) {       // warp-uniform trip count
                        const uint32_t j = j0 + tid;
                        const bool act = j >= o0 && j < o1;
                        handle(act ? kb_lds64(stage_a + j * 8) : 0ULL, act);
                    }
                }
            }
            pos = seg_end;
            if (pos == be) {
                // ---- bucket b is complete: flush the queues, evaluate, emit, clear -------------------------------
                if (__any_sync(0xFFFFFFFFu, kb_lds32(ctl_a + 4 * par))) qn = 0;
                while (qn) drain();
                kb_consumer_sync();
                const bool over = s_ctl[par] != 0;
                uint32_t flags = 0, n_closed = 0, n_present = 0;
                if (!over) {
                    for (uint32_t q = 0, slot = tid; slot < S; q++, slot += KB_HS_CONSUMERS) {
                        const unsigned long long kk0 = tab[slot].key;
                        if (kk0 == KB_KH_EMPTY) continue;
                        n_closed++;
                        bool present = true;
#pragma unroll
                        for (int j = 0; j < PWN; j++) present = present && (tab[slot].pres[j] == a.full[j]);
                        if (!present) continue;
                        n_present++;
                        bool ok = true;
                        if (lo.D) {
                            const uint32_t m_in = packed ? ((uint32_t)(kk0 >> 56) & 0xFu) << 28 : tab[slot].msk[0];
                            const uint32_t m_out = packed ? ((uint32_t)(kk0 >> 60) & 0xFu) << 28 : tab[slot].msk[1];
                            uint32_t y = m_in & m_out;
                            y |= y >> 1; y |= y >> 2;
                            ok = (~y & 0x11111111u & colmask) != 0;
                        }
                        if (ok) flags |= 1u << q;
                    }
                    if (flags) atomicAdd(&s_ctl[4 + par], (uint32_t)__popc(flags));
                }
                kb_consumer_sync();
                const bool defer = over || s_ctl[4 + par] > KB_HS_MAX_INLINE;
                for (uint32_t q = 0, slot = tid; slot < S; q++, slot += KB_HS_CONSUMERS) {
                    const unsigned long long kk0 = tab[slot].key;
                    if (kk0 == KB_KH_EMPTY) continue;
                    if (!defer && ((flags >> q) & 1u)) {
                        const unsigned long long gs = atomicAdd(a.n_res, 1ULL);
                        if (gs < a.cap) {
                            uint64_t kk = packed ? (kk0 & KB_HS_KEYMASK) : kk0;
                            if (lo.mix) kk = kb_unmix(kk, lo.FB, lo.shs);
                            a.res_flank[gs] = kk << kshift;
                            if (lo.MW) {
                                a.res_in[gs] = packed ? ((uint32_t)(kk0 >> 56) & 0xFu) << 28 : tab[slot].msk[0];
                                a.res_out[gs] = packed ? ((uint32_t)(kk0 >> 60) & 0xFu) << 28 : tab[slot].msk[1];
                            }
                            a.res_run[2 * gs] = bs;
                            a.res_run[2 * gs + 1] = be - bs;
                            a.res_size[gs] = 0xFFFFFFFFu;                    // filled by kb_hsize_kernel
                        }
                    }
                    clear_slot(slot);
                }
                if (!defer) { n_closed_t += n_closed; n_present_t += n_present; }
                if (tid == 0) {
                    if (defer) { const unsigned long long d = atomicAdd(xs.n_deferred, 1ULL); xs.deferred[d] = b; s_defer++; }
                    s_ctl[par ^ 1u] = 0; s_ctl[2 + (par ^ 1u)] = 0; s_ctl[4 + (par ^ 1u)] = 0; s_rounds++;   // the next bucket's words
                }
                kb_consumer_sync();
                par ^= 1u;
                // next non-empty bucket
                bs = be;
                if (b + 1 < b1) {
                    do { b++; be = x.bstart[b + 1]; } while (be == bs && b + 1 < b1);
                }
            }
        }
        // ---- release the stage: the last of the 8 warps re-arms it with the chunk STAGES ahead ---------------------
        __syncwarp();
        if (lane == 0) {
            if (kb_atoms_add(ctl_a + 4 * (8 + s), 1u) == KB_HS_WARPS - 1) {
                kb_sts32(ctl_a + 4 * (8 + s), 0u);
                const uint32_t kn = k + KB_HS_STAGES;
                if (kn < nchunks) {
                    kb_mbar_expect_tx(bars_a + 8 * s, KB_HS_CHUNK * 8);
                    kb_bulk_g2s(stage_a, src + (uint64_t)kn * KB_HS_CHUNK, KB_HS_CHUNK * 8, bars_a + 8 * s);
                }
            }
        }
    }

    n_closed_t = __reduce_add_sync(0xFFFFFFFFu, n_closed_t);
    n_present_t = __reduce_add_sync(0xFFFFFFFFu, n_present_t);
    if (lane == 0) { if (n_closed_t) atomicAdd(&s_closed, n_closed_t); if (n_present_t) atomicAdd(&s_present, n_present_t); }
    kb_consumer_sync();
    if (tid == 0) {
        if (s_closed) atomicAdd(a.stats + 0, (unsigned long long)s_closed);
        if (s_rounds - s_defer) atomicAdd(a.stats + 1, (unsigned long long)(s_rounds - s_defer));
        if (s_present) atomicAdd(a.stats + 2, (unsigned long long)s_present);
    }
}

static inline size_t kb_hash_stream_smem(uint32_t slots_log2, int pwn) {
    const size_t slot = pwn == 2 ? sizeof(KbHsSlotT<2>) : (pwn == 4 ? sizeof(KbHsSlotT<4>) : sizeof(KbHsSlotT<8>));
    return (size_t)KB_HS_STAGES * KB_HS_CHUNK * 8 + (size_t)KB_HS_WARPS * KB_HS_QCAP * 8 + ((size_t)1 << slots_log2) * slot + 64 + 8 * KB_HS_STAGES + 16;
}
static_assert(sizeof(KbHsSlotT<2>) == 24 && sizeof(KbHsSlotT<4>) == 40 && sizeof(KbHsSlotT<8>) == 56, "odd multiples of 8 bytes");

// ---- group sizes of the survivors emitted by the stream kernel: one warp per survivor ----------------------
struct KbHSizeArgs {
    const uint64_t* ent;
    const unsigned long long* n_res;
    uint64_t cap;
    const uint64_t* res_flank;
    const uint64_t* res_run;
    uint32_t* res_size;
    KbLayout lo;
};

__global__ void __launch_bounds__(256) kb_hsize_kernel(const KbHSizeArgs a) {
    const KbLayout& lo = a.lo;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n = min((uint64_t)*a.n_res, a.cap);
    const uint32_t kshift = 64 - lo.FB;
    for (uint64_t g = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); g < n; g += (uint64_t)gridDim.x * 8) {
        if (a.res_size[g] != 0xFFFFFFFFu) continue;
        uint64_t key = a.res_flank[g] >> kshift;
        if (lo.mix) key = kb_mix(key, lo.FB, lo.shs);
        const uint64_t start = a.res_run[2 * g], len = a.res_run[2 * g + 1];
        uint32_t c = 0;
        for (uint64_t i = lane; i < len; i += 32) c += ((a.ent[start + i] >> kshift) == key) ? 1u : 0u;
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) a.res_size[g] = c;
    }
}
