// kb_hash.cuh — K3 (search path): per-bucket hash aggregation = intersection + diagnostic filter + compaction.
//
// Same contract as kb_group.cuh (the reference stages it replaces are listed there: simplifyStream /
// alignmentStream shared.py:210-240,:442-475, intersectSortedStreams :321-347 folded by mergeFiles
// intersectAmplicons.py:232-310, filterAlignments.py:4-28 + ingroupUniqueColumns Amplicon.py:495-521,
// consensus :550-558), but on PARTITIONED instead of sorted elements (kb_part.cuh): all records of one
// (left,right) key sit in the same bucket of a few thousand records, in no particular order.
//
// One CTA per bucket: the bucket's records are streamed ONCE (coalesced) through a shared-memory hash
// table keyed by the COMPLETE flank key — so grouping is exact whatever the bucket bits are — whose
// slots accumulate, with shared-memory atomics, the file-presence bitmap and the per-column ingroup /
// outgroup base sets.  Then the table is scanned: a key survives iff it is present in every file (S6)
// and, when D > 0, some column has disjoint ingroup / outgroup base sets (S7); survivors are appended to
// the result table.  Buckets with at least one survivor are streamed a second time (from L2) to count
// the survivors' records (group size, needed to size the --out_align gather).
//
// A bucket with more distinct keys than the table holds is split by further hash bits and streamed
// once per part (explicit work stack; never happens at the default bucket size on genome panels).
#pragma once
#include "kb_group.cuh"

#define KB_KH_THREADS 256
#define KB_KH_STACK 48
#define KB_KH_EMPTY 0xFFFFFFFFFFFFFFFFULL     // DIRECT keys have at most 56 bits
#define KB_KH_NONE 0xFFFFFFFFu
#define KB_KH_MAXSPLIT 20

struct KbHashArgs {
    KbGroupArgs g;                       // ent = partitioned elements; results; stats
    const unsigned long long* bstart;    // [n_buckets + 1]
    uint32_t n_buckets;
    uint32_t bb;                         // bucket bits: every element of a bucket has the same top bb bits
    uint32_t slots_log2;
    uint64_t ingroup64, full64;          // fast kernel (<= 64 files)
    unsigned long long* err;             // != 0: a bucket could not be resolved
    const uint32_t* list;                // != null: process only the buckets list[0 .. *n_list)
    const unsigned long long* n_list;
    uint32_t abort_above;                // list mode: more deferred buckets than this = the plan was too coarse for this input: set *err = 2
                                         // and leave (the host re-plans with more bucket bits instead of splitting every bucket here)
    const unsigned long long* brun;      // != null (generic kernel): bucket b = elements [brun[2b], brun[2b] + brun[2b+1]) (kb_prefilter.cuh)
    const unsigned long long* bend;      // != null: bucket b ends at bend[b] instead of bstart[b + 1]
    uint64_t bcap;                       // != 0 (slab layout, kb_extract_part.cuh): bucket b = [b * bcap, min(bend[b], (b + 1) * bcap)), bstart unused
    uint32_t bucket0;                    // kb_hash_warp_kernel: this launch covers buckets [bucket0, bucket0 + n_buckets)
};

// element range of bucket b in every layout
__device__ __forceinline__ void kb_bucket_range(const KbHashArgs& x, uint32_t b, uint64_t& bs, uint64_t& be) {
    if (x.brun) { bs = x.brun[2 * (size_t)b]; be = bs + x.brun[2 * (size_t)b + 1]; }
    else if (x.bcap) { bs = (uint64_t)b * x.bcap; be = min((uint64_t)x.bend[b], bs + x.bcap); }
    else { bs = x.bstart[b]; be = x.bend ? x.bend[b] : x.bstart[b + 1]; }
    if (be < bs) be = bs;
}

// The 32 key (or flank-hash) bits right below the bucket bits, left-aligned.  Mixed keys (kb_mix) and flank
// hashes are uniform there, so these bits index the table directly: the first `nb` bits select the part of a
// split bucket, the next slots_log2 bits the home slot.
__device__ __forceinline__ uint32_t kb_kh_bits(uint64_t e, uint32_t bb, uint32_t hmask) { return (uint32_t)((e << bb) >> 32) & hmask; }
__device__ __forceinline__ uint32_t kb_kh_hmask(uint32_t keybits, uint32_t bb) {
    const uint32_t rem = keybits - bb;
    return rem >= 32 ? 0xFFFFFFFFu : (rem == 0 ? 0u : (0xFFFFFFFFu << (32 - rem)));
}
__device__ __forceinline__ uint32_t kb_kh_part(uint32_t hh, uint32_t nb) { return nb ? (hh >> (32 - nb)) : 0u; }
__device__ __forceinline__ uint32_t kb_kh_slot(uint32_t hh, uint32_t nb, uint32_t slots_log2) { return (hh << nb) >> (32 - slots_log2); }

__device__ __forceinline__ uint32_t kb_ld_shared_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// work-stack bookkeeping shared by both kernels (thread 0 between barriers)
struct KbKhCtl {
    uint32_t sp, over, nkeys, nsurv;
    uint32_t stack[KB_KH_STACK];         // part (24 bits) | nbits << 24
};

// ===== fast kernel: one-word records, <= 64 files, D <= 8 ================================================
// slot (24 bytes): key u64 | files 0-31 | files 32-63 | ingroup column sets | outgroup column sets
struct KbKhSlot { unsigned long long key; uint32_t pres[2]; uint32_t msk[2]; };

template <bool D1>
__global__ void __launch_bounds__(KB_KH_THREADS) kb_hash_fast_kernel(const KbHashArgs x) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    const uint32_t S = 1u << x.slots_log2, smask = S - 1u;
    KbKhSlot* tab = reinterpret_cast<KbKhSlot*>(kb_smem_raw);
    __shared__ KbKhCtl ctl;
    __shared__ uint32_t s_closed, s_present, s_rounds, s_splits;

    const uint32_t tid = threadIdx.x;
    const uint32_t kshift = 64 - lo.FB;
    const uint32_t D2 = 2 * lo.D;
    const uint32_t mshift = 64 - lo.FB - D2;
    const uint32_t colmask = lo.D ? (0xFFFFFFFFu << (4 * (8 - lo.D))) : 0u;
    const uint32_t limit = S - (S >> 2);
    const uint32_t hmask = kb_kh_hmask((uint32_t)lo.FB, x.bb);
    const uint32_t ing_lo = (uint32_t)x.ingroup64, ing_hi = (uint32_t)(x.ingroup64 >> 32);
    if (tid == 0) { s_closed = 0; s_present = 0; s_rounds = 0; s_splits = 0; }

    // rest of the probe sequence after the home slot missed: find the key or claim an empty slot
    auto probe = [&](uint64_t key, uint32_t slot, bool insert) -> uint32_t {
        for (uint32_t step = 0; step <= S; step++) {
            const unsigned long long k = tab[slot].key;
            if (k == key) return slot;
            if (k == KB_KH_EMPTY) {
                if (!insert) return KB_KH_NONE;
                const unsigned long long old = atomicCAS(&tab[slot].key, KB_KH_EMPTY, (unsigned long long)key);
                if (old == KB_KH_EMPTY) {
                    if (atomicAdd(&ctl.nkeys, 1u) >= limit) ctl.over = 1;
                    return slot;
                }
                if (old == key) return slot;
            }
            slot = (slot + 1) & smask;
        }
        ctl.over = 1;
        return KB_KH_NONE;
    };

    if (x.list && x.abort_above && *x.n_list > (unsigned long long)x.abort_above) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(x.err, 0ULL, 2ULL);
        return;
    }
    const uint32_t n_work = x.list ? (uint32_t)min((unsigned long long)x.n_buckets, *x.n_list) : x.n_buckets;
    for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const uint32_t b = x.list ? x.list[wi] : wi;
        uint64_t bs, be;
        kb_bucket_range(x, b, bs, be);
        if (be == bs) continue;
        __syncthreads();
        if (tid == 0) { ctl.sp = 1; ctl.stack[0] = 0; }
        __syncthreads();
        while (true) {
            const uint32_t sp = ctl.sp;
            if (sp == 0) break;
            const uint32_t item = ctl.stack[sp - 1];
            __syncthreads();
            const uint32_t part = item & 0xFFFFFFu, nb = item >> 24;
            if (tid == 0) { ctl.sp = sp - 1; ctl.over = 0; ctl.nkeys = 0; ctl.nsurv = 0; s_rounds++; }
            for (uint32_t i = tid; i < S; i += KB_KH_THREADS) { tab[i].key = KB_KH_EMPTY; tab[i].pres[0] = 0; tab[i].pres[1] = 0; tab[i].msk[0] = 0; tab[i].msk[1] = 0; }
            __syncthreads();

            // ---- stream the bucket through the table (4 coalesced loads in flight per thread) ---------------
            for (uint64_t i0 = bs; i0 < be; i0 += 4 * KB_KH_THREADS) {
                const uint32_t left = (uint32_t)min((uint64_t)(4 * KB_KH_THREADS), be - i0);
                uint64_t r[4];
#pragma unroll
                for (int u = 0; u < 4; u++) r[u] = (u * KB_KH_THREADS + tid < left) ? kb_ld_stream(a.ent + i0 + u * KB_KH_THREADS + tid) : 0ULL;
                if (__any_sync(0xFFFFFFFFu, kb_ld_shared_volatile(&ctl.over))) break;      // (warp-uniform exit: the loop body has __syncwarp)
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint64_t e = r[u];
                    const uint64_t key = e >> kshift;
                    const uint32_t hh = kb_kh_bits(e, x.bb, hmask);
                    const bool act = (u * KB_KH_THREADS + tid < left) && kb_kh_part(hh, nb) == part;
                    uint32_t slot = kb_kh_slot(hh, nb, x.slots_log2);
                    bool found = false;
                    if (act) found = tab[slot].key == key;                   // the common case: the key sits in its home slot
                    if (act && !found) { slot = probe(key, slot, true); found = slot != KB_KH_NONE; }
                    __syncwarp();                                            // reconverge before the accumulation
                    if (found) {
                        const uint32_t id = (uint32_t)e & 0xFFu;
                        atomicOr(&tab[slot].pres[id >> 5], 1u << (id & 31));
                        if (D2) {
                            uint32_t oh;
                            if (D1) oh = 0x10000000u << ((uint32_t)(e >> mshift) & 3u);
                            else oh = kb_onehot8(((uint32_t)(e >> mshift) & ((1u << D2) - 1u)) << (16 - D2)) & colmask;
                            const uint32_t isin = (((id & 32u) ? ing_hi : ing_lo) >> (id & 31)) & 1u;
                            uint32_t* p = &tab[slot].msk[isin ^ 1u];
                            if ((kb_ld_shared_volatile(p) & oh) != oh) atomicOr(p, oh);
                        }
                    }
                }
            }
            __syncthreads();
            if (ctl.over) {                                   // too many distinct keys: split this part by one more bit
                __syncthreads();
                if (tid == 0) {
                    s_splits++;
                    if (nb >= KB_KH_MAXSPLIT || ctl.sp + 2 > KB_KH_STACK) atomicExch(x.err, 1ULL);
                    else { ctl.stack[ctl.sp] = (part << 1) | ((nb + 1) << 24); ctl.stack[ctl.sp + 1] = ((part << 1) | 1u) | ((nb + 1) << 24); ctl.sp += 2; }
                }
                __syncthreads();
                continue;
            }

            // ---- scan the table: S6 / S7, emit survivors; afterwards pres[0] = record count, pres[1] = mark ----
            uint32_t n_closed = 0, n_present = 0, any = 0;
            for (uint32_t slot = tid; slot < S; slot += KB_KH_THREADS) {
                const unsigned long long k = tab[slot].key;
                if (k == KB_KH_EMPTY) continue;
                uint32_t mark = 0;
                n_closed++;
                const uint64_t P = (uint64_t)tab[slot].pres[0] | ((uint64_t)tab[slot].pres[1] << 32);
                if (P == x.full64) {
                    n_present++;
                    const uint32_t in = tab[slot].msk[0], out = tab[slot].msk[1];
                    bool ok = true;
                    if (lo.D) {
                        uint32_t y = in & out;
                        y |= y >> 1; y |= y >> 2;
                        ok = (~y & 0x11111111u & colmask) != 0;
                    }
                    if (ok) {
                        const unsigned long long gs = atomicAdd(a.n_res, 1ULL);
                        any = 1;
                        if (gs < a.cap) {
                            uint64_t kk = k;
                            if (lo.mix) kk = kb_unmix(kk, lo.FB, lo.shs);
                            a.res_flank[gs] = kk << kshift;
                            if (lo.MW) { a.res_in[gs] = in; a.res_out[gs] = out; }
                            a.res_run[2 * gs] = bs;
                            a.res_run[2 * gs + 1] = be - bs;
                            mark = (uint32_t)gs + 1u;
                        }
                    }
                }
                tab[slot].pres[0] = 0; tab[slot].pres[1] = mark;
            }
            if (any) ctl.nsurv = 1;
            if (n_closed) atomicAdd(&s_closed, n_closed);
            if (n_present) atomicAdd(&s_present, n_present);
            __syncthreads();

            // ---- survivors' group sizes: second stream of the bucket (L2) ---------------------------------------
            if (ctl.nsurv) {
                for (uint64_t i = bs + tid; i < be; i += KB_KH_THREADS) {
                    const uint64_t e = a.ent[i];
                    const uint64_t key = e >> kshift;
                    const uint32_t hh = kb_kh_bits(e, x.bb, hmask);
                    if (kb_kh_part(hh, nb) != part) continue;
                    const uint32_t slot = probe(key, kb_kh_slot(hh, nb, x.slots_log2), false);
                    if (slot != KB_KH_NONE && tab[slot].pres[1]) atomicAdd(&tab[slot].pres[0], 1u);
                }
                __syncthreads();
                for (uint32_t slot = tid; slot < S; slot += KB_KH_THREADS) {
                    const uint32_t mark = tab[slot].key != KB_KH_EMPTY ? tab[slot].pres[1] : 0u;
                    if (mark) a.res_size[mark - 1] = tab[slot].pres[0];
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (s_closed) atomicAdd(a.stats + 0, (unsigned long long)s_closed);
        if (s_rounds) atomicAdd(a.stats + 1, (unsigned long long)s_rounds);
        if (s_present) atomicAdd(a.stats + 2, (unsigned long long)s_present);
        if (s_splits) atomicAdd(a.stats + 3, (unsigned long long)s_splits);
    }
}

static inline size_t kb_hash_fast_smem(uint32_t slots_log2) { return ((size_t)1 << slots_log2) * sizeof(KbKhSlot) + 16; }

// ===== generic kernel: any record width, up to 256 files, D <= 128 ====================================
// Slot = tag (0 empty, 1 being written, else 0x80000000 | key bits) + FW key words + PW presence words +
// MW ingroup + MW outgroup mask words + 2 words (survivor's record count, mark).  A slot is claimed by CAS
// on the tag; the winner writes the key words and publishes the tag inside the same loop iteration, readers
// retry while the tag says "being written" (no thread ever waits inside a critical section).
#define KB_KH_TAG_EMPTY 0u
#define KB_KH_TAG_BUSY 1u

static inline size_t kb_hash_slot_bytes(const KbLayout& lo) { return 4 + 8 + 8 * (size_t)lo.FW + 4 * (size_t)lo.PW + 8 * (size_t)lo.MW; }
static inline size_t kb_hash_smem(const KbLayout& lo, uint32_t slots_log2) { return ((size_t)1 << slots_log2) * kb_hash_slot_bytes(lo) + 16; }

template <int WN>
__global__ void __launch_bounds__(KB_KH_THREADS) kb_hash_kernel(const KbHashArgs x) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    const KbGroupArgs& a = x.g;
    const KbLayout& lo = a.lo;
    const uint32_t S = 1u << x.slots_log2, smask = S - 1u;
    const int FW = lo.FW, PW = lo.PW, MW = lo.MW;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(kb_smem_raw);    // S * FW
    uint32_t* tags = reinterpret_cast<uint32_t*>(keys + (size_t)S * FW);               // S
    uint32_t* aux = tags + S;                                                           // 2S: survivor's record count, mark (result row + 1)
    uint32_t* pres = aux + 2 * S;                                                       // S * PW
    uint32_t* min_ = pres + (size_t)S * PW;                                             // S * MW
    uint32_t* mout = min_ + (size_t)S * MW;                                             // S * MW
    __shared__ KbKhCtl ctl;
    __shared__ uint32_t s_closed, s_present, s_rounds, s_splits;

    const uint32_t tid = threadIdx.x;
    const uint32_t limit = S - (S >> 2);
    const uint32_t hmask = kb_kh_hmask(WN == 1 ? (uint32_t)lo.FB : 32u, x.bb);
    if (tid == 0) { s_closed = 0; s_present = 0; s_rounds = 0; s_splits = 0; }

    auto key_matches = [&](uint32_t slot, const KbKey<WN>& key) -> bool {
        bool eq = true;
#pragma unroll
        for (int j = 0; j < WN; j++) if (j < FW) eq = eq && (keys[(size_t)slot * FW + j] == key.w[j]);
        return eq;
    };
    auto find_or_insert = [&](const KbKey<WN>& key, uint32_t hh, uint32_t slot, bool insert) -> uint32_t {
        const uint32_t want = 0x80000000u | hh;
        uint32_t step = 0;
        while (step <= S) {
            const uint32_t t = kb_ld_shared_volatile(&tags[slot]);
            if (t == KB_KH_TAG_EMPTY) {
                if (!insert) return KB_KH_NONE;
                if (atomicCAS(&tags[slot], KB_KH_TAG_EMPTY, KB_KH_TAG_BUSY) == KB_KH_TAG_EMPTY) {
#pragma unroll
                    for (int j = 0; j < WN; j++) if (j < FW) keys[(size_t)slot * FW + j] = key.w[j];
                    __threadfence_block();
                    *reinterpret_cast<volatile uint32_t*>(&tags[slot]) = want;
                    if (atomicAdd(&ctl.nkeys, 1u) >= limit) ctl.over = 1;
                    return slot;
                }
                continue;                                  // lost the race: look at the same slot again
            }
            if (t == KB_KH_TAG_BUSY) continue;             // being written by another thread
            if (t == want) { __threadfence_block(); if (key_matches(slot, key)) return slot; }
            slot = (slot + 1) & smask; step++;
        }
        ctl.over = 1;
        return KB_KH_NONE;
    };

    if (x.list && x.abort_above && *x.n_list > (unsigned long long)x.abort_above) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(x.err, 0ULL, 2ULL);
        return;
    }
    const uint32_t n_work = x.list ? (uint32_t)min((unsigned long long)x.n_buckets, *x.n_list) : x.n_buckets;
    for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const uint32_t b = x.list ? x.list[wi] : wi;
        uint64_t bs, be;
        kb_bucket_range(x, b, bs, be);
        if (be == bs) continue;
        __syncthreads();
        if (tid == 0) { ctl.sp = 1; ctl.stack[0] = 0; }
        __syncthreads();
        while (true) {
            const uint32_t sp = ctl.sp;
            if (sp == 0) break;
            const uint32_t item = ctl.stack[sp - 1];
            __syncthreads();
            const uint32_t part = item & 0xFFFFFFu, nb = item >> 24;
            if (tid == 0) { ctl.sp = sp - 1; ctl.over = 0; ctl.nkeys = 0; ctl.nsurv = 0; s_rounds++; }
            for (uint32_t i = tid; i < S; i += KB_KH_THREADS) tags[i] = KB_KH_TAG_EMPTY;
            for (uint32_t i = tid; i < S * (uint32_t)(2 + PW + 2 * MW); i += KB_KH_THREADS) aux[i] = 0;   // aux, pres, min_, mout are contiguous
            __syncthreads();

            // software pipeline: the element and the (random, 8 W bytes) record of the NEXT round are in flight while the
            // current one goes through the table
            auto load = [&](uint64_t i, uint64_t& e, uint32_t& hh, bool& act, uint64_t (&rec)[WN]) {
                act = i < be;
                e = 0; hh = 0;
#pragma unroll
                for (int j = 0; j < WN; j++) rec[j] = 0;
                if (act) { e = kb_ld_stream(a.ent + i); hh = kb_kh_bits(e, x.bb, hmask); act = kb_kh_part(hh, nb) == part; }
                if (act) kb_fetch_rec<WN>(a, e, rec);
            };
            uint64_t e_n, rec_n[WN]; uint32_t hh_n; bool act_n;
            load(bs + tid, e_n, hh_n, act_n, rec_n);
            for (uint64_t i0 = bs; i0 < be; i0 += KB_KH_THREADS) {
                const uint32_t hh = hh_n; const bool act = act_n;
                uint64_t rec[WN];
#pragma unroll
                for (int j = 0; j < WN; j++) rec[j] = rec_n[j];
                load(i0 + KB_KH_THREADS + tid, e_n, hh_n, act_n, rec_n);
                if (__any_sync(0xFFFFFFFFu, kb_ld_shared_volatile(&ctl.over))) break;      // (warp-uniform exit: the loop body has __syncwarp)
                KbKey<WN> key;
                kb_key_of<WN>(lo, rec, key);
                uint32_t slot = KB_KH_NONE;
                if (act) slot = find_or_insert(key, hh, kb_kh_slot(hh, nb, x.slots_log2), true);
                __syncwarp();
                if (slot != KB_KH_NONE) {
                    const uint32_t id = (uint32_t)rec[WN - 1] & 0xFFu;
                    atomicOr(&pres[(size_t)slot * PW + (id >> 5)], 1u << (id & 31));
                    const bool isin = (a.ingroup[id >> 5] >> (id & 31)) & 1u;
                    uint32_t* m = (isin ? min_ : mout) + (size_t)slot * MW;
                    for (int j = 0; j < MW; j++) {
                        const int ncol = min(8, lo.D - 8 * j);
                        const uint32_t v = (uint32_t)kb_rec_bits<WN>(rec, lo.FB + 16 * j, 2 * ncol) << (16 - 2 * ncol);
                        const uint32_t oh = kb_onehot8(v) & (0xFFFFFFFFu << (4 * (8 - ncol)));
                        if ((kb_ld_shared_volatile(m + j) & oh) != oh) atomicOr(m + j, oh);
                    }
                }
            }
            __syncthreads();
            if (ctl.over) {
                __syncthreads();
                if (tid == 0) {
                    s_splits++;
                    if (nb >= KB_KH_MAXSPLIT || ctl.sp + 2 > KB_KH_STACK) atomicExch(x.err, 1ULL);
                    else { ctl.stack[ctl.sp] = (part << 1) | ((nb + 1) << 24); ctl.stack[ctl.sp + 1] = ((part << 1) | 1u) | ((nb + 1) << 24); ctl.sp += 2; }
                }
                __syncthreads();
                continue;
            }

            // ---- scan: S6 / S7, emit; aux[2 slot + 1] = mark of a survivor ------------------------------------------
            uint32_t n_closed = 0, n_present = 0, any = 0;
            for (uint32_t slot = tid; slot < S; slot += KB_KH_THREADS) {
                const uint32_t t = tags[slot];
                if (t == KB_KH_TAG_EMPTY) continue;
                n_closed++;
                uint32_t mark = 0;
                bool present = true;
                for (int j = 0; j < PW; j++) present = present && (pres[(size_t)slot * PW + j] == a.full[j]);
                if (present) {
                    n_present++;
                    bool diag = (lo.D == 0);
                    for (int j = 0; j < MW; j++) {
                        const int ncol = min(8, lo.D - 8 * j);
                        uint32_t y = min_[(size_t)slot * MW + j] & mout[(size_t)slot * MW + j];
                        y |= y >> 1; y |= y >> 2;
                        diag = diag || ((~y & 0x11111111u & (0xFFFFFFFFu << (4 * (8 - ncol)))) != 0);
                    }
                    if (diag) {
                        const unsigned long long gs = atomicAdd(a.n_res, 1ULL);
                        any = 1;
                        if (gs < a.cap) {
                            if constexpr (WN == 1) {
                                uint64_t kk = keys[slot];
                                if (lo.FB) { if (lo.mix) kk = kb_unmix(kk, lo.FB, lo.shs); kk <<= (64 - lo.FB); }
                                a.res_flank[gs] = kk;
                            } else {
                                for (int j = 0; j < FW; j++) a.res_flank[gs * FW + j] = keys[(size_t)slot * FW + j];
                            }
                            for (int j = 0; j < MW; j++) { a.res_in[gs * MW + j] = min_[(size_t)slot * MW + j]; a.res_out[gs * MW + j] = mout[(size_t)slot * MW + j]; }
                            a.res_run[2 * gs] = bs;
                            a.res_run[2 * gs + 1] = be - bs;
                            mark = (uint32_t)gs + 1u;
                        }
                    }
                }
                aux[2 * slot + 1] = mark;
            }
            if (any) ctl.nsurv = 1;
            if (n_closed) atomicAdd(&s_closed, n_closed);
            if (n_present) atomicAdd(&s_present, n_present);
            __syncthreads();

            if (ctl.nsurv) {
                for (uint64_t i = bs + tid; i < be; i += KB_KH_THREADS) {
                    const uint64_t e = a.ent[i];
                    const uint32_t hh = kb_kh_bits(e, x.bb, hmask);
                    if (kb_kh_part(hh, nb) != part) continue;
                    uint64_t rec[WN]; KbKey<WN> key;
                    kb_fetch<WN>(a, e, rec, key);
                    const uint32_t slot = find_or_insert(key, hh, kb_kh_slot(hh, nb, x.slots_log2), false);
                    if (slot != KB_KH_NONE && aux[2 * slot + 1]) atomicAdd(&aux[2 * slot], 1u);
                }
                __syncthreads();
                for (uint32_t slot = tid; slot < S; slot += KB_KH_THREADS) {
                    const uint32_t mark = aux[2 * slot + 1];
                    if (mark) a.res_size[mark - 1] = aux[2 * slot];
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (s_closed) atomicAdd(a.stats + 0, (unsigned long long)s_closed);
        if (s_rounds) atomicAdd(a.stats + 1, (unsigned long long)s_rounds);
        if (s_present) atomicAdd(a.stats + 2, (unsigned long long)s_present);
        if (s_splits) atomicAdd(a.stats + 3, (unsigned long long)s_splits);
    }
}

// ---- records of the surviving groups (for --out_align): one warp per survivor scans its bucket and keeps
//      the records whose complete flank key matches ---------------------------------------------------------
struct KbHGatherArgs {
    const uint64_t* ent;
    const uint64_t* recs;
    const uint64_t* res_run;     // [n_groups][2] bucket start, bucket length
    const uint64_t* res_flank;   // [n_groups][FW]
    const uint64_t* off;         // [n_groups] exclusive prefix of group sizes
    uint64_t n_groups;
    uint64_t* out;
    KbLayout lo;
};

template <int WN>
__global__ void __launch_bounds__(256) kb_hgather_kernel(const KbHGatherArgs a) {
    const KbLayout& lo = a.lo;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= a.n_groups) return;
    const uint64_t start = a.res_run[2 * g], len = a.res_run[2 * g + 1];
    uint64_t o = a.off[g];
    uint64_t want[WN];
#pragma unroll
    for (int j = 0; j < WN; j++) want[j] = j < lo.FW ? a.res_flank[g * lo.FW + j] : 0ULL;
    KbGroupArgs ga{};
    ga.recs = a.recs; ga.lo = lo;
    for (uint64_t i0 = 0; i0 < len; i0 += 32) {
        const uint64_t i = i0 + lane;
        bool match = false;
        uint64_t rec[WN];
#pragma unroll
        for (int j = 0; j < WN; j++) rec[j] = 0;
        if (i < len) {
            const uint64_t e = a.ent[start + i];
            if constexpr (WN == 1) {
                uint64_t v = e;
                if (lo.FB && lo.mix) {
                    const uint64_t low = lo.FB < 64 ? (v & kb_lowmask(64 - lo.FB)) : 0ULL;
                    v = (kb_unmix(v >> (64 - lo.FB), lo.FB, lo.shs) << (64 - lo.FB)) | low;
                }
                rec[0] = v;
                match = lo.FB ? ((v >> (64 - lo.FB)) == (want[0] >> (64 - lo.FB))) : true;
            } else {
                KbKey<WN> key;
                kb_fetch<WN>(ga, e, rec, key);
                match = true;
#pragma unroll
                for (int j = 0; j < WN; j++) if (j < lo.FW) match = match && (key.w[j] == want[j]);
            }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, match);
        if (match) {
            const uint64_t dst = o + __popc(m & kb_lanemask_lt());
#pragma unroll
            for (int j = 0; j < WN; j++) a.out[dst * WN + j] = rec[j];
        }
        o += __popc(m);
    }
}
