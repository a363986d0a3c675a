// kb_part.cuh — K2 (search path): most-significant-digit radix PARTITION of the 64-bit sort elements.
//
// The reference sorts every k-mer table with GNU sort (kstream/kstream.py:83-119) only so that the
// merge can walk equal (left,right) keys together (shared.py:442-475).  The search needs grouping, not
// order: records are partitioned by the top `bb` bits of the (mixed) flank key / flank hash into 2^bb
// buckets of a few thousand records, and the group pass (kb_hash.cuh) resolves the complete keys of one
// bucket in a shared-memory hash table.  bb <= 24 is split into at most three levels of <= 8 bits.
//
// One level = ONE read + ONE write of every element, like an LSD onesweep pass, but because no order
// has to be kept inside a bucket the pass needs neither stable ranking nor a look-back chain:
//   * rank inside the tile  : shared-memory atomicAdd on 256 digit counters (1 instruction per element
//                             instead of 8 ballots + counter bookkeeping)
//   * tile offset per digit : ONE global atomicAdd per (tile, digit) on a cursor initialised with the
//                             exclusive prefix of the digit histogram (exact, dense output)
//   * the elements are staged in shared memory in digit order, so every digit's run leaves the CTA as a
//     contiguous, coalesced store.
// Level l partitions every bucket ("parent") of level l-1 on its own: tiles never straddle parents, the
// cursor of (parent p, digit d) is child c = p * 2^bits + d.  Child counts come from kb_part_hist_kernel
// (level >= 2; level 1 is fused into K1) and are turned into offsets / cursors / tile maps by
// kb_plan_kernel + kb_tilemap_kernel.
#pragma once
#include "kb_common.cuh"

#define KB_PT_THREADS 512
#define KB_PT_ITEMS 16
#define KB_PT_TILE (KB_PT_THREADS * KB_PT_ITEMS)
#define KB_PT_MAXR 512                               // digits of at most 9 bits

struct KbPartArgs {
    const uint64_t* in;
    uint64_t* out;
    const unsigned long long* pstart;   // [n_parents + 1] parent ranges in `in`
    const uint32_t* ptile0;             // [n_parents + 1] first tile of each parent
    const uint32_t* tile_parent;        // [tiles]  (ignored when n_parents == 1)
    const uint32_t* prow;               // [n_parents] cursor row of each parent; null = the parent index itself.  (Multi-GPU: the
                                        // pieces of one level-0 digit received from different ranks share a row.)
    uint32_t n_parents;
    uint32_t shift, bits;               // digit = (e >> shift) & (2^bits - 1)
    unsigned long long* cursor;         // [n_parents << bits] absolute output offsets, advanced atomically
    unsigned long long* hist;           // kb_part_hist_kernel: [n_parents << bits] child counts (zeroed)
    int pair_mode;                      // 1: only the sizes of sibling PAIRS of children (2p, 2p + 1) are known (one bit more than the
                                        // histogram K1 could count): the pair's range is exact, the even child fills it from the left end,
                                        // the odd child from the right end — cursor[2p] starts at the pair's start and grows, cursor[2p+1]
                                        // starts at its end and shrinks; where they meet is the boundary (kb_pair_fix_kernel)
    const unsigned long long* out_elems; // != null: child c is written to the buffer at element address out_elems[c] (= pointer / 8)
                                        // instead of `out` — multi-GPU: the owner's receive buffer, a peer mapping over NVLink; the
                                        // cursor of c then counts from the first element of this rank's piece in that buffer
    // Slab layout (kb_extract_part.cuh): buckets are fixed-capacity regions filled through their cursors, not exact ranges.
    const unsigned long long* pend;     // != null: parent p = elements [p * pcap, min(pend[p], (p + 1) * pcap)) of `in` (pstart unused)
    const unsigned long long* pbegin;   // != null (with pend): only the segment from pbegin[p] on — what a batch of input files appended to the slab
    uint64_t pcap;
    uint32_t psel_n, psel_j0, psel_dps;  // psel_n != 0 (tile-by-division mode): the q-th parent of this launch is slab src * psel_dps + psel_j0 + q % psel_n,
    uint32_t psel_skip;                 // src = q / psel_n, + 1 from psel_skip on (psel_skip = 0xFFFFFFFF: no source is left out)
                                        // (multi-GPU: the slabs (source rank, digit j0 .. j0 + n) of one group of level-0 digits; this rank's own
                                        //  slabs, resident since K1, go through a launch of their own before the first group has arrived)
    uint64_t ccap;                      // != 0: child c owns [c * ccap, (c + 1) * ccap) of `out` (its cursor starts at c * ccap); a run that
    unsigned long long* ovf;            // does not fit is dropped and *ovf is raised (the host repeats the search on the exact path)
};

// element range of parent p
__device__ __forceinline__ void kb_part_parent(const KbPartArgs& a, uint32_t parent, uint64_t& ps, uint64_t& pe) {
    if (a.pend) {
        ps = (uint64_t)parent * a.pcap;
        pe = min((uint64_t)a.pend[parent], ps + a.pcap);
        if (a.pbegin) ps = max(ps, (uint64_t)a.pbegin[parent]);
        if (pe < ps) pe = ps;
    } else { ps = a.pstart[parent]; pe = a.pstart[parent + 1]; }
}

// tile -> its range [s, s + n_tile) and the cursor row of its parent
__device__ __forceinline__ bool kb_part_tile(const KbPartArgs& a, uint32_t tile, uint32_t& row, uint64_t& s, uint32_t& n_tile) {
    if (a.pend && !a.ptile0) {                                                    // (never together with pbegin)
        // slab parents whose capacity is a multiple of the tile: tile -> parent by division, no tile map (tiles past the fill level leave)
        const uint32_t tpp = (uint32_t)(a.pcap / KB_PT_TILE);
        const uint32_t q = tile / tpp;
        if (q >= a.n_parents) return false;
        uint32_t parent = q;
        if (a.psel_n) {
            uint32_t src = q / a.psel_n;
            if (src >= a.psel_skip) src++;
            parent = src * a.psel_dps + a.psel_j0 + q % a.psel_n;
        }
        uint64_t ps, pe;
        kb_part_parent(a, parent, ps, pe);
        s = ps + (uint64_t)(tile - q * tpp) * KB_PT_TILE;
        if (s >= pe) return false;
        n_tile = (uint32_t)min((uint64_t)KB_PT_TILE, pe - s);
        row = a.prow ? __ldg(a.prow + parent) : parent;
        return true;
    }
    if (tile >= __ldg(a.ptile0 + a.n_parents)) return false;
    const uint32_t parent = a.n_parents == 1 ? 0u : __ldg(a.tile_parent + tile);
    uint64_t ps, pe;
    kb_part_parent(a, parent, ps, pe);
    s = ps + (uint64_t)(tile - __ldg(a.ptile0 + parent)) * KB_PT_TILE;
    n_tile = (uint32_t)min((uint64_t)KB_PT_TILE, pe - s);
    row = a.prow ? __ldg(a.prow + parent) : parent;
    return true;
}

// ---- child histogram of one level (reads the previous level's output) ----------------------------
__global__ void __launch_bounds__(KB_PT_THREADS) kb_part_hist_kernel(const KbPartArgs a) {
    __shared__ uint32_t cnt[KB_PT_MAXR];
    const uint32_t tid = threadIdx.x;
    uint32_t parent, n_tile; uint64_t s;
    if (!kb_part_tile(a, blockIdx.x, parent, s, n_tile)) return;
    if (tid < KB_PT_MAXR) cnt[tid] = 0;
    __syncthreads();
    const uint32_t dmask = (1u << a.bits) - 1u;
#pragma unroll 4
    for (uint32_t i = tid; i < n_tile; i += KB_PT_THREADS) {
        const uint64_t e = kb_ld_stream(a.in + s + i);
        atomicAdd(&cnt[(uint32_t)(e >> a.shift) & dmask], 1u);
    }
    __syncthreads();
    if (tid <= dmask) {
        const uint32_t c = cnt[tid];
        if (c) atomicAdd(a.hist + (((size_t)parent << a.bits) | tid), (unsigned long long)c);
    }
}

// ---- counts -> offsets, cursors, tile prefix: three small launches (block sums, scan of the block sums, apply) -----------
// counts[nc] (u64) -> start[nc + 1] exclusive prefix, cursor[c] = start[c], tile0[nc + 1] = exclusive prefix of ceil(count / TILE).
// `cursor` may alias `counts` (each element is read before it is overwritten).
#define KB_PLAN_BLOCK 1024
struct KbPlanArgs {
    const unsigned long long* counts;
    uint32_t nc;
    unsigned long long* start;          // [nc + 1]
    unsigned long long* cursor;         // [nc] (may be null)
    uint32_t* tile0;                    // [nc + 1] (may be null)
    unsigned long long* part;           // scratch: [2 * ceil(nc / 1024)] block sums of counts / of tile counts
    // optional: also fold groups of `fold` consecutive counts (fold divides 1024) into parent counts: the level-0 histogram
    // from K1's two-level histogram; null = off
    unsigned long long* folded; uint32_t fold;
};

__device__ __forceinline__ void kb_block_scan2(unsigned long long& x, unsigned long long& y, unsigned long long* ws, unsigned long long* wt,
                                               unsigned long long& tx, unsigned long long& ty) {
    // inclusive scan of (x, y) over the 1024 threads of the CTA; (tx, ty) = CTA totals
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long ox = __shfl_up_sync(0xFFFFFFFFu, x, d), oy = __shfl_up_sync(0xFFFFFFFFu, y, d);
        if (lane >= (uint32_t)d) { x += ox; y += oy; }
    }
    if (lane == 31) { ws[warp] = x; wt[warp] = y; }
    __syncthreads();
    unsigned long long ax = 0, ay = 0; tx = 0; ty = 0;
    for (uint32_t w = 0; w < 32; w++) { if (w < warp) { ax += ws[w]; ay += wt[w]; } tx += ws[w]; ty += wt[w]; }
    x += ax; y += ay;
    __syncthreads();
}

__global__ void __launch_bounds__(KB_PLAN_BLOCK) kb_plan_reduce_kernel(const KbPlanArgs a) {
    __shared__ unsigned long long ws[32], wt[32], sv[KB_PLAN_BLOCK];
    const uint32_t i = blockIdx.x * KB_PLAN_BLOCK + threadIdx.x;
    const unsigned long long v = i < a.nc ? a.counts[i] : 0ULL;
    unsigned long long x = v, y = (v + KB_PT_TILE - 1) / KB_PT_TILE, tx, ty;
    if (a.folded) sv[threadIdx.x] = v;
    kb_block_scan2(x, y, ws, wt, tx, ty);
    if (threadIdx.x == 0) { a.part[2 * blockIdx.x] = tx; a.part[2 * blockIdx.x + 1] = ty; }
    if (a.folded && (threadIdx.x % a.fold) == 0 && i < a.nc) {
        unsigned long long s = 0;
        for (uint32_t j = 0; j < a.fold; j++) s += sv[threadIdx.x + j];
        a.folded[i / a.fold] = s;
    }
}

// exclusive scan of the nb (count sum, tile sum) pairs in place; ONE CTA
__global__ void __launch_bounds__(KB_PLAN_BLOCK) kb_plan_scan_kernel(unsigned long long* part, uint32_t nb) {
    __shared__ unsigned long long ws[32], wt[32];
    const uint32_t per = (nb + KB_PLAN_BLOCK - 1) / KB_PLAN_BLOCK;
    const uint32_t c0 = min(nb, threadIdx.x * per), c1 = min(nb, c0 + per);
    unsigned long long sx = 0, sy = 0;
    for (uint32_t c = c0; c < c1; c++) { sx += part[2 * c]; sy += part[2 * c + 1]; }
    unsigned long long x = sx, y = sy, tx, ty;
    kb_block_scan2(x, y, ws, wt, tx, ty);
    unsigned long long rx = x - sx, ry = y - sy;
    for (uint32_t c = c0; c < c1; c++) {
        const unsigned long long vx = part[2 * c], vy = part[2 * c + 1];
        part[2 * c] = rx; part[2 * c + 1] = ry;
        rx += vx; ry += vy;
    }
}

__global__ void __launch_bounds__(KB_PLAN_BLOCK) kb_plan_apply_kernel(const KbPlanArgs a) {
    __shared__ unsigned long long ws[32], wt[32];
    const uint32_t i = blockIdx.x * KB_PLAN_BLOCK + threadIdx.x;
    const unsigned long long v = i < a.nc ? a.counts[i] : 0ULL;
    const unsigned long long t = (v + KB_PT_TILE - 1) / KB_PT_TILE;
    unsigned long long x = v, y = t, tx, ty;
    kb_block_scan2(x, y, ws, wt, tx, ty);
    const unsigned long long bx = a.part[2 * blockIdx.x], by = a.part[2 * blockIdx.x + 1];
    if (i < a.nc) {
        a.start[i] = bx + x - v;
        if (a.cursor) a.cursor[i] = bx + x - v;
        if (a.tile0) a.tile0[i] = (uint32_t)(by + y - t);
    }
    if (i + 1 == a.nc) {
        a.start[a.nc] = bx + x;
        if (a.tile0) a.tile0[a.nc] = (uint32_t)(by + y);
    }
}

// root parent table of level 0 from K1's device-resident record counter: {0, n}, {0, tiles}
__global__ void kb_root_kernel(const unsigned long long* n_ptr, unsigned long long* root, uint32_t* roottile) {
    if (threadIdx.x == 0) {
        const unsigned long long n = *n_ptr;
        root[0] = 0; root[1] = n;
        roottile[0] = 0; roottile[1] = (uint32_t)((n + KB_PT_TILE - 1) / KB_PT_TILE);
    }
}

// ---- level 0 per batch of input files (sequences still arriving from the host) -------------------------------------------
// K1 runs per batch and counts the two-level children of the batch (hist2_b).  kb_batch_fold_kernel: level-0 counts of the
// batch = row sums, and the whole search's level-1 counts accumulate in total2.  One block per level-0 digit.
__global__ void __launch_bounds__(512) kb_batch_fold_kernel(const unsigned long long* hist2_b, unsigned long long* total2,
                                                            unsigned long long* counts0, uint32_t fold) {
    __shared__ unsigned long long ws[16];
    const uint32_t d0 = blockIdx.x, t = threadIdx.x;
    unsigned long long v = 0;
    if (t < fold) { v = hist2_b[(size_t)d0 * fold + t]; total2[(size_t)d0 * fold + t] += v; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((t & 31) == 0) ws[t >> 5] = v;
    __syncthreads();
    if (t == 0) { unsigned long long s = 0; for (int w = 0; w < 16; w++) s += ws[w]; counts0[d0] = s; }
}

// offsets of the batch's level-0 buckets: the batch's records are [*base, *base + sum) of the element array (K1 appends batch
// after batch), its bucket d starts at start[d]; start[nc0] = end = the next batch's base.  One block, nc0 <= 512.
__global__ void __launch_bounds__(512) kb_batch_plan_kernel(const unsigned long long* counts0, uint32_t nc0, const unsigned long long* base,
                                                            unsigned long long* start, unsigned long long* cursor,
                                                            unsigned long long* root, uint32_t* roottile) {
    __shared__ unsigned long long ws[16];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long v = t < nc0 ? counts0[t] : 0ULL;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x += o; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned long long add = *base, tot = 0;
    for (uint32_t w = 0; w < 16; w++) { if (w < warp) add += ws[w]; tot += ws[w]; }
    if (t < nc0) { start[t] = add + x - v; cursor[t] = add + x - v; }
    if (t == 0) {
        const unsigned long long b0 = *base;
        start[nc0] = b0 + tot;
        root[0] = b0; root[1] = b0 + tot;
        roottile[0] = 0; roottile[1] = (uint32_t)((tot + KB_PT_TILE - 1) / KB_PT_TILE);
    }
}

// tile -> parent map: one warp per parent
__global__ void __launch_bounds__(256) kb_tilemap_kernel(const uint32_t* tile0, uint32_t n_parents, uint32_t* tile_parent) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t p = blockIdx.x * 8 + (threadIdx.x >> 5); p < n_parents; p += gridDim.x * 8) {
        const uint32_t t0 = tile0[p], t1 = tile0[p + 1];
        for (uint32_t t = t0 + lane; t < t1; t += 32) tile_parent[t] = p;
    }
}

// ---- one partition level -------------------------------------------------------------------------------
#define KB_PT_DROP 0xFFFFFFFFFFFFFFFFULL
template <int MINB, bool PAIR = false, bool SLAB = false>
__global__ void __launch_bounds__(KB_PT_THREADS, MINB) kb_part_kernel(const KbPartArgs a) {
    extern __shared__ __align__(16) unsigned char kb_smem_raw[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(kb_smem_raw);                   // TILE
    uint64_t* dbase = skeys + KB_PT_TILE;                                         // MAXR: global offset - local start
    uint32_t* cnt = reinterpret_cast<uint32_t*>(dbase + KB_PT_MAXR);              // MAXR: digit counters, then local starts
    uint32_t* wsum = cnt + KB_PT_MAXR;                                            // MAXR / 32

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t parent, n_tile; uint64_t s;
    if (!kb_part_tile(a, blockIdx.x, parent, s, n_tile)) return;
    if (tid < KB_PT_MAXR) cnt[tid] = 0;
    __syncthreads();

    const uint32_t dmask = (1u << a.bits) - 1u;
    uint64_t key[KB_PT_ITEMS];
    uint16_t rank[KB_PT_ITEMS];
    if (n_tile == KB_PT_TILE) {
#pragma unroll
        for (int i = 0; i < KB_PT_ITEMS; i++) key[i] = kb_ld_stream(a.in + s + i * KB_PT_THREADS + tid);
#pragma unroll
        for (int i = 0; i < KB_PT_ITEMS; i++) rank[i] = (uint16_t)atomicAdd(&cnt[(uint32_t)(key[i] >> a.shift) & dmask], 1u);
    } else {
#pragma unroll
        for (int i = 0; i < KB_PT_ITEMS; i++) {
            const uint32_t idx = i * KB_PT_THREADS + tid;
            key[i] = idx < n_tile ? kb_ld_stream(a.in + s + idx) : 0ULL;
        }
#pragma unroll
        for (int i = 0; i < KB_PT_ITEMS; i++) {
            const uint32_t idx = i * KB_PT_THREADS + tid;
            rank[i] = idx < n_tile ? (uint16_t)atomicAdd(&cnt[(uint32_t)(key[i] >> a.shift) & dmask], 1u) : (uint16_t)0;
        }
    }
    __syncthreads();

    // ---- per digit: claim the output range (global cursor), local start -----------------------------
    uint32_t c = 0, lstart = 0;
    unsigned long long g = 0;
    bool drop = false;
    if (tid < KB_PT_MAXR) {
        c = cnt[tid];
        if (c) {
            if (PAIR && (tid & 1u)) g = atomicAdd(a.cursor + (((size_t)parent << a.bits) | tid), 0ULL - (unsigned long long)c) - (unsigned long long)c;   // claim downwards
            else g = atomicAdd(a.cursor + (((size_t)parent << a.bits) | tid), (unsigned long long)c);
            if (SLAB && g + c > ((((unsigned long long)parent << a.bits) | tid) + 1ULL) * a.ccap) { drop = true; *a.ovf = 1ULL; }
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

    // ---- stage in digit order ------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < KB_PT_ITEMS; i++) {
        const uint32_t idx = i * KB_PT_THREADS + tid;
        if (n_tile == KB_PT_TILE || idx < n_tile) skeys[cnt[(uint32_t)(key[i] >> a.shift) & dmask] + rank[i]] = key[i];
    }
    if (tid < KB_PT_MAXR) {
        // element address (pointer / 8) of the digit's run minus its position in the staged tile
        const unsigned long long base = a.out_elems ? (c ? a.out_elems[((size_t)parent << a.bits) | tid] : 0ULL)
                                                    : (unsigned long long)(reinterpret_cast<uintptr_t>(a.out) >> 3);
        dbase[tid] = (SLAB && drop) ? KB_PT_DROP : base + g - (unsigned long long)lstart;
    }
    __syncthreads();

    // ---- coalesced store ------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < KB_PT_ITEMS; i++) {
        const uint32_t pos = i * KB_PT_THREADS + tid;
        if (n_tile == KB_PT_TILE || pos < n_tile) {
            const uint64_t kv = skeys[pos];
            const unsigned long long db = dbase[(uint32_t)(kv >> a.shift) & dmask];
            if (!SLAB || db != KB_PT_DROP) *reinterpret_cast<uint64_t*>((db + pos) << 3) = kv;
        }
    }
}

// pair mode: cursors and bucket table from the exclusive prefix of the PAIR counts (pstart[n_pairs + 1])
__global__ void __launch_bounds__(256) kb_pair_expand_kernel(const unsigned long long* pstart, uint32_t n_pairs, unsigned long long* cursor,
                                                             unsigned long long* start) {
    for (uint32_t p = blockIdx.x * 256 + threadIdx.x; p < n_pairs; p += gridDim.x * 256) {
        const unsigned long long s0 = pstart[p], s1 = pstart[p + 1];
        cursor[2 * (size_t)p] = s0; cursor[2 * (size_t)p + 1] = s1;
        start[2 * (size_t)p] = s0; start[2 * (size_t)p + 1] = s1;        // (odd entries: fixed after the pass)
        if (p == n_pairs - 1) start[2 * (size_t)n_pairs] = s1;
    }
}
// after the pass the even child's cursor stands at the boundary between the two siblings
__global__ void __launch_bounds__(256) kb_pair_fix_kernel(const unsigned long long* cursor, uint32_t n_pairs, unsigned long long* start) {
    for (uint32_t p = blockIdx.x * 256 + threadIdx.x; p < n_pairs; p += gridDim.x * 256) start[2 * (size_t)p + 1] = cursor[2 * (size_t)p];
}

// ---- slab layout helpers ----------------------------------------------------------------------------------------------
// cursors of nc slabs of `cap` elements: cursor[c] = c * cap
__global__ void __launch_bounds__(256) kb_slab_init_kernel(unsigned long long* cursor, uint32_t nc, uint64_t cap) {
    for (uint32_t c = blockIdx.x * 256 + threadIdx.x; c < nc; c += gridDim.x * 256) cursor[c] = (unsigned long long)c * cap;
}
// fill levels after a pass: counts[c] = min(cursor[c], (c + 1) * cap) - c * cap (input of kb_plan_*: tile prefix of the next level)
__global__ void __launch_bounds__(256) kb_slab_counts_kernel(const unsigned long long* cursor, const unsigned long long* begin, uint32_t nc, uint64_t cap,
                                                             unsigned long long* counts) {
    for (uint32_t c = blockIdx.x * 256 + threadIdx.x; c < nc; c += gridDim.x * 256) {
        unsigned long long s = (unsigned long long)c * cap;
        const unsigned long long e = min(cursor[c], s + cap);
        if (begin) s = max(s, begin[c]);
        counts[c] = e > s ? e - s : 0ULL;
    }
}

// kb_slab_counts_kernel + kb_plan_{reduce,scan,apply}_kernel + kb_tilemap_kernel in ONE launch for up to KB_PLAN_BLOCK parents (one
// CTA): segment sizes of the parents -> exclusive prefix of their tile counts (tile0) -> tile -> parent map.  This is the planner
// of level 1 per batch of arriving files, where five tiny launches per batch were a tenth of the batch's GPU time.
__global__ void __launch_bounds__(KB_PLAN_BLOCK) kb_slab_plan_fused_kernel(const unsigned long long* cursor, const unsigned long long* begin, uint32_t nc, uint64_t cap,
                                                                           unsigned long long* start, uint32_t* tile0, uint32_t* tile_parent) {
    __shared__ unsigned long long ws[32], wt[32];
    __shared__ uint32_t st0[KB_PLAN_BLOCK + 1];
    const uint32_t i = threadIdx.x;
    unsigned long long v = 0;
    if (i < nc) {
        unsigned long long s = (unsigned long long)i * cap;
        const unsigned long long e = min(cursor[i], s + cap);
        if (begin) s = max(s, begin[i]);
        v = e > s ? e - s : 0ULL;
    }
    const unsigned long long t = (v + KB_PT_TILE - 1) / KB_PT_TILE;
    unsigned long long x = v, y = t, tx, ty;
    kb_block_scan2(x, y, ws, wt, tx, ty);
    if (i < nc) { start[i] = x - v; tile0[i] = (uint32_t)(y - t); st0[i] = (uint32_t)(y - t); }
    if (i + 1 == nc) { start[nc] = x; tile0[nc] = (uint32_t)y; st0[nc] = (uint32_t)y; }
    __syncthreads();
    const uint32_t lane = i & 31;
    for (uint32_t p = i >> 5; p < nc; p += KB_PLAN_BLOCK / 32) {
        const uint32_t t0 = st0[p], t1 = st0[p + 1];
        for (uint32_t q = t0 + lane; q < t1; q += 32) tile_parent[q] = p;
    }
}

static inline size_t kb_part_smem() { return (size_t)KB_PT_TILE * 8 + KB_PT_MAXR * 8 + KB_PT_MAXR * 4 + (KB_PT_MAXR / 32) * 4 + 16; }
static_assert(KB_PT_THREADS >= KB_PT_MAXR, "one thread per digit");
