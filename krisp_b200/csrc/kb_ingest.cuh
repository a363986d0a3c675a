// kb_ingest.cuh — GPU-side FASTA de-lining: raw (decompressed) file bytes -> the byte layout K1 consumes.
//
// Replaces the host loop of kstream._parse_FASTA (kstream/kstream.py:556-583; plain input: _parse_seqs :539-554) for the
// common file shape: lines ended by "\n" or "\r\n", no other whitespace.  The caller has already dropped the first line
// (kstream.py:450: the FASTA probe consumes it) and tells whether it held a '>' (FASTA) or not (plain: one record per line).
//   FASTA: every line that starts with '>' ends the current record -> its '>' becomes ONE separator byte, the rest of the
//          header line, every '\n' and every '\r' before a '\n' vanish; everything else is kept in order.
//   plain: '\n' stays (it separates the records), '\r' before '\n' vanishes.
// Output length <= input length: the packed bytes are written at the start of the file's slot of the sequence buffer, the
// caller pre-fills the slot with separators, so no length ever travels to the host.
// Anything the fast path does not reproduce exactly (other whitespace that str.strip() would remove, 'U'/'u' — RNA) only
// raises a flag; the host layer then re-ingests that input with its line-by-line restatement.
//
// Tiles of 4096 bytes.  A byte is inside a header iff the line that contains it starts with '>'.  Three launches per file:
// (1) kb_fa_count_kernel: last '\n' of every tile + its kept-byte count under BOTH hypotheses for the header state at the tile start;
// (2) kb_fa_offsets_kernel (one CTA): exclusive prefix-max over tiles -> line start before each tile -> header state at the tile
//     start -> the right count -> exclusive scan = output offset of every tile;
// (3) kb_fa_pack_kernel<true>: compaction through shared memory, coalesced byte stores.
// (kb_fa_lastnl_kernel / kb_fa_scan_kernel / kb_fa_pack_kernel<false> + the kb_plan_* scan are the same steps as seven launches.)
#pragma once
#include "kb_common.cuh"

#define KB_FA_THREADS 256
#define KB_FA_PER 16
#define KB_FA_TILE (KB_FA_THREADS * KB_FA_PER)
#define KB_FA_NONE 0xFFFFFFFFFFFFFFFFULL

#define KB_FA_FLAG_RNA 1u          // a 'U' / 'u' among the kept bytes
#define KB_FA_FLAG_SPACE 2u        // whitespace the fast path does not strip like str.strip() (blank, tab, VT, FF, lone CR)

struct KbFastaArgs {
    const uint8_t* in;             // body: the file content after its first line
    uint64_t n;
    int fasta;                     // 1 = FASTA, 0 = plain (one record per line)
    unsigned long long* last_nl;   // [tiles] position of the last '\n' of the tile (KB_FA_NONE: none) -> after the scan: line start before the tile
    uint8_t* hdr0;                 // [tiles] header state at the tile start
    unsigned long long* counts;    // [tiles] kept bytes
    const unsigned long long* start;   // [tiles + 1] exclusive prefix of counts
    uint8_t* out;                  // the file's slot (pre-filled with separators)
    unsigned int* flags;
};

__device__ __forceinline__ void kb_fa_load(const KbFastaArgs& a, uint64_t tile_base, uint32_t tid, uint8_t (&b)[KB_FA_PER], uint32_t& nv) {
    const uint64_t p = tile_base + (uint64_t)tid * KB_FA_PER;
    nv = p >= a.n ? 0u : (uint32_t)min((uint64_t)KB_FA_PER, a.n - p);
    if (nv == KB_FA_PER && ((reinterpret_cast<uintptr_t>(a.in + p) & 15) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4*>(a.in + p);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < KB_FA_PER; i++) b[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
    } else {
#pragma unroll
        for (int i = 0; i < KB_FA_PER; i++) b[i] = i < (int)nv ? a.in[p + i] : (uint8_t)'\n';
    }
}

// (1) last newline of every tile
__global__ void __launch_bounds__(KB_FA_THREADS) kb_fa_lastnl_kernel(const KbFastaArgs a) {
    __shared__ unsigned long long ws[KB_FA_THREADS / 32];
    const uint32_t tid = threadIdx.x;
    const uint64_t tile_base = (uint64_t)blockIdx.x * KB_FA_TILE;
    uint8_t b[KB_FA_PER]; uint32_t nv;
    kb_fa_load(a, tile_base, tid, b, nv);
    unsigned long long last = 0;                       // position + 1, 0 = none
#pragma unroll
    for (int i = 0; i < KB_FA_PER; i++) if (i < (int)nv && b[i] == '\n') last = tile_base + (uint64_t)tid * KB_FA_PER + i + 1;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, d));
    if ((tid & 31) == 0) ws[tid >> 5] = last;
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = 0;
        for (int w = 0; w < KB_FA_THREADS / 32; w++) m = max(m, ws[w]);
        a.last_nl[blockIdx.x] = m;                     // position + 1 (0 = no newline in this tile)
    }
}

// (2) exclusive prefix-max over the tiles (ONE CTA): last_nl[t] := start of the line that runs into tile t; hdr0[t]
__global__ void __launch_bounds__(1024) kb_fa_scan_kernel(const KbFastaArgs a, uint32_t n_tiles) {
    __shared__ unsigned long long ws[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_tiles + 1023u) / 1024u;
    const uint32_t c0 = min(n_tiles, tid * per), c1 = min(n_tiles, c0 + per);
    unsigned long long m = 0;
    for (uint32_t c = c0; c < c1; c++) m = max(m, a.last_nl[c]);
    unsigned long long x = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x = max(x, o); }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned long long before = 0;
    for (uint32_t w = 0; w < warp; w++) before = max(before, ws[w]);
    const unsigned long long ex = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (lane > 0) before = max(before, ex);
    unsigned long long run = before;                   // line start (= last newline + 1, 0 if none) before tile c0
    for (uint32_t c = c0; c < c1; c++) {
        const unsigned long long mine = a.last_nl[c];
        a.last_nl[c] = run;
        a.hdr0[c] = (a.fasta && run < a.n && a.in[run] == '>') ? 1 : 0;
        run = max(run, mine);
    }
}

// per-thread walk over its 16 bytes: keep mask, separator mask, flags
struct KbFaWalk { uint32_t keep, sep, flags, end_state, has_ls; };

__device__ __forceinline__ KbFaWalk kb_fa_walk(const KbFastaArgs& a, const uint8_t (&b)[KB_FA_PER], uint32_t nv, uint8_t prev, uint8_t next, uint32_t state) {
    // prev = the byte before b[0] ('\n' at the very start), next = the byte after b[nv-1] ('\n' at the very end)
    KbFaWalk w{0, 0, 0, state, 0};
#pragma unroll
    for (int i = 0; i < KB_FA_PER; i++) {
        if (i >= (int)nv) continue;
        const uint8_t c = b[i];
        const uint8_t p = i == 0 ? prev : b[i - 1];
        const uint8_t nx = (i + 1 < (int)nv) ? b[i + 1] : next;
        const bool line_start = p == '\n';
        if (line_start) { w.has_ls = 1; w.end_state = (a.fasta && c == '>') ? 1u : 0u; }
        const bool hdr = w.end_state != 0;
        if (c == '\n') { if (!a.fasta) { w.keep |= 1u << i; w.sep |= 1u << i; } continue; }
        if (hdr) { if (line_start) { w.keep |= 1u << i; w.sep |= 1u << i; } continue; }      // the '>' becomes the separator
        if (c == '\r' && nx == '\n') continue;
        if (c == ' ' || c == '\t' || c == '\r' || c == 0x0B || c == 0x0C) w.flags |= KB_FA_FLAG_SPACE;
        if (c == 'U' || c == 'u') w.flags |= KB_FA_FLAG_RNA;
        w.keep |= 1u << i;
    }
    return w;
}

// (1') last newline of the tile AND its kept-byte counts for both header states flowing into it: counts[t] (not in a header),
//      counts[n_tiles + t] (inside a header line)
__global__ void __launch_bounds__(KB_FA_THREADS) kb_fa_count_kernel(const KbFastaArgs a, uint32_t n_tiles) {
    __shared__ unsigned long long ws[KB_FA_THREADS / 32];
    __shared__ uint32_t wstate[KB_FA_THREADS / 32], wsum0[KB_FA_THREADS / 32], wsum1[KB_FA_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_base = (uint64_t)blockIdx.x * KB_FA_TILE;
    uint8_t b[KB_FA_PER]; uint32_t nv;
    kb_fa_load(a, tile_base, tid, b, nv);
    unsigned long long last = 0;                       // position + 1, 0 = none
#pragma unroll
    for (int i = 0; i < KB_FA_PER; i++) if (i < (int)nv && b[i] == '\n') last = tile_base + (uint64_t)tid * KB_FA_PER + i + 1;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, d));
    if (lane == 0) ws[warp] = last;
    const uint64_t p0 = tile_base + (uint64_t)tid * KB_FA_PER;
    const uint8_t prev = (p0 == 0 || p0 > a.n) ? (uint8_t)'\n' : a.in[p0 - 1];
    const uint8_t next = (p0 + nv < a.n) ? a.in[p0 + nv] : (uint8_t)'\n';
    const KbFaWalk probe = kb_fa_walk(a, b, nv, prev, next, 0);
    uint32_t x = probe.has_ls ? (2u | probe.end_state) : 0u;             // bit 1: defines the state, bit 0: the state
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d && !(x & 2u)) x = o; }
    if (lane == 31) wstate[warp] = x;
    __syncthreads();
    uint32_t in = 0;                                                     // bit 1 set: a line start earlier in this tile fixes the state
    for (uint32_t w = 0; w < warp; w++) if (wstate[w] & 2u) in = wstate[w];
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (lane > 0 && (up & 2u)) in = up;
    uint32_t c0, c1;
    if (in & 2u) { c0 = c1 = __popc(kb_fa_walk(a, b, nv, prev, next, in & 1u).keep); }
    else { c0 = __popc(probe.keep); c1 = __popc(kb_fa_walk(a, b, nv, prev, next, 1u).keep); }
    c0 = __reduce_add_sync(0xFFFFFFFFu, c0); c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
    if (lane == 0) { wsum0[warp] = c0; wsum1[warp] = c1; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = 0; uint32_t t0 = 0, t1 = 0;
        for (int w = 0; w < KB_FA_THREADS / 32; w++) { m = max(m, ws[w]); t0 += wsum0[w]; t1 += wsum1[w]; }
        a.last_nl[blockIdx.x] = m;
        a.counts[blockIdx.x] = t0;
        a.counts[n_tiles + blockIdx.x] = t1;
    }
}

// (2') ONE CTA: line start before every tile, header state at its start, the matching count, exclusive scan -> start[0 .. n_tiles]
__global__ void __launch_bounds__(1024) kb_fa_offsets_kernel(const KbFastaArgs a, uint32_t n_tiles, unsigned long long* start) {
    __shared__ unsigned long long ws[32], wc[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_tiles + 1023u) / 1024u;
    const uint32_t c0 = min(n_tiles, tid * per), c1 = min(n_tiles, c0 + per);
    unsigned long long m = 0;
    for (uint32_t c = c0; c < c1; c++) m = max(m, a.last_nl[c]);
    unsigned long long x = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d) x = max(x, o); }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned long long before = 0;
    for (uint32_t w = 0; w < warp; w++) before = max(before, ws[w]);
    const unsigned long long ex = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (lane > 0) before = max(before, ex);
    unsigned long long run = before, sum = 0;          // line start before tile c0; kept bytes of this thread's tiles
    for (uint32_t c = c0; c < c1; c++) {
        const unsigned long long mine = a.last_nl[c];
        const uint8_t h = (a.fasta && run < a.n && a.in[run] == '>') ? 1 : 0;
        a.last_nl[c] = run;
        a.hdr0[c] = h;
        sum += a.counts[(h ? n_tiles : 0u) + c];
        run = max(run, mine);
    }
    unsigned long long y = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, y, d); if (lane >= (uint32_t)d) y += o; }
    if (lane == 31) wc[warp] = y;
    __syncthreads();
    unsigned long long off = y - sum;
    for (uint32_t w = 0; w < warp; w++) off += wc[w];
    for (uint32_t c = c0; c < c1; c++) {
        start[c] = off;
        off += a.counts[(a.hdr0[c] ? n_tiles : 0u) + c];
    }
    if (tid == 1023) start[n_tiles] = off;
}

// (3) kept bytes per tile / (5) compaction.  WRITE = false: counts only.
template <bool WRITE>
__global__ void __launch_bounds__(KB_FA_THREADS) kb_fa_pack_kernel(const KbFastaArgs a) {
    __shared__ uint32_t wstate[KB_FA_THREADS / 32], wsum[KB_FA_THREADS / 32];
    __shared__ uint8_t stage[KB_FA_TILE];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_base = (uint64_t)blockIdx.x * KB_FA_TILE;
    uint8_t b[KB_FA_PER]; uint32_t nv;
    kb_fa_load(a, tile_base, tid, b, nv);
    const uint64_t p0 = tile_base + (uint64_t)tid * KB_FA_PER;
    const uint8_t prev = (p0 == 0 || p0 > a.n) ? (uint8_t)'\n' : a.in[p0 - 1];
    const uint8_t next = (p0 + nv < a.n) ? a.in[p0 + nv] : (uint8_t)'\n';
    // header state at this thread's first byte: "last line start wins" scan over the threads of the tile
    KbFaWalk probe = kb_fa_walk(a, b, nv, prev, next, 0);
    uint32_t code = probe.has_ls ? (2u | probe.end_state) : 0u;          // bit 1: defines the state, bit 0: the state
    uint32_t x = code;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (uint32_t)d && !(x & 2u)) x = o; }
    if (lane == 31) wstate[warp] = x;
    __syncthreads();
    uint32_t in = 2u | a.hdr0[blockIdx.x];                               // state flowing into the tile
    for (uint32_t w = 0; w < warp; w++) if (wstate[w] & 2u) in = wstate[w];
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (lane > 0 && (up & 2u)) in = up;
    const KbFaWalk w = kb_fa_walk(a, b, nv, prev, next, in & 1u);
    const uint32_t cnt = __popc(w.keep);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += o; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t add = 0, total = 0;
    for (uint32_t q = 0; q < KB_FA_THREADS / 32; q++) { if (q < warp) add += wsum[q]; total += wsum[q]; }
    if (!WRITE) {
        if (tid == 0) a.counts[blockIdx.x] = total;
        if (w.flags) atomicOr(a.flags, w.flags);
        return;
    }
    if (w.flags) atomicOr(a.flags, w.flags);
    uint32_t o = add + inc - cnt;
#pragma unroll
    for (int i = 0; i < KB_FA_PER; i++) if ((w.keep >> i) & 1u) stage[o++] = ((w.sep >> i) & 1u) ? (uint8_t)'\n' : b[i];
    __syncthreads();
    uint8_t* dst = a.out + a.start[blockIdx.x];
    for (uint32_t i = tid; i < total; i += KB_FA_THREADS) dst[i] = stage[i];
}
