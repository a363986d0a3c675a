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
// Tiles of 4096 bytes.  A byte is inside a header iff the line that contains it starts with '>'.  Three launches per batch of
// arrived files (KbFastaBatch: up to KB_FA_MAXF files share a launch):
// (1) kb_fa_count_kernel: last '\n' of every tile + its kept-byte count under BOTH hypotheses for the header state at the tile start;
// (2) kb_fa_offsets_kernel (one CTA per file): exclusive prefix-max over tiles -> line start before each tile -> header state at the tile
//     start -> the right count -> exclusive scan = output offset of every tile;
// (3) kb_fa_pack_kernel<true>: compaction through shared memory, coalesced byte stores.
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

// per-thread walk over its 16 bytes: keep mask, separator mask, flags
struct KbFaWalk { uint32_t keep, sep, flags, end_state, has_ls; };

// The walk is bit-parallel: the 16 bytes are classified four at a time (SWAR byte compares -> 16-bit position masks), the header
// state "the last line start before a byte decides" is a prefix propagation over those masks, and keep / separator masks are a
// few logic operations — about 10 instructions per byte instead of a 16-step byte loop.  Bytes below 0x21 other than '\n' (CR,
// blanks: rare) take a per-byte slow path.
__host__ __device__ __forceinline__ uint32_t kb_fa_gather4(uint32_t z) { return (((z >> 7) * 0x01020408u) >> 24) & 0xFu; }   // bit 7 of byte j -> bit j
__host__ __device__ __forceinline__ uint32_t kb_fa_eq4(uint32_t x, uint32_t pat) {                                              // bytes equal to the pattern's
    const uint32_t t = x ^ pat;
    return kb_fa_gather4(~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u);
}
__host__ __device__ __forceinline__ uint32_t kb_fa_lt4(uint32_t x, uint32_t bound7) {                                           // bytes < bound (bound <= 0x80)
    return kb_fa_gather4(~(((x & 0x7F7F7F7Fu) + (0x80808080u - bound7)) | x) & 0x80808080u);
}

struct KbFaBits {
    uint32_t V;        // valid bytes
    uint32_t NL, LOW;  // '\n'; bytes < 0x21 ('\n' included)
    uint32_t RU;       // 'U' / 'u'
    uint32_t LS;       // line starts: the byte before is '\n'
    uint32_t H;        // header starts: a line start holding '>' (FASTA only)
    uint32_t inh0;     // bytes inside a header line that STARTS among these 16 bytes
    uint32_t F;        // bytes before the first line start: they inherit the header state that flows in
};

__host__ __device__ __forceinline__ KbFaBits kb_fa_bits(const uint32_t (&w)[4], uint32_t nv, bool prev_nl, int fasta) {
    KbFaBits m;
    m.V = nv >= 16u ? 0xFFFFu : ((1u << nv) - 1u);
    uint32_t nl = 0, gt = 0, low = 0, ru = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        nl |= kb_fa_eq4(w[j], 0x0A0A0A0Au) << (4 * j);
        gt |= kb_fa_eq4(w[j], 0x3E3E3E3Eu) << (4 * j);
        low |= kb_fa_lt4(w[j], 0x21212121u) << (4 * j);
        ru |= kb_fa_eq4(w[j] | 0x20202020u, 0x75757575u) << (4 * j);
    }
    m.NL = nl & m.V; m.LOW = low & m.V; m.RU = ru & m.V;
    m.LS = ((m.NL << 1) | (prev_nl ? 1u : 0u)) & m.V;
    m.H = fasta ? (m.LS & gt) : 0u;
    uint32_t G = m.H, P = ~m.LS & 0xFFFFu;                     // Kogge-Stone: a byte that is no line start copies its left neighbour's state
    G |= P & (G << 1); P &= P << 1;
    G |= P & (G << 2); P &= P << 2;
    G |= P & (G << 4); P &= P << 4;
    G |= P & (G << 8);
    m.inh0 = G & m.V;
    m.F = (m.LS ? ((m.LS & (0u - m.LS)) - 1u) : 0xFFFFu) & m.V;
    return m;
}

// header state after these bytes: the last line start decides; none -> what flowed in
__host__ __device__ __forceinline__ uint32_t kb_fa_end_state(const KbFaBits& m, uint32_t state) {
    if (!m.LS) return state;
    uint32_t top = m.LS;                                       // isolate the highest line start
    top |= top >> 1; top |= top >> 2; top |= top >> 4; top |= top >> 8;
    top ^= top >> 1;
    return (m.H & top) ? 1u : 0u;
}

__host__ __device__ __forceinline__ KbFaWalk kb_fa_walk_bits(const KbFaBits& m, const uint32_t (&w)[4], uint32_t nv, bool next_nl, int fasta, uint32_t state) {
    KbFaWalk r{0, 0, 0, 0, 0};
    r.has_ls = m.LS ? 1u : 0u;
    r.end_state = kb_fa_end_state(m, state);
    const uint32_t inh = m.inh0 | (state ? m.F : 0u);
    r.keep = m.H; r.sep = m.H;                                 // the '>' of a header line becomes the record separator
    if (!fasta) { r.keep |= m.NL; r.sep |= m.NL; }             // plain input: the newline is the separator
    const uint32_t body = m.V & ~m.NL & ~inh;
    uint32_t drop = 0;
    uint32_t ctl = m.LOW & body;                               // CR / blanks outside headers: rare
    if (ctl) {
        const uint32_t nl_next = (m.NL >> 1) | ((next_nl && nv) ? (1u << (nv - 1u)) : 0u);
        while (ctl) {
            uint32_t i = 0;
            while (!((ctl >> i) & 1u)) i++;
            ctl &= ctl - 1u;
            const uint32_t q = i >> 2, wd = q == 0 ? w[0] : (q == 1 ? w[1] : (q == 2 ? w[2] : w[3]));   // (no dynamic indexing: registers)
            const uint32_t c = (wd >> (8u * (i & 3u))) & 0xFFu;
            if (c == '\r' && ((nl_next >> i) & 1u)) drop |= 1u << i;
            else if (c == ' ' || c == '\t' || c == '\r' || c == 0x0Bu || c == 0x0Cu) r.flags |= KB_FA_FLAG_SPACE;
        }
    }
    const uint32_t rest = body & ~drop;
    if (m.RU & rest) r.flags |= KB_FA_FLAG_RNA;
    r.keep |= rest;
    return r;
}

// reference form of the walk (one byte at a time): kept for tools/fa_walk_check.cu, which compares the two on random inputs
__host__ __device__ inline KbFaWalk kb_fa_walk_bytes(int fasta, const uint8_t (&b)[KB_FA_PER], uint32_t nv, uint8_t prev, uint8_t next, uint32_t state) {
    // prev = the byte before b[0] ('\n' at the very start), next = the byte after b[nv-1] ('\n' at the very end)
    KbFaWalk w{0, 0, 0, state, 0};
    for (int i = 0; i < KB_FA_PER; i++) {
        if (i >= (int)nv) continue;
        const uint8_t c = b[i];
        const uint8_t p = i == 0 ? prev : b[i - 1];
        const uint8_t nx = (i + 1 < (int)nv) ? b[i + 1] : next;
        const bool line_start = p == '\n';
        if (line_start) { w.has_ls = 1; w.end_state = (fasta && c == '>') ? 1u : 0u; }
        const bool hdr = w.end_state != 0;
        if (c == '\n') { if (!fasta) { w.keep |= 1u << i; w.sep |= 1u << i; } continue; }
        if (hdr) { if (line_start) { w.keep |= 1u << i; w.sep |= 1u << i; } continue; }      // the '>' becomes the separator
        if (c == '\r' && nx == '\n') continue;
        if (c == ' ' || c == '\t' || c == '\r' || c == 0x0B || c == 0x0C) w.flags |= KB_FA_FLAG_SPACE;
        if (c == 'U' || c == 'u') w.flags |= KB_FA_FLAG_RNA;
        w.keep |= 1u << i;
    }
    return w;
}

// this thread's 16 bytes as four little-endian words (past the end of the file: '\n'), the byte before and the byte after
struct KbFaChunk { uint32_t w[4]; uint32_t nv; bool prev_nl, next_nl; };

__device__ __forceinline__ KbFaChunk kb_fa_chunk(const KbFastaArgs& a, uint64_t tile_base, uint32_t tid) {
    KbFaChunk c;
    const uint64_t p = tile_base + (uint64_t)tid * KB_FA_PER;
    c.nv = p >= a.n ? 0u : (uint32_t)min((uint64_t)KB_FA_PER, a.n - p);
    if (c.nv == KB_FA_PER && ((reinterpret_cast<uintptr_t>(a.in + p) & 15) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4*>(a.in + p);
        c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t x = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) x |= (uint32_t)((4 * j + i) < (int)c.nv ? a.in[p + 4 * j + i] : (uint8_t)'\n') << (8 * i);
            c.w[j] = x;
        }
    }
    c.prev_nl = (p == 0 || p > a.n) ? true : a.in[p - 1] == '\n';
    c.next_nl = (p + c.nv < a.n) ? a.in[p + c.nv] == '\n' : true;
    return c;
}

// Header state flowing into this lane from the lanes below it (inside one warp): the nearest lower lane with a line start decides.
// Returns bit 1 = some lower lane decides, bit 0 = its state; *warp_code = what the whole warp hands on (same encoding).
__device__ __forceinline__ uint32_t kb_fa_lane_state_in(bool defines, uint32_t state, uint32_t lane, uint32_t* warp_code) {
    const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, defines), smask = __ballot_sync(0xFFFFFFFFu, defines && state);
    *warp_code = dmask ? (2u | ((smask >> (31u - (uint32_t)__clz(dmask))) & 1u)) : 0u;
    const uint32_t below = dmask & ((1u << lane) - 1u);
    return below ? (2u | ((smask >> (31u - (uint32_t)__clz(below))) & 1u)) : 0u;
}

// (1) last newline of the tile AND its kept-byte counts for both header states flowing into it: counts[t] (not in a header),
//      counts[n_tiles + t] (inside a header line)
__device__ __forceinline__ void kb_fa_count_tile(const KbFastaArgs& a, uint32_t n_tiles, uint32_t tile) {
    __shared__ unsigned long long ws[KB_FA_THREADS / 32];
    __shared__ uint32_t wstate[KB_FA_THREADS / 32], wsum0[KB_FA_THREADS / 32], wsum1[KB_FA_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_base = (uint64_t)tile * KB_FA_TILE;
    const KbFaChunk ch = kb_fa_chunk(a, tile_base, tid);
    const KbFaBits m = kb_fa_bits(ch.w, ch.nv, ch.prev_nl, a.fasta);
    // last newline of the tile: tile-relative position + 1 (0 = none), one reduction per warp
    const uint32_t last_rel = __reduce_max_sync(0xFFFFFFFFu, m.NL ? tid * KB_FA_PER + (32u - (uint32_t)__clz(m.NL)) : 0u);
    if (lane == 0) ws[warp] = last_rel ? tile_base + last_rel : 0ULL;
    uint32_t wcode;
    uint32_t in = kb_fa_lane_state_in(m.LS != 0, m.LS ? kb_fa_end_state(m, 0) : 0u, lane, &wcode);   // bit 1 set: a line start earlier in this tile fixes the state
    if (lane == 31) wstate[warp] = wcode;
    __syncthreads();
    if (!(in & 2u)) for (uint32_t w = 0; w < warp; w++) if (wstate[w] & 2u) in = wstate[w];
    uint32_t c0, c1;
    if (in & 2u) { c0 = c1 = __popc(kb_fa_walk_bits(m, ch.w, ch.nv, ch.next_nl, a.fasta, in & 1u).keep); }
    else {
        c0 = __popc(kb_fa_walk_bits(m, ch.w, ch.nv, ch.next_nl, a.fasta, 0u).keep);
        c1 = __popc(kb_fa_walk_bits(m, ch.w, ch.nv, ch.next_nl, a.fasta, 1u).keep);
    }
    c0 = __reduce_add_sync(0xFFFFFFFFu, c0); c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
    if (lane == 0) { wsum0[warp] = c0; wsum1[warp] = c1; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long mx = 0; uint32_t t0 = 0, t1 = 0;
        for (int w = 0; w < KB_FA_THREADS / 32; w++) { mx = max(mx, ws[w]); t0 += wsum0[w]; t1 += wsum1[w]; }
        a.last_nl[tile] = mx;
        a.counts[tile] = t0;
        a.counts[n_tiles + tile] = t1;
    }
}

// (2) ONE CTA: line start before every tile, header state at its start, the matching count, exclusive scan -> start[0 .. n_tiles]
__device__ __forceinline__ void kb_fa_offsets_file(const KbFastaArgs& a, uint32_t n_tiles, unsigned long long* start) {
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

// (3) compaction.  (WRITE = false: kept bytes per tile only.)
template <bool WRITE>
__device__ __forceinline__ void kb_fa_pack_tile(const KbFastaArgs& a, uint32_t tile) {
    __shared__ uint32_t wstate[KB_FA_THREADS / 32], wsum[KB_FA_THREADS / 32];
    __shared__ __align__(16) uint8_t stage[KB_FA_TILE + 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_base = (uint64_t)tile * KB_FA_TILE;
    const KbFaChunk ch = kb_fa_chunk(a, tile_base, tid);
    const KbFaBits m = kb_fa_bits(ch.w, ch.nv, ch.prev_nl, a.fasta);
    // header state at this thread's first byte: the nearest line start before it decides — in this warp, in an earlier warp of the
    // tile, or what flows into the tile
    uint32_t wcode;
    uint32_t in = kb_fa_lane_state_in(m.LS != 0, m.LS ? kb_fa_end_state(m, 0) : 0u, lane, &wcode);
    if (lane == 31) wstate[warp] = wcode;
    __syncthreads();
    if (!(in & 2u)) {
        in = 2u | a.hdr0[tile];                                           // state flowing into the tile
        for (uint32_t w = 0; w < warp; w++) if (wstate[w] & 2u) in = wstate[w];
    }
    const KbFaWalk w = kb_fa_walk_bits(m, ch.w, ch.nv, ch.next_nl, a.fasta, in & 1u);
    const uint32_t cnt = __popc(w.keep);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += o; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t add = 0, total = 0;
    for (uint32_t q = 0; q < KB_FA_THREADS / 32; q++) { if (q < warp) add += wsum[q]; total += wsum[q]; }
    if (!WRITE) {
        if (tid == 0) a.counts[tile] = total;
        if (w.flags) atomicOr(a.flags, w.flags);
        return;
    }
    if (w.flags) atomicOr(a.flags, w.flags);
    uint32_t o = add + inc - cnt;
#pragma unroll
    for (int i = 0; i < KB_FA_PER; i++)
        if ((w.keep >> i) & 1u) stage[o++] = ((w.sep >> i) & 1u) ? (uint8_t)'\n' : (uint8_t)(ch.w[i >> 2] >> (8 * (i & 3)));
    __syncthreads();
    // out of the CTA: bytes up to the first 16-byte boundary of the destination, then aligned 16-byte stores assembled from the
    // (differently aligned) staged bytes, then the remainder
    uint8_t* dst = a.out + a.start[tile];
    const uint32_t head = min(total, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    if (tid < head) dst[tid] = stage[tid];
    const uint32_t nvec = (total - head) / 16u;
    const uint32_t sh = 8u * (head & 3u);
    for (uint32_t v = tid; v < nvec; v += KB_FA_THREADS) {
        const uint32_t so = head + 16u * v;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage + (so & ~3u));
        const uint32_t x0 = sw[0], x1 = sw[1], x2 = sw[2], x3 = sw[3], x4 = sw[4];
        uint4 ov;
        ov.x = __funnelshift_r(x0, x1, sh); ov.y = __funnelshift_r(x1, x2, sh);
        ov.z = __funnelshift_r(x2, x3, sh); ov.w = __funnelshift_r(x3, x4, sh);
        *reinterpret_cast<uint4*>(dst + so) = ov;
    }
    for (uint32_t i = head + 16u * nvec + tid; i < total; i += KB_FA_THREADS) dst[i] = stage[i];
}

// ---- launches: the files of one arrived batch share each of the three launches (<= KB_FA_MAXF files per launch) -----------------
#define KB_FA_MAXF 8
struct KbFastaBatch {
    int n_files;
    uint32_t tile0[KB_FA_MAXF + 1];              // first tile of every file in the grid of the tile kernels
    unsigned long long* start_w[KB_FA_MAXF];     // = f[i].start, writable for the offsets kernel
    KbFastaArgs f[KB_FA_MAXF];
};

// file and file-relative tile of grid tile `g` (field-by-field selects: no dynamic indexing into the parameter space)
__device__ __forceinline__ KbFastaArgs kb_fa_locate(const KbFastaBatch& b, uint32_t g, uint32_t& tile, uint32_t& n_tiles) {
    KbFastaArgs a = b.f[0];
    uint32_t t0 = b.tile0[0], t1 = b.tile0[1];
#pragma unroll
    for (int i = 1; i < KB_FA_MAXF; i++)
        if (i < b.n_files && g >= b.tile0[i]) { a = b.f[i]; t0 = b.tile0[i]; t1 = b.tile0[i + 1]; }
    tile = g - t0; n_tiles = t1 - t0;
    return a;
}

__global__ void __launch_bounds__(KB_FA_THREADS) kb_fa_count_kernel(const KbFastaBatch b) {
    uint32_t tile, n_tiles;
    const KbFastaArgs a = kb_fa_locate(b, blockIdx.x, tile, n_tiles);
    kb_fa_count_tile(a, n_tiles, tile);
}

__global__ void __launch_bounds__(1024) kb_fa_offsets_kernel(const KbFastaBatch b) {                 // one CTA per file
    KbFastaArgs a = b.f[0];
    unsigned long long* start = b.start_w[0];
    uint32_t n_tiles = b.tile0[1] - b.tile0[0];
#pragma unroll
    for (int i = 1; i < KB_FA_MAXF; i++)
        if ((int)blockIdx.x == i) { a = b.f[i]; start = b.start_w[i]; n_tiles = b.tile0[i + 1] - b.tile0[i]; }
    kb_fa_offsets_file(a, n_tiles, start);
}

__global__ void __launch_bounds__(KB_FA_THREADS) kb_fa_pack_kernel(const KbFastaBatch b) {
    uint32_t tile, n_tiles;
    const KbFastaArgs a = kb_fa_locate(b, blockIdx.x, tile, n_tiles);
    kb_fa_pack_tile<true>(a, tile);
}
