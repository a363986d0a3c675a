// kb_api.cu — host side of libkrisp_b200.so: context, device workspaces, stage sequencing, C ABI.
// (interface and the reference entry points each call replaces: include/krisp_b200.h)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/krisp_b200.h"
#include "kb_common.cuh"
#include "kb_extract.cuh"
#include "kb_sort.cuh"
#include "kb_group.cuh"
#include "kb_group_fast.cuh"
#include "kb_part.cuh"
#include "kb_hash.cuh"
#include "kb_hash_stream.cuh"
#include "kb_hash_warp.cuh"
#include "kb_extract_part.cuh"
#include "kb_extract_sym.cuh"
#include "kb_ingest.cuh"
#include "kb_prefilter.cuh"
#include "kb_rows.cuh"

#define KB_VERSION_STR "krisp_b200 0.1.0 sm_100a"

// ---- small RAII-free device buffer that only grows --------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct kb_result {
    kb_result_view v;
    // flank / masks / group sizes / rows live in the context's pinned result arena (valid until the next search on that context)
    std::vector<uint64_t> run_offset, records;
    const char* rows = nullptr;          // CSV rows, ascending (left, right)
    size_t rows_len = 0;
    int row_bytes = 0;
};

struct kb_table {
    std::vector<uint64_t> records;
    int record_words = 1;
};

// partition plan of the search path (kb_part.cuh / kb_hash.cuh); see make_plan
struct PartPlan {
    int levels = 0, bits[3] = {0, 0, 0}, bb = 0;
    uint32_t slots_log2 = 11;
    bool fast = false;                   // kb_hash_fast_kernel applies (one-word records, <= 64 files, D <= 8)
    bool stream = false;                 // kb_hash_stream_kernel applies (one-word records, D <= 8, any file count)
    // device tables inside ctx->plan (byte offsets), per level: counts/cursors [NC], starts [NC + 1], tile prefix [NC + 1]
    size_t off_cnt[3] = {0, 0, 0}, off_start[3] = {0, 0, 0}, off_tile0[3] = {0, 0, 0}, off_part = 0, off_tilemap = 0, bytes = 0;
    size_t off_pair = 0;                 // pair mode (17 bits in two levels): 2^16 pair counts | 2^16 + 1 pair offsets
    uint32_t nc[3] = {0, 0, 0};          // children per level over the whole key space: 2^(bits[0] + ... + bits[l])
    uint32_t ncl[3] = {0, 0, 0};         // children per level this GPU works on (= nc on one GPU; its shard's part on several)
};

// slab search path (kb_extract_part.cuh): per level the number of slabs, their capacity, and where their tables live in ctx->plan
struct SlabPlan {
    int levels = 0, bits[3] = {0, 0, 0};
    uint32_t nc[3] = {0, 0, 0};
    uint64_t cap[3] = {0, 0, 0};
    size_t off_cur[3] = {0, 0, 0}, off_counts = 0, off_start = 0, off_part = 0, off_tab0 = 0, off_tile0[3] = {0, 0, 0}, off_tilemap = 0, off_snap = 0, bytes = 0;
    uint64_t max_tiles = 0;
    bool sym = false;                    // level 0 moves one window item per window, ranked by the strand-symmetric core digit (kb_extract_sym.cuh)
    uint64_t core_mask = 0;
};

// Level 0 of the partition per batch of input files (host buffers still in flight): see kb_batch_*_kernel in kb_part.cuh.
#define KB_MAX_BATCHES 16
struct BatchL0 {
    bool enabled = false;
    int n_batches = 0;
    uint32_t nc0 = 0, nc1 = 0, bits0 = 0, bits1 = 0;
    unsigned long long *hist2 = nullptr, *total2 = nullptr, *counts_all = nullptr, *start_all = nullptr, *cursors = nullptr, *roots = nullptr;
    uint32_t *roottiles = nullptr, *ptile0_all = nullptr, *prow = nullptr;
    unsigned long long* scratch_start = nullptr;
    uint64_t* out = nullptr;             // partition output (the other ping-pong buffer)
};

struct CustomParents {                  // parents of the first level run, when they are not the previous level's children
    const unsigned long long* pstart;   // [n + 1]
    const uint32_t* ptile0;             // [n + 1]
    const uint32_t* prow;               // [n]
    uint32_t n;
    uint64_t max_tiles;
};

struct kb_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host -> device sequence copies run here, so that K1 can start on the first files while the rest is in flight
    std::vector<cudaEvent_t> copy_events;   // event pool: copy_events[i] = "local file i is resident" (null entry = copied on the main stream)
    std::vector<cudaEvent_t> file_event;
    cudaEvent_t main_event = nullptr;
    std::string err;

    // configuration
    bool configured = false;
    KbLayout lo{};
    int soft_mode = 0;
    uint8_t is_ingroup[KB_MAX_FILES]{};
    long long opt_sort_bits = 32, opt_mix = 1, opt_want_records = 0, opt_profile = 0, opt_result_cap = 1 << 16, opt_sort_variant = 0, opt_fast_group = 1;
    long long opt_group_algo = 1;        // 1 = partition + bucket hash (kb_part.cuh, kb_hash.cuh), 0 = radix sort + segmented pass
    long long opt_bucket_bits = -1;      // -1 = from the input size
    long long opt_hash_slots_log2 = 0;   // 0 = default
    long long opt_hash_stream = 1;       // 1 = persistent TMA-fed bucket hash kernel for the fast shape
    long long opt_pair_hist = 1;         // 1 = 17 bucket bits in two levels run on K1's 16-bit histogram (sibling pairs fill their range from both ends)
    long long opt_fused_hist = 1;        // 1 = K1 also counts the level-1 children (single-GPU search path)
    long long opt_batch_level0 = 1;      // 1 = partition level 0 per batch of arriving files (hidden under the host -> device copy)
    long long opt_render_rows = 1;       // CSV rows of the survivors rendered + ordered on the device (kb_result_rows)
    long long opt_have_outgroup = 1;     // consensus letters: ingroup only (an outgroup was given) / every occurrence
    long long opt_strands = 0;           // kb_extract_sorted only: 0 = windows + reverse complements, 1 = windows only, 2 = canonical k-mers
    long long opt_lazy_records = 1;      // multi-word records: 1 = filter by flank hash, build records only for what is left (kb_prefilter.cuh)
    long long opt_slab = 1;              // one-word records: 1 = K1 fused with partition level 0 into fixed-capacity slabs (kb_extract_part.cuh)
    long long opt_hash_warp = 1;         // 1 = bucket hash kernel with per-warp streaming (kb_hash_warp.cuh) instead of the CTA-wide stream kernel
    long long opt_group_sizes = 0;       // survivors' group sizes on the bucket-hash path: 1 = always, 0 = only with want_records (the record gather needs them)
    long long opt_rank_rows = 1;         // rows of <= KB_RANK_MAX survivors ordered by counting ranks (one launch) instead of the chunked LSD sort
    long long opt_hash_warps = 0;        // kb_hash_warp.cuh, shared table: warps per CTA (8 or 10); 0 = 8
    long long opt_hash_shared = -1;      // kb_hash_warp.cuh: 1 = one table per CTA, 0 = one per warp, -1 = by table size (>= 1024 slots: per CTA)
    long long opt_slab_cap = 0;          // != 0: force the capacity of every slab (tests: overflow -> exact path)
    long long opt_sym = -1;              // slab path: 1 = strand-symmetric level 0 on window items where the layout allows it (kb_extract_sym.cuh),
                                         // 0 = records, -1 = items only where the exchange is the bottleneck (>= 4 GPUs: half the NVLink bytes;
                                         // on one GPU forming the records in level 1 costs more than level 0 saves)
    bool slab_off = false;               // a slab overflowed on these sequences: searches use the exact path until they change
    bool lazy_now = false;               // the running search uses it
    int strands_now = 0;                 // strand mode of the running K1 (kb_extract_sorted; searches always use both strands)
    int bb_extra = 0;                    // bucket bits added to the size-based plan: learnt when a search deferred too many buckets
                                         // (divergent genomes: far more distinct keys per record than the plan assumes); kept per layout
    bool replan_ok = false;              // the running search may give up on a too coarse plan (kb_search retries with bb_extra + 2)
    long long opt_shard_bits0 = 0;       // multi-GPU: bits of partition level 0 (the exchange); 0 = log2(shards) + 2

    // sequences
    DevBuf bases;
    uint64_t n_bases = 0;
    std::vector<uint64_t> file_table_host;
    std::vector<uint32_t> batch_rows_host;
    std::vector<uint64_t> file_starts;   // local files
    std::vector<uint32_t> file_gid;
    DevBuf d_file_starts, d_file_gid;

    // FASTA files whose raw bytes are on their way (copy stream) and whose de-lining kernels have not been enqueued yet
    struct PendingFasta { size_t raw_off; uint64_t nb; int fasta; uint64_t slot_off; cudaEvent_t ev; size_t file_idx; };
    std::vector<PendingFasta> pending_fasta;    // one per kb_add_fasta since the last clear, in file order
    size_t fasta_done = 0;                       // how many of them have been de-lined
    size_t fa_round = 0;                         // de-lining launches so far (they take the side streams in turn)
    std::vector<int> fasta_of_file;              // local file index -> index into pending_fasta (-1: added as parsed sequence)
    size_t raw_used = 0;                         // bytes of rawbuf in use
#define KB_FA_STREAMS 8
    cudaStream_t fa_stream[KB_FA_STREAMS] = {};  // the de-lining chains of different files overlap (each is a handful of tiny dependent launches)
    DevBuf fa_work_s[KB_FA_STREAMS];
    std::vector<cudaEvent_t> fa_done_ev;         // event pool: "file i is de-lined"

    // workspaces
    DevBuf rowkeyA, rowkeyB, rowtext;    // row rendering: survivor indices being sorted, the text
    DevBuf brun;                         // lazy records: per bucket (start, length) of the kept elements
    DevBuf batchbuf, rawbuf, fa_work, fa_flags;   // FASTA de-lining: raw file bytes, tile tables, flag word
    DevBuf entC;                         // sharded slab search with three levels: output of level 2 (entA is the staging buffer the copies still read)
    DevBuf entA, entB, recs, status, small, res_flank, res_in, res_out, res_size, res_run, gather_off, gather_out, taint, plan, deferred;
    uint64_t* h_pinned = nullptr;        // 64 x u64 scratch for small D2H reads
    uint8_t* h_arena = nullptr;          // pinned result arena: survivor table + rows of the LAST search (no pageable copies)
    size_t h_arena_cap = 0;
    uint64_t result_cap = 0;

    // last-search bookkeeping
    uint64_t launches = 0, alg_bytes = 0, alg_rec_bytes = 0;   // alg_rec_bytes: per-record bytes of stages enqueued before the record count is known
    int passes = 0;
    std::vector<std::pair<std::string, float>> profile;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> prof_events;

    // shard state
    SlabPlan shard_sp;
    PartPlan shard_plan;
    int shard_n = 0, shard_index = 0;
    std::vector<uint64_t> shard_tab_host;
    DevBuf shard_tab;
    DevBuf recvbuf;                      // IPC-exported receive buffer of the fused partition + exchange
    std::vector<void*> peer_ptr;         // peer_ptr[r] = rank r's receive buffer mapped here (own entry = recvbuf.p)
    std::vector<uint64_t> scatter_host;  // staging of the per-digit tables of kb_shard_scatter
    std::vector<uint64_t> shard_child_host;   // kb_shard_count: K1's two-level histogram (children of level 1 over the whole key space), if counted
    std::vector<uint64_t> shard_child_set;    // kb_shard_set_child_counts: level-1 child counts of this shard's digits, summed over the source ranks
    bool sep_filled = false;             // the sequence buffer holds separators in [n_bases, sep_upto) (host-buffer adds rely on it)
    uint64_t sep_upto = 0, reserve_hint = 0;
    int own_first = 0, own_count = -1;   // replicated sequences: K1 of the shard calls covers only these local files (-1: all)
    int shard_direct = 0;                // the last exchange went through kb_shard_scatter (input of kb_shard_search = recvbuf)
    bool shard_slab = false;             // kb_shard_slab_plan succeeded: shard_sp / shard_plan describe the slab exchange
    bool shard_own_done = false;         // kb_shard_slab_own ran level 1 on this rank's own slabs: the group passes skip that source
    long long opt_shard_bb_extra = 0;    // bucket bits added to the sharded slab plan (set on every rank alike after a re-plan vote)
    uint64_t shard_total_bases = 0;
    uint64_t shard_n_records = 0;
    int shard_send_in_B = 0;             // partitioned records are in entB (else entA)
};

// layout of the `small` device buffer (u64 units)
enum { SM_NOUT = 0, SM_NRES = 1, SM_STATS = 2 /*4*/, SM_NTAINT = 6, SM_ERR = 7, SM_OVF = 8 /* a slab overflowed */, SM_TICKET = 16 /* u32 x 16 */,
       SM_HIST = 24 /* 9*256 */, SM_ROOT = SM_HIST + 9 * 256 /* u64 x 2: {0, n} */, SM_ROOTTILE = SM_ROOT + 2 /* u32 x 2: {0, tiles} */, SM_TOTAL = SM_ROOT + 4 };

#define KB_REPLAN 1                      // internal: the search gave up on its plan (never leaves the library)
#define KB_SLABOVF 2                     // internal: a slab overflowed; repeat on the exact path
#define KB_RESGROW 3                     // internal (pipelined multi-GPU search): the survivor table was too small and has been grown

static int fail(kb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? KB_ENOMEM : KB_ECUDA,                    \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                            \
    } while (0)

static int launch_plan(kb_ctx* ctx, KbPlanArgs pa, const PartPlan& pl);

static int ensure(kb_ctx* ctx, DevBuf& b, size_t bytes, bool keep = false) {
    if (bytes <= b.cap) return KB_OK;
    size_t want = bytes + bytes / 16 + 256;
    void* np = nullptr;
    CU(cudaMalloc(&np, want));
    if (keep && b.p && b.cap) CU(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, ctx->stream));
    if (b.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(b.p)); }
    b.p = np; b.cap = want;
    return KB_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__ != KB_OK) return rc__; } while (0)

// ---- profiling helpers ---------------------------------------------------------------------------
static void prof_begin(kb_ctx* ctx, const char* name) {
    if (!ctx->opt_profile) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
    ctx->prof_events.push_back({name, {a, b}});
}
static void prof_end(kb_ctx* ctx) {
    if (!ctx->opt_profile || ctx->prof_events.empty()) return;
    cudaEventRecord(ctx->prof_events.back().second.second, ctx->stream);
}
static void prof_collect(kb_ctx* ctx) {
    ctx->profile.clear();
    for (auto& e : ctx->prof_events) {
        float ms = 0;
        cudaEventSynchronize(e.second.second);
        cudaEventElapsedTime(&ms, e.second.first, e.second.second);
        ctx->profile.push_back({e.first, ms});
        cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second);
    }
    ctx->prof_events.clear();
}

extern "C" {

const char* kb_version(void) { return KB_VERSION_STR; }

int kb_create(int device, kb_ctx** out) {
    if (!out) return KB_EINVAL;
    *out = nullptr;
    kb_ctx* ctx = new (std::nothrow) kb_ctx();
    if (!ctx) return KB_ENOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { ctx->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); *out = ctx; return KB_ECUDA; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { ctx->err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e); *out = ctx; return KB_ECUDA; }
    ctx->n_sm = prop.multiProcessorCount;
    if (prop.major < 10) { ctx->err = "krisp_b200 needs an sm_100a device (found sm_" + std::to_string(prop.major * 10 + prop.minor) + ")"; *out = ctx; return KB_EUNSUPPORTED; }
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { ctx->err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); *out = ctx; return KB_ECUDA; }
    ctx->stream = ctx->own_stream;
    e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->main_event, cudaEventDisableTiming);
    if (e != cudaSuccess) { ctx->err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); *out = ctx; return KB_ECUDA; }
    e = cudaMallocHost(&ctx->h_pinned, 64 * sizeof(uint64_t));
    if (e != cudaSuccess) { ctx->err = std::string("cudaMallocHost: ") + cudaGetErrorString(e); *out = ctx; return KB_ENOMEM; }
    *out = ctx;
    return KB_OK;
}

void kb_destroy(kb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->bases, &ctx->d_file_starts, &ctx->d_file_gid, &ctx->entA, &ctx->entB, &ctx->entC, &ctx->recs, &ctx->status,
                      &ctx->small, &ctx->res_flank, &ctx->res_in, &ctx->res_out, &ctx->res_size, &ctx->res_run,
                      &ctx->gather_off, &ctx->gather_out, &ctx->taint, &ctx->plan, &ctx->deferred, &ctx->shard_tab, &ctx->batchbuf, &ctx->rawbuf, &ctx->fa_work, &ctx->fa_flags, &ctx->brun, &ctx->rowkeyA, &ctx->rowkeyB, &ctx->rowtext};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
    for (size_t r = 0; r < ctx->peer_ptr.size(); r++) if (ctx->peer_ptr[r] && (int)r != ctx->shard_index) cudaIpcCloseMemHandle(ctx->peer_ptr[r]);
    if (ctx->recvbuf.p) cudaFree(ctx->recvbuf.p);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (int i = 0; i < KB_FA_STREAMS; i++) {
        if (ctx->fa_stream[i]) { cudaStreamSynchronize(ctx->fa_stream[i]); cudaStreamDestroy(ctx->fa_stream[i]); }
        if (ctx->fa_work_s[i].p) cudaFree(ctx->fa_work_s[i].p);
    }
    for (cudaEvent_t ev : ctx->fa_done_ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->copy_events) cudaEventDestroy(ev);
    if (ctx->main_event) cudaEventDestroy(ctx->main_event);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* kb_last_error(const kb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int kb_set_stream(kb_ctx* ctx, void* s) {
    if (!ctx) return KB_EINVAL;
    ctx->stream = (s == KB_STREAM_OWN) ? ctx->own_stream : (cudaStream_t)s;   // NULL = the default stream
    return KB_OK;
}

int kb_set_option(kb_ctx* ctx, const char* name, long long value) {
    if (!ctx || !name) return KB_EINVAL;
    std::string n(name);
    if (n == "sort_bits") { if (value < 1 || value > 64) return fail(ctx, KB_EINVAL, "sort_bits must be in 1..64"); ctx->opt_sort_bits = value; }
    else if (n == "mix") ctx->opt_mix = value ? 1 : 0;
    else if (n == "want_records") ctx->opt_want_records = value ? 1 : 0;
    else if (n == "profile") ctx->opt_profile = value ? 1 : 0;
    else if (n == "sort_variant") ctx->opt_sort_variant = value;
    else if (n == "fast_group") ctx->opt_fast_group = value ? 1 : 0;
    else if (n == "group_algo") ctx->opt_group_algo = value ? 1 : 0;
    else if (n == "hash_stream") ctx->opt_hash_stream = value ? 1 : 0;
    else if (n == "fused_hist") ctx->opt_fused_hist = value ? 1 : 0;
    else if (n == "pair_hist") ctx->opt_pair_hist = value ? 1 : 0;
    else if (n == "batch_level0") ctx->opt_batch_level0 = value ? 1 : 0;
    else if (n == "lazy_records") ctx->opt_lazy_records = value ? 1 : 0;
    else if (n == "strands") { if (value < 0 || value > 2) return fail(ctx, KB_EINVAL, "strands: 0 (both), 1 (forward) or 2 (canonical)"); ctx->opt_strands = value; }
    else if (n == "slab") { ctx->opt_slab = value ? 1 : 0; ctx->slab_off = false; }
    else if (n == "hash_warp") ctx->opt_hash_warp = value ? 1 : 0;
    else if (n == "sym") ctx->opt_sym = value < 0 ? -1 : (value ? 1 : 0);
    else if (n == "hash_shared") ctx->opt_hash_shared = value < 0 ? -1 : (value ? 1 : 0);
    else if (n == "group_sizes") ctx->opt_group_sizes = value ? 1 : 0;
    else if (n == "rank_rows") ctx->opt_rank_rows = value ? 1 : 0;
    else if (n == "hash_warps") {
        if (value != 0 && value != KB_HW_WARPS && value != KB_HW_WARPS_WIDE) return fail(ctx, KB_EINVAL, "hash_warps: 0 (automatic), 8 or 10");
        ctx->opt_hash_warps = value;
    }
    else if (n == "shard_bb_extra") { if (value < 0 || value > 12) return fail(ctx, KB_EINVAL, "shard_bb_extra must be in 0..12"); ctx->opt_shard_bb_extra = value; }
    else if (n == "slab_cap") { if (value < 0) return fail(ctx, KB_EINVAL, "slab_cap must be >= 0"); ctx->opt_slab_cap = value; ctx->slab_off = false; }
    else if (n == "render_rows") ctx->opt_render_rows = value ? 1 : 0;
    else if (n == "have_outgroup") ctx->opt_have_outgroup = value ? 1 : 0;
    else if (n == "shard_bits0") { if (value < 0 || value > 9) return fail(ctx, KB_EINVAL, "shard_bits0 must be in 0..9"); ctx->opt_shard_bits0 = value; }
    else if (n == "bucket_bits") { if (value < -1 || value > 24) return fail(ctx, KB_EINVAL, "bucket_bits must be in -1..24"); ctx->opt_bucket_bits = value; }
    else if (n == "hash_slots_log2") { if (value != 0 && (value < 4 || value > 12)) return fail(ctx, KB_EINVAL, "hash_slots_log2 must be 0 or in 4..12"); ctx->opt_hash_slots_log2 = value; }
    else if (n == "result_cap") { if (value < 1) return fail(ctx, KB_EINVAL, "result_cap must be >= 1"); ctx->opt_result_cap = value; }
    else return fail(ctx, KB_EINVAL, "unknown option " + n);
    if (ctx->configured) {   // re-derive the sort plan
        uint8_t ing[KB_MAX_FILES];
        memcpy(ing, ctx->is_ingroup, sizeof ing);
        return kb_configure(ctx, ctx->lo.L, ctx->lo.D, ctx->lo.R, ctx->soft_mode, ctx->lo.n_files, ing);
    }
    return KB_OK;
}

int kb_configure(kb_ctx* ctx, int L, int D, int R, int soft_mode, int n_files, const uint8_t* is_ingroup) {
    if (!ctx) return KB_EINVAL;
    if (L < 0 || D < 0 || R < 0 || L + D + R < 1) return fail(ctx, KB_EINVAL, "need L, D, R >= 0 and k = L+D+R >= 1");
    if (n_files < 1 || !is_ingroup) return fail(ctx, KB_EINVAL, "need n_files >= 1 and an is_ingroup array");
    if (n_files > KB_MAX_FILES) return fail(ctx, KB_EUNSUPPORTED, "more than 256 input files (8-bit file id)");
    const int k = L + D + R;
    const int total_bits = 2 * k + KB_IDBITS;
    if (total_bits > 64 * KB_MAX_W) return fail(ctx, KB_EUNSUPPORTED, "k-mer longer than 252 bases");
    if (D > 8 * KB_MAX_MW) return fail(ctx, KB_EUNSUPPORTED, "diagnostic region longer than 128 bases");
    KbLayout lo{};
    lo.L = L; lo.D = D; lo.R = R; lo.k = k;
    lo.FB = 2 * (L + R);
    lo.FW = std::max(1, (lo.FB + 63) / 64);
    lo.direct = total_bits <= 64;
    if (lo.direct) lo.W = 1;
    else { int w = (total_bits + 63) / 64; lo.W = w <= 2 ? 2 : (w <= 4 ? 4 : 8); }
    lo.mix = lo.direct ? (int)ctx->opt_mix : 0;
    lo.shs = (uint32_t)((lo.FB + 1) / 2);
    const int keybits = lo.direct ? lo.FB : 32;
    const int prefix = std::min<int>(keybits, (int)ctx->opt_sort_bits);
    lo.P = (prefix + 7) / 8;
    lo.cmpbits = lo.direct ? std::min(lo.FB, 8 * lo.P) : 8 * lo.P;
    lo.n_files = n_files;
    lo.PW = (n_files + 31) / 32;
    lo.MW = (D + 7) / 8;
    if (!ctx->configured || ctx->lo.L != lo.L || ctx->lo.D != lo.D || ctx->lo.R != lo.R || ctx->lo.n_files != lo.n_files) { ctx->bb_extra = 0; ctx->slab_off = false; }
    ctx->lo = lo;
    ctx->soft_mode = soft_mode ? 1 : 0;
    memset(ctx->is_ingroup, 0, sizeof ctx->is_ingroup);
    for (int f = 0; f < n_files; f++) ctx->is_ingroup[f] = is_ingroup[f] ? 1 : 0;
    ctx->configured = true;
    return KB_OK;
}

int kb_clear_sequences(kb_ctx* ctx) {
    if (!ctx) return KB_EINVAL;
    ctx->n_bases = 0;
    ctx->file_starts.clear();
    ctx->file_gid.clear();
    ctx->file_event.clear();
    ctx->own_first = 0; ctx->own_count = -1;
    ctx->sep_filled = false;
    ctx->slab_off = false;
    ctx->pending_fasta.clear(); ctx->fasta_done = 0; ctx->fasta_of_file.clear(); ctx->raw_used = 0;
    if (ctx->fa_flags.p) { cudaSetDevice(ctx->device); cudaMemsetAsync(ctx->fa_flags.p, 0, 8, ctx->stream); }
    return KB_OK;
}

static size_t padded_len(uint64_t n_bases) {
    return (size_t)((n_bases + KB_K1_TB - 1) / KB_K1_TB) * KB_K1_TB + KB_K1_PAD;
}

int kb_reserve(kb_ctx* ctx, uint64_t total_bytes) {
    if (!ctx) return KB_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    ctx->reserve_hint = total_bytes + KB_MAX_FILES;
    ctx->sep_filled = false;
    return ensure(ctx, ctx->bases, padded_len(total_bytes + KB_MAX_FILES), true);
}

int kb_add_sequence(kb_ctx* ctx, int file_id, const uint8_t* bytes, uint64_t n_bytes, int on_device) {
    if (!ctx) return KB_EINVAL;
    if (file_id < 0 || file_id >= KB_MAX_FILES) return fail(ctx, KB_EINVAL, "file_id out of range");
    if (n_bytes && !bytes) return fail(ctx, KB_EINVAL, "null sequence pointer");
    CU(cudaSetDevice(ctx->device));
    if (ctx->bases.cap < padded_len(ctx->n_bases + n_bytes + 1)) {
        CU(cudaStreamSynchronize(ctx->copy_stream));                 // the buffer moves: no copy may be in flight
        TRY(ensure(ctx, ctx->bases, padded_len(ctx->n_bases + n_bytes + 1), true));
        ctx->sep_filled = false;
    }
    uint8_t* dst = (uint8_t*)ctx->bases.p + ctx->n_bases;
    cudaEvent_t ev = nullptr;
    if (on_device) {
        if (n_bytes) CU(cudaMemcpyAsync(dst, bytes, n_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemsetAsync(dst + n_bytes, '\n', 1, ctx->stream));      // separator between files
    } else {
        // host buffers go over the copy stream; kb_search launches K1 on the files that have arrived
        const size_t idx = ctx->file_starts.size();
        if (idx == 0 || !ctx->sep_filled || ctx->n_bases + n_bytes + 1 > ctx->sep_upto) {
            // the copy stream starts after whatever the main stream still does with the buffer.
            // The region the files will land in becomes separators ONCE, so that the copy stream carries nothing but the file
            // copies (a one-byte memset between two copies costs the DMA engine a bubble per file)
            const uint64_t upto = std::min<uint64_t>(ctx->bases.cap, std::max<uint64_t>(ctx->n_bases + n_bytes + 1 + 4096, ctx->reserve_hint));
            CU(cudaMemsetAsync(dst, '\n', upto - ctx->n_bases, ctx->stream));
            ctx->sep_filled = true; ctx->sep_upto = upto;
            CU(cudaEventRecord(ctx->main_event, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->main_event, 0));
        }
        while (ctx->copy_events.size() <= idx) {
            cudaEvent_t ne;
            CU(cudaEventCreateWithFlags(&ne, cudaEventDisableTiming));
            ctx->copy_events.push_back(ne);
        }
        ev = ctx->copy_events[idx];
        if (n_bytes) CU(cudaMemcpyAsync(dst, bytes, n_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaEventRecord(ev, ctx->copy_stream));
    }
    ctx->file_starts.push_back(ctx->n_bases);
    ctx->file_gid.push_back((uint32_t)file_id);
    ctx->file_event.push_back(ev);
    ctx->fasta_of_file.push_back(-1);
    ctx->n_bases += n_bytes + 1;
    return KB_OK;
}

static int deline_pending(kb_ctx* ctx, size_t upto);

int kb_add_fasta(kb_ctx* ctx, int file_id, const uint8_t* bytes, uint64_t n_bytes) {
    if (!ctx) return KB_EINVAL;
    if (file_id < 0 || file_id >= KB_MAX_FILES) return fail(ctx, KB_EINVAL, "file_id out of range");
    if (n_bytes && !bytes) return fail(ctx, KB_EINVAL, "null file pointer");
    CU(cudaSetDevice(ctx->device));
    // the first line decides FASTA vs plain and is consumed by the probe (kstream.py:447-452, :510-537)
    const uint8_t* nl = n_bytes ? (const uint8_t*)memchr(bytes, '\n', n_bytes) : nullptr;
    const uint64_t first = nl ? (uint64_t)(nl - bytes) + 1 : n_bytes;
    const int fasta = (n_bytes && memchr(bytes, '>', nl ? (size_t)(nl - bytes) : (size_t)n_bytes)) ? 1 : 0;
    const uint8_t* body = bytes + first;
    const uint64_t nb = n_bytes - first;
    // Asynchronous like kb_add_sequence: the raw bytes travel on the copy stream into their own region of the raw buffer (one event
    // per file); the de-lining kernels are enqueued later, per batch of arrived files, right before K1 needs the bytes
    // (deline_pending), so that nothing here waits for the device.  `bytes` must stay valid until the next synchronising call.
    if (ctx->bases.cap < padded_len(ctx->n_bases + nb + 1)) {
        CU(cudaStreamSynchronize(ctx->copy_stream));
        TRY(deline_pending(ctx, ctx->pending_fasta.size()));                  // (their slots move with the buffer: write them first)
        TRY(ensure(ctx, ctx->bases, padded_len(ctx->n_bases + nb + 1), true));
    }
    const size_t raw_off = (ctx->raw_used + 15) & ~(size_t)15;
    if (ctx->rawbuf.cap < raw_off + nb + 64) {
        CU(cudaStreamSynchronize(ctx->copy_stream));
        TRY(deline_pending(ctx, ctx->pending_fasta.size()));                  // (they read the old raw buffer)
        TRY(ensure(ctx, ctx->rawbuf, std::max<size_t>(raw_off + nb + 64, (size_t)ctx->reserve_hint + 64 * (ctx->file_starts.size() + 2)), true));
    }
    const size_t idx = ctx->file_starts.size();
    if (ctx->pending_fasta.empty()) {
        // the copy stream starts after whatever the main stream still does with the raw buffer (the previous search's de-lining)
        CU(cudaEventRecord(ctx->main_event, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->main_event, 0));
    }
    while (ctx->copy_events.size() <= idx) {
        cudaEvent_t ne;
        CU(cudaEventCreateWithFlags(&ne, cudaEventDisableTiming));
        ctx->copy_events.push_back(ne);
    }
    cudaEvent_t ev = ctx->copy_events[idx];
    if (nb) CU(cudaMemcpyAsync((uint8_t*)ctx->rawbuf.p + raw_off, body, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(ev, ctx->copy_stream));
    ctx->pending_fasta.push_back({raw_off, nb, fasta, ctx->n_bases, ev, idx});
    ctx->raw_used = raw_off + nb;
    ctx->fasta_of_file.push_back((int)ctx->pending_fasta.size() - 1);
    ctx->file_starts.push_back(ctx->n_bases);
    ctx->file_gid.push_back((uint32_t)file_id);
    ctx->file_event.push_back(ev);
    ctx->n_bases += nb + 1;
    return KB_OK;
}

int kb_fasta_flags(kb_ctx* ctx, unsigned int* flags) {
    if (!ctx || !flags) return KB_EINVAL;
    *flags = 0;
    CU(cudaSetDevice(ctx->device));
    TRY(deline_pending(ctx, ctx->pending_fasta.size()));
    if (!ctx->fa_flags.p) return KB_OK;
    CU(cudaMemcpyAsync(ctx->h_pinned + 16, ctx->fa_flags.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *flags = (unsigned int)ctx->h_pinned[16];
    return KB_OK;
}

int kb_get_sequence(kb_ctx* ctx, int local_index, uint8_t* out, uint64_t cap, uint64_t* n_bytes) {
    if (!ctx || !n_bytes) return KB_EINVAL;
    if (local_index < 0 || local_index >= (int)ctx->file_starts.size()) return fail(ctx, KB_EINVAL, "no such sequence");
    const uint64_t lo = ctx->file_starts[local_index];
    const uint64_t hi = local_index + 1 < (int)ctx->file_starts.size() ? ctx->file_starts[local_index + 1] : ctx->n_bases;
    *n_bytes = hi - lo;
    if (!out || cap < hi - lo) return KB_OK;                              // size query
    CU(cudaSetDevice(ctx->device));
    TRY(deline_pending(ctx, ctx->pending_fasta.size()));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaMemcpyAsync(out, (const uint8_t*)ctx->bases.p + lo, hi - lo, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_synchronize(kb_ctx* ctx) {
    if (!ctx) return KB_EINVAL;
    CU(cudaSetDevice(ctx->device));
    TRY(deline_pending(ctx, ctx->pending_fasta.size()));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

}  // extern "C"

// ---- stages --------------------------------------------------------------------------------------
static int upload_file_table(kb_ctx* ctx) {
    const size_t nf = ctx->file_starts.size();
    std::vector<uint64_t>& fs = ctx->file_table_host;        // lives in the ctx: the async copy needs no synchronisation
    fs = ctx->file_starts;
    fs.push_back(ctx->n_bases);
    TRY(ensure(ctx, ctx->d_file_starts, fs.size() * 8));
    TRY(ensure(ctx, ctx->d_file_gid, std::max<size_t>(nf, 1) * 4));
    CU(cudaMemcpyAsync(ctx->d_file_starts.p, fs.data(), fs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_file_gid.p, ctx->file_gid.data(), nf * 4, cudaMemcpyHostToDevice, ctx->stream));
    return KB_OK;
}

static int prepare_small(kb_ctx* ctx) {
    TRY(ensure(ctx, ctx->small, SM_TOTAL * 8));
    CU(cudaMemsetAsync(ctx->small.p, 0, SM_TOTAL * 8, ctx->stream));
    return KB_OK;
}

// De-lining kernels (kb_ingest.cuh) of the FASTA files added with kb_add_fasta, up to (not including) pending index `upto`, on the
// main stream, each behind the event of its host-to-device copy.  Called per batch of arrived files right before K1 runs on them.
static int deline_pending(kb_ctx* ctx, size_t upto) {
    upto = std::min(upto, ctx->pending_fasta.size());
    if (ctx->fasta_done >= upto) return KB_OK;
    if (!ctx->fa_flags.p) { TRY(ensure(ctx, ctx->fa_flags, 8)); CU(cudaMemsetAsync(ctx->fa_flags.p, 0, 8, ctx->stream)); }
    // the de-lining streams start after what the main stream had queued when the first file of this round came up (flag word,
    // earlier users of the buffers) — not after the K1 batches queued since: those touch other files
    const bool first_round = ctx->fasta_done == 0;
    if (first_round) CU(cudaEventRecord(ctx->main_event, ctx->stream));
    static const int n_fa = []() { const char* e = getenv("KRISP_FA_STREAMS"); const int v = e ? atoi(e) : 2; return std::max(1, std::min(v, KB_FA_STREAMS)); }();
    // the files [fasta_done, upto) have arrived together (one batch): they share the three launches, KB_FA_MAXF files at a time
    while (ctx->fasta_done < upto) {
        const size_t i0 = ctx->fasta_done, i1 = std::min(upto, i0 + (size_t)KB_FA_MAXF);
        const int k = (int)(ctx->fa_round++ % (size_t)n_fa);
        if (!ctx->fa_stream[k]) CU(cudaStreamCreateWithFlags(&ctx->fa_stream[k], cudaStreamNonBlocking));
        cudaStream_t st = ctx->fa_stream[k];
        while (ctx->fa_done_ev.size() <= i0) {
            cudaEvent_t ne;
            CU(cudaEventCreateWithFlags(&ne, cudaEventDisableTiming));
            ctx->fa_done_ev.push_back(ne);
        }
        if (first_round) CU(cudaStreamWaitEvent(st, ctx->main_event, 0));
        CU(cudaStreamWaitEvent(st, ctx->pending_fasta[i1 - 1].ev, 0));          // (the copies arrive in order: the last file's event covers the batch)
        // separators wherever the packed bytes will not reach: the slots of the batch are contiguous
        {
            const kb_ctx::PendingFasta& fa = ctx->pending_fasta[i0];
            const kb_ctx::PendingFasta& fb = ctx->pending_fasta[i1 - 1];
            CU(cudaMemsetAsync((uint8_t*)ctx->bases.p + fa.slot_off, '\n', fb.slot_off + fb.nb + 1 - fa.slot_off, st));
        }
        KbFastaBatch b{};
        size_t words = 0;
        uint32_t tiles_total = 0;
        for (size_t i = i0; i < i1; i++) {
            const kb_ctx::PendingFasta& f = ctx->pending_fasta[i];
            if (!f.nb) continue;
            const uint32_t tiles = (uint32_t)((f.nb + KB_FA_TILE - 1) / KB_FA_TILE);
            words += (size_t)tiles * 4 + 2 + (tiles + 7) / 8 + 8;                // last_nl | counts x 2 | start (+1) | hdr0 bytes
        }
        DevBuf& wk = ctx->fa_work_s[k];
        if (wk.cap < words * 8) {                                                 // (rare: grows to the largest batch; the stream's earlier chain is done first)
            CU(cudaStreamSynchronize(st));
            if (wk.p) CU(cudaFree(wk.p));
            wk.p = nullptr; wk.cap = 0;
            CU(cudaMalloc(&wk.p, words * 8 + words));
            wk.cap = words * 8 + words;
        }
        unsigned long long* q = (unsigned long long*)wk.p;
        for (size_t i = i0; i < i1; i++) {
            const kb_ctx::PendingFasta& f = ctx->pending_fasta[i];
            if (!f.nb) continue;
            const uint32_t tiles = (uint32_t)((f.nb + KB_FA_TILE - 1) / KB_FA_TILE);
            KbFastaArgs& a = b.f[b.n_files];
            a.in = (const uint8_t*)ctx->rawbuf.p + f.raw_off; a.n = f.nb; a.fasta = f.fasta;
            a.last_nl = q; q += tiles;
            a.counts = q; q += 2 * (size_t)tiles;
            b.start_w[b.n_files] = q; a.start = q; q += tiles + 1;
            a.hdr0 = (uint8_t*)q; q += (tiles + 7) / 8 + 1;
            a.out = (uint8_t*)ctx->bases.p + f.slot_off;
            a.flags = (unsigned int*)ctx->fa_flags.p;
            b.tile0[b.n_files] = tiles_total;
            tiles_total += tiles;
            b.n_files++;
            b.tile0[b.n_files] = tiles_total;
        }
        if (b.n_files) {
            kb_fa_count_kernel<<<tiles_total, KB_FA_THREADS, 0, st>>>(b);
            CU(cudaGetLastError());
            kb_fa_offsets_kernel<<<(unsigned)b.n_files, 1024, 0, st>>>(b);
            CU(cudaGetLastError());
            kb_fa_pack_kernel<<<tiles_total, KB_FA_THREADS, 0, st>>>(b);
            CU(cudaGetLastError());
            ctx->launches += 3;
        }
        CU(cudaEventRecord(ctx->fa_done_ev[i0], st));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->fa_done_ev[i0], 0));             // whatever the main stream does next sees these files
        ctx->fasta_done = i1;
    }
    return KB_OK;
}

struct KBatch {
    uint32_t first;          // end tile
    cudaEvent_t second;      // event to wait for (null: none)
    size_t last_file;        // local files [.., last_file] are complete once the event fired
};

// de-lining of every FASTA file up to local file `last_file` (kb_add_fasta defers it to here)
static int deline_upto_file(kb_ctx* ctx, size_t last_file) {
    size_t upto = ctx->fasta_done;
    while (upto < ctx->pending_fasta.size() && ctx->pending_fasta[upto].file_idx <= last_file) upto++;
    return deline_pending(ctx, upto);
}

// Batches of arriving files (host buffers on the copy stream): about ten, each ending at a file boundary.  Uniform on purpose — per
// file the GPU's work (K1 + level 1 of the batch, ~80 us per 5 Mbp) is about what its copy takes (~92 us), so the last batch's work
// after the end of the copy is what is exposed; uneven schedules (small first / last batches) were measured and lose: the big
// batches in the middle make the GPU wait for data it could already have worked on (parsed sequences 6.36 -> 6.64 ms).
static std::vector<KBatch> extract_batches(kb_ctx* ctx, uint32_t tile0, uint32_t n_tiles) {
    std::vector<KBatch> batches;
    const size_t nf = ctx->file_starts.size();
    const uint32_t t_end = tile0 + n_tiles;
    const uint64_t step = std::max<uint64_t>(ctx->n_bases / 10, 8ull << 20);
    uint64_t next = step;
    for (size_t f = 0; f < nf; f++) {
        if (!ctx->file_event[f]) continue;
        const uint64_t resident = f + 1 < nf ? ctx->file_starts[f + 1] : ctx->n_bases;   // bytes [0, resident) are there once event f fired
        if (f + 1 < nf && resident < next) continue;
        next = resident + step;
        uint32_t t1 = f + 1 < nf ? (uint32_t)std::min<uint64_t>(t_end, resident > KB_K1_PAD ? (resident - KB_K1_MAXHALO - 64) / KB_K1_TB : 0) : t_end;
        if (f + 1 == nf) { batches.push_back({t_end, ctx->file_event[f], f}); break; }
        if (t1 > tile0 && (batches.empty() || t1 > batches.back().first)) batches.push_back({t1, ctx->file_event[f], f});
    }
    if (batches.empty() || batches.back().first < t_end) {
        // device-resident sequences (or none pending): everything at once, after every pending copy
        cudaEvent_t last = nullptr;
        for (size_t f = 0; f < nf; f++) if (ctx->file_event[f]) last = ctx->file_event[f];
        batches.push_back({t_end, last, nf ? nf - 1 : 0});
    }
    return batches;
}

// K1 over tiles [tile0, tile0+n_tiles), windows starting in [pos_lo, pos_hi); *n_out = records written
static int run_extract(kb_ctx* ctx, const KbLayout& lo, uint32_t tile0, uint32_t n_tiles, uint64_t pos_lo, uint64_t pos_hi, uint64_t* n_out,
                       unsigned long long* hist = nullptr, uint32_t hist_shift = 0, uint32_t hist_bits = 0, bool no_sync = false,
                       BatchL0* bl = nullptr) {
    const uint64_t n_max = 2 * (std::min<uint64_t>(pos_hi, ctx->n_bases) - pos_lo) + 64;
    TRY(ensure(ctx, ctx->entA, (n_max + 2048) * 8));   // slack: bulk copies of the stream kernel read whole 4 KB stages
    if (!lo.direct && !ctx->lazy_now) TRY(ensure(ctx, ctx->recs, n_max * 8 * lo.W));
    // separator padding past the data
    const size_t plen = padded_len(ctx->n_bases);
    TRY(ensure(ctx, ctx->bases, plen, true));
    CU(cudaMemsetAsync((uint8_t*)ctx->bases.p + ctx->n_bases, '\n', plen - ctx->n_bases, ctx->stream));
    TRY(upload_file_table(ctx));

    KbExtractArgs a{};
    a.bases = (const uint8_t*)ctx->bases.p;
    a.n_bases = ctx->n_bases;
    a.file_starts = (const uint64_t*)ctx->d_file_starts.p;
    a.file_gid = (const uint32_t*)ctx->d_file_gid.p;
    a.n_local_files = (int)ctx->file_gid.size();
    a.soft_omit = ctx->soft_mode;
    a.lo = lo;
    a.out_entries = (uint64_t*)ctx->entA.p;
    a.out_recs = (uint64_t*)ctx->recs.p;
    a.lazy = ctx->lazy_now ? 1 : 0;
    a.strand_mode = ctx->strands_now;
    a.n_out = (unsigned long long*)ctx->small.p + SM_NOUT;
    a.tile0 = tile0; a.n_tiles = n_tiles;
    a.pos_lo = pos_lo; a.pos_hi = pos_hi;
    a.hist = hist; a.hist_shift = hist_shift; a.hist_bits = hist_bits;
    const bool wide_hist = hist && hist_bits > 9;            // 4-group CTAs with the packed shared-memory histogram (<= 16 bits)
    const size_t smem = wide_hist ? kb_extract_smem(lo.k, 4, (int)hist_bits) : kb_extract_smem(lo.k);
    prof_begin(ctx, "K1 extract");
    std::vector<KBatch> batches = extract_batches(ctx, tile0, n_tiles);
    uint32_t t0 = tile0;
    const bool batched_l0 = bl && bl->enabled && wide_hist && batches.size() > 1 && batches.size() <= KB_MAX_BATCHES;
    if (batched_l0) CU(cudaFuncSetAttribute(kb_part_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb_part_smem()));
    for (auto& bt : batches) {
        if (bt.second) CU(cudaStreamWaitEvent(ctx->stream, bt.second, 0));
        TRY(deline_upto_file(ctx, bt.last_file));
        const uint32_t nt = bt.first > t0 ? bt.first - t0 : 0;
        if (!nt) continue;
        a.tile0 = t0; a.n_tiles = nt;
        if (batched_l0) a.hist = bl->hist2 + (size_t)bl->n_batches * bl->nc1;
        if (wide_hist) {
            const uint32_t grid = std::min<uint32_t>((nt + 3) / 4, (uint32_t)ctx->n_sm);
            switch (lo.W) {
                case 1: CU(cudaFuncSetAttribute(kb_extract_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        kb_extract_kernel<1, 4><<<grid, KB_K1_THREADS * 4, smem, ctx->stream>>>(a); break;
                case 2: CU(cudaFuncSetAttribute(kb_extract_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        kb_extract_kernel<2, 4><<<grid, KB_K1_THREADS * 4, smem, ctx->stream>>>(a); break;
                case 4: CU(cudaFuncSetAttribute(kb_extract_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        kb_extract_kernel<4, 4><<<grid, KB_K1_THREADS * 4, smem, ctx->stream>>>(a); break;
                default: CU(cudaFuncSetAttribute(kb_extract_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        kb_extract_kernel<8, 4><<<grid, KB_K1_THREADS * 4, smem, ctx->stream>>>(a); break;
            }
        } else {
            const uint32_t grid = std::min<uint32_t>(nt, (uint32_t)ctx->n_sm * 8);
            switch (lo.W) {
                case 1: kb_extract_kernel<1, 1><<<grid, KB_K1_THREADS, smem, ctx->stream>>>(a); break;
                case 2: kb_extract_kernel<2, 1><<<grid, KB_K1_THREADS, smem, ctx->stream>>>(a); break;
                case 4: kb_extract_kernel<4, 1><<<grid, KB_K1_THREADS, smem, ctx->stream>>>(a); break;
                default: kb_extract_kernel<8, 1><<<grid, KB_K1_THREADS, smem, ctx->stream>>>(a); break;
            }
        }
        CU(cudaGetLastError());
        ctx->launches++;
        if (batched_l0) {
            // level 0 of this batch right away: its records are the tail of the element array
            const int b = bl->n_batches++;
            unsigned long long* c0 = bl->counts_all + (size_t)b * bl->nc0;
            kb_batch_fold_kernel<<<bl->nc0, 512, 0, ctx->stream>>>(a.hist, bl->total2, c0, 1u << bl->bits1);
            CU(cudaGetLastError());
            kb_batch_plan_kernel<<<1, 512, 0, ctx->stream>>>(c0, bl->nc0, bl->start_all + (size_t)b * bl->nc0, bl->start_all + (size_t)b * bl->nc0,
                                                            bl->cursors + (size_t)b * bl->nc0, bl->roots + 2 * b, bl->roottiles + 2 * b);
            CU(cudaGetLastError());
            KbPartArgs pa{};
            pa.in = (const uint64_t*)ctx->entA.p; pa.out = bl->out;
            pa.pstart = bl->roots + 2 * b; pa.ptile0 = bl->roottiles + 2 * b; pa.n_parents = 1;
            pa.shift = 64 - bl->bits0; pa.bits = bl->bits0;
            pa.cursor = bl->cursors + (size_t)b * bl->nc0; pa.hist = pa.cursor;
            const uint64_t bound = 2ull * nt * KB_K1_TB / KB_PT_TILE + 2;
            kb_part_kernel<2><<<(unsigned)bound, KB_PT_THREADS, kb_part_smem(), ctx->stream>>>(pa);
            CU(cudaGetLastError());
            ctx->launches += 3;
        }
        t0 = bt.first;
    }
    prof_end(ctx);
    if (no_sync) {                                         // the count stays on the device; the caller sizes grids with the bound
        *n_out = n_max;
        ctx->alg_bytes += std::min<uint64_t>(pos_hi, ctx->n_bases) - pos_lo;
        ctx->alg_rec_bytes += 8 * ((lo.direct || ctx->lazy_now) ? 1 : (1 + lo.W));
        return KB_OK;
    }
    CU(cudaMemcpyAsync(ctx->h_pinned, (uint64_t*)ctx->small.p + SM_NOUT, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *n_out = ctx->h_pinned[0];
    if (*n_out > n_max) return fail(ctx, KB_EINTERNAL, "K1 wrote more records than the bound");
    ctx->alg_bytes += (std::min<uint64_t>(pos_hi, ctx->n_bases) - pos_lo) + *n_out * 8 * (lo.direct ? 1 : (1 + lo.W));
    return KB_OK;
}

template <typename ST, int THREADS, int ITEMS, int MINB, bool HWMATCH>
static int launch_pass_v(kb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t n, uint32_t shift, uint32_t shard_n, int hist_row, int ticket_idx) {
    constexpr uint64_t TILE = (uint64_t)THREADS * ITEMS;
    const uint64_t n_tiles = (n + TILE - 1) / TILE;
    TRY(ensure(ctx, ctx->status, n_tiles * KB_RADIX * sizeof(ST)));
    CU(cudaMemsetAsync(ctx->status.p, 0, n_tiles * KB_RADIX * sizeof(ST), ctx->stream));
    KbSortArgs<ST> a{};
    a.in = in; a.out = out; a.n = n; a.shift = shift; a.shard_n = shard_n;
    a.base = (const unsigned long long*)ctx->small.p + SM_HIST + (size_t)hist_row * KB_RADIX;
    a.status = (ST*)ctx->status.p;
    a.ticket = (uint32_t*)((uint64_t*)ctx->small.p + SM_TICKET) + ticket_idx;
    const size_t smem = kb_onesweep_smem<THREADS, ITEMS>();
    CU(cudaFuncSetAttribute(kb_onesweep_kernel<ST, THREADS, ITEMS, MINB, HWMATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kb_onesweep_kernel<ST, THREADS, ITEMS, MINB, HWMATCH><<<(unsigned)n_tiles, THREADS, smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
    return KB_OK;
}

template <typename ST>
static int launch_pass(kb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t n, uint32_t shift, uint32_t shard_n, int hist_row, int ticket_idx) {
    switch (ctx->opt_sort_variant) {
        case 1: return launch_pass_v<ST, 256, 16, 3, false>(ctx, in, out, n, shift, shard_n, hist_row, ticket_idx);
        case 2: return launch_pass_v<ST, 512, 16, 2, true>(ctx, in, out, n, shift, shard_n, hist_row, ticket_idx);
        case 3: return launch_pass_v<ST, 256, 16, 3, true>(ctx, in, out, n, shift, shard_n, hist_row, ticket_idx);
        case 4: return launch_pass_v<ST, 512, 12, 2, true>(ctx, in, out, n, shift, shard_n, hist_row, ticket_idx);
        case 5: return launch_pass_v<ST, 256, 20, 3, true>(ctx, in, out, n, shift, shard_n, hist_row, ticket_idx);
        default: return launch_pass_v<ST, 512, 16, 2, false>(ctx, in, out, n, shift, shard_n, hist_row, ticket_idx);
    }
}

// sort in[0..n) by the top 8P bits, ping-ponging with `other` (same capacity) -> *sorted = buffer holding the result
static int run_sort(kb_ctx* ctx, DevBuf& in, DevBuf& other, uint64_t n, int P, uint64_t** sorted) {
    uint64_t* cur = (uint64_t*)in.p;
    *sorted = cur;
    if (n == 0 || P == 0) return KB_OK;
    if (P > 8) return fail(ctx, KB_EINTERNAL, "more than 8 radix passes");
    TRY(ensure(ctx, other, (size_t)(n + 2048) * 8));        // (not in.cap: capacities carry growth slack and would chase each other)
    uint64_t* alt = (uint64_t*)other.p;
    const bool wide = n >= (1ULL << 30);
    const uint32_t shift0 = 64 - 8 * P;

    prof_begin(ctx, "K2 histogram");
    KbHistArgs h{};
    h.in = cur; h.n = n; h.P = P; h.shift0 = shift0; h.shard_n = 0;
    h.hist = (unsigned long long*)ctx->small.p + SM_HIST;
    const unsigned hgrid = (unsigned)std::min<uint64_t>((n / 2 + KB_HIST_THREADS - 1) / KB_HIST_THREADS + 1, (uint64_t)ctx->n_sm * 4);
    kb_hist_kernel<<<hgrid, KB_HIST_THREADS, 0, ctx->stream>>>(h);
    CU(cudaGetLastError());
    kb_scan_kernel<<<P, KB_RADIX, 0, ctx->stream>>>(h.hist);
    CU(cudaGetLastError());
    ctx->launches += 2;
    prof_end(ctx);
    ctx->alg_bytes += n * 8;

    for (int p = 0; p < P; p++) {
        static const char* names[8] = {"K2 pass 0", "K2 pass 1", "K2 pass 2", "K2 pass 3", "K2 pass 4", "K2 pass 5", "K2 pass 6", "K2 pass 7"};
        prof_begin(ctx, names[p]);
        if (wide) TRY(launch_pass<unsigned long long>(ctx, cur, alt, n, shift0 + 8 * p, 0, p, p));
        else TRY(launch_pass<uint32_t>(ctx, cur, alt, n, shift0 + 8 * p, 0, p, p));
        prof_end(ctx);
        std::swap(cur, alt);
        ctx->alg_bytes += n * 16;
    }
    ctx->passes = P;
    *sorted = cur;
    return KB_OK;
}


// ---- search path: radix partition (kb_part.cuh) + bucket hash aggregation (kb_hash.cuh) ----------------
static bool hash_fast_ok(const kb_ctx* ctx) {
    const KbLayout& lo = ctx->lo;
    return ctx->opt_fast_group && lo.direct && lo.FB >= 1 && lo.MW <= 1 && lo.n_files <= 64;
}

// n_est: records over the whole key space (all GPUs).  bits0 > 0 (multi-GPU): level 0 — the exchange — separates exactly bits0 bits
// (>= log2 of the shard count: it decides the owner; few digits = long runs = efficient peer stores), the levels after the exchange
// share the remaining bucket bits.
static PartPlan make_plan(const kb_ctx* ctx, uint64_t n_est, int bits0 = 0, bool use_hint = true) {
    const KbLayout& lo = ctx->lo;
    PartPlan pl;
    pl.fast = hash_fast_ok(ctx);
    pl.stream = ctx->opt_hash_stream && ctx->opt_fast_group && lo.direct && lo.FB >= 1 && lo.MW <= 1;
    if (ctx->opt_hash_slots_log2) pl.slots_log2 = (uint32_t)ctx->opt_hash_slots_log2;
    else if (pl.stream) pl.slots_log2 = ctx->opt_hash_warp ? (lo.n_files <= 64 ? 11 : 10) : 10;   // kb_hash_warp.cuh, one table per CTA (32 - 40 KB,
                                                                // fewer bucket bits = cheaper partition levels) / stream kernel: 24 - 56 KB of table + 28 KB ring and queues
    else if (pl.fast) pl.slots_log2 = 11;                       // 48 KB of table: 4 CTAs per SM
    else {
        const size_t sb = kb_hash_slot_bytes(lo);
        uint32_t l2 = 11;
        while (l2 > 6 && ((size_t)1 << l2) * sb > 96 * 1024) l2--;
        pl.slots_log2 = l2;
    }
    const int keybits = lo.direct ? lo.FB : 32;
    int bb;
    if (ctx->opt_bucket_bits >= 0) bb = (int)ctx->opt_bucket_bits;
    else {
        // records per bucket: half a table of distinct keys, each expected in a good part of the files (every file for small panels,
        // half of them — at most 24 — for large ones: 19..31 occurrences per key on the 40..200-genome panels of BASELINE config 5).
        // Optimistic on purpose: a panel with fewer shared keys is caught by the re-plan of kb_search
        const int nf = std::max(lo.n_files, 1);
        const uint64_t mult = (uint64_t)(nf <= 12 ? nf : std::min(std::max(nf / 2, 12), 24));
        const uint64_t target = ((uint64_t)1 << pl.slots_log2) / 2 * mult;
        bb = 0;
        while (bb < 24 && (n_est >> bb) > target) bb++;
    }
    if (ctx->opt_bucket_bits < 0 && use_hint) bb += ctx->bb_extra;      // (never in the sharded plans: every rank must derive the same one)
    bb = std::min(bb, std::min(keybits, 24));
    if (bits0) {
        bb = std::min(std::max(bb, bits0 + 1), std::min(keybits, bits0 + 18));
        const int rb = bb - bits0, rl = (rb + 8) / 9;
        pl.bb = bb;
        pl.levels = 1 + rl;
        pl.bits[0] = bits0;
        for (int l = 1; l < pl.levels; l++) pl.bits[l] = rb / rl + (l - 1 < rb % rl ? 1 : 0);
    } else {
        pl.bb = bb;
        pl.levels = (bb + 8) / 9;
        for (int l = 0; l < pl.levels; l++) pl.bits[l] = bb / pl.levels + (l < bb % pl.levels ? 1 : 0);
        if (pl.levels == 2 && bb == 17) { pl.bits[0] = 8; pl.bits[1] = 9; }     // a 256-way level runs faster than a 512-way one: give level 0,
                                                                                  // which cannot use the pair trick, the narrow digit
    }
    const uint64_t max_tiles = n_est / KB_PT_TILE + ((uint64_t)1 << bb) + 2;
    size_t off = 0;
    int acc = 0;
    for (int l = 0; l < pl.levels; l++) {
        acc += pl.bits[l];
        pl.nc[l] = 1u << acc;
        pl.ncl[l] = pl.nc[l];
        pl.off_cnt[l] = off; off += (size_t)pl.nc[l] * 8;
        pl.off_start[l] = off; off += ((size_t)pl.nc[l] + 1) * 8;
        pl.off_tile0[l] = off; off += (((size_t)pl.nc[l] + 1) * 4 + 7) & ~(size_t)7;
    }
    pl.off_part = off; off += ((((size_t)1 << bb) + KB_PLAN_BLOCK - 1) / KB_PLAN_BLOCK + 1) * 16;
    pl.off_tilemap = off; off += (max_tiles * 4 + 7) & ~(size_t)7;
    pl.off_pair = off; off += ((size_t)2 << 16) * 8 + 64;
    pl.bytes = off;
    return pl;
}

// Partition levels [l_begin, l_end) of `in` (at most n elements) -> *parted, child table of the last level -> *bstart / *n_buckets.
// l_begin == 0: the exact element count is K1's device counter and the level-0 histogram is already in the plan buffer;
// l_begin > 0 needs `cp` (the parents of that level).
static int launch_plan(kb_ctx* ctx, KbPlanArgs pa, const PartPlan& pl) {
    if (!pa.part) pa.part = (unsigned long long*)((uint8_t*)ctx->plan.p + pl.off_part);
    const uint32_t nb = (pa.nc + KB_PLAN_BLOCK - 1) / KB_PLAN_BLOCK;
    kb_plan_reduce_kernel<<<nb, KB_PLAN_BLOCK, 0, ctx->stream>>>(pa);
    CU(cudaGetLastError());
    kb_plan_scan_kernel<<<1, KB_PLAN_BLOCK, 0, ctx->stream>>>(pa.part, nb);
    CU(cudaGetLastError());
    kb_plan_apply_kernel<<<nb, KB_PLAN_BLOCK, 0, ctx->stream>>>(pa);
    CU(cudaGetLastError());
    ctx->launches += 3;
    return KB_OK;
}

// have_hist1: K1 already counted the level-1 children (plan buffer, level-1 counts); level 0's counts are their row sums.
static int run_partition(kb_ctx* ctx, const PartPlan& pl, DevBuf& in, DevBuf& other, uint64_t n, int l_begin, int l_end,
                         const CustomParents* cp, uint64_t** parted, const unsigned long long** bstart, uint32_t* n_buckets,
                         bool have_hist1 = false, bool counts_given = false, bool pair = false) {
    uint64_t* cur = (uint64_t*)in.p;
    uint8_t* P = (uint8_t*)ctx->plan.p;
    unsigned long long* root = (unsigned long long*)ctx->small.p + SM_ROOT;
    uint32_t* roottile = (uint32_t*)((uint64_t*)ctx->small.p + SM_ROOTTILE);
    // root parent {0, n}, {0, tiles} from the device-resident record counter (n is only an upper bound here: no host round trip)
    kb_root_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long*)ctx->small.p + SM_NOUT, root, roottile);
    CU(cudaGetLastError());
    ctx->launches++;
    *parted = cur; *bstart = root; *n_buckets = 1;
    if (l_end <= l_begin || n == 0) return KB_OK;
    TRY(ensure(ctx, other, (size_t)(n + 2048) * 8));        // (not in.cap: capacities carry growth slack and would chase each other)
    uint64_t* alt = (uint64_t*)other.p;
    const size_t smem = kb_part_smem();
    CU(cudaFuncSetAttribute(kb_part_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int shift = 64;
    for (int l = 0; l < l_begin; l++) shift -= pl.bits[l];
    static const char* pnames[3] = {"K2 partition 0", "K2 partition 1", "K2 partition 2"};
    static const char* hnames[3] = {"K2 plan 0", "K2 histogram 1", "K2 histogram 2"};
    for (int l = l_begin; l < l_end; l++) {
        shift -= pl.bits[l];
        KbPartArgs a{};
        a.in = cur; a.out = alt;
        a.shift = (uint32_t)shift; a.bits = (uint32_t)pl.bits[l];
        a.cursor = (unsigned long long*)(P + pl.off_cnt[l]);
        a.hist = a.cursor;
        uint64_t grid;
        if (l == 0) {
            a.pstart = root; a.ptile0 = roottile; a.tile_parent = nullptr; a.n_parents = 1;
            grid = (n + KB_PT_TILE - 1) / KB_PT_TILE;
        } else if (l == l_begin) {
            a.pstart = cp->pstart; a.ptile0 = cp->ptile0; a.prow = cp->prow; a.n_parents = cp->n;
            a.tile_parent = (const uint32_t*)(P + pl.off_tilemap);
            grid = cp->max_tiles;
        } else {
            a.pstart = (const unsigned long long*)(P + pl.off_start[l - 1]);
            a.ptile0 = (const uint32_t*)(P + pl.off_tile0[l - 1]);
            a.tile_parent = (const uint32_t*)(P + pl.off_tilemap);
            a.n_parents = pl.ncl[l - 1];
            grid = n / KB_PT_TILE + pl.ncl[l - 1] + 1;
        }
        prof_begin(ctx, hnames[l]);
        if (l == 0 && pair) {
            // K1 counted 16 bits, level 0 + 1 have 17: the counts are those of sibling pairs.  Pair offsets -> the cursors of both
            // siblings (left end / right end) and the even entries of the bucket table; row sums = the level-0 counts
            unsigned long long* h16 = (unsigned long long*)(P + pl.off_pair);
            unsigned long long* s16 = h16 + ((size_t)1 << 16);
            const uint32_t n_pairs = pl.ncl[1] / 2;
            KbPlanArgs pp{};
            pp.counts = h16; pp.nc = n_pairs; pp.start = s16; pp.cursor = nullptr; pp.tile0 = nullptr;
            pp.folded = a.cursor; pp.fold = 1u << (pl.bits[1] - 1);
            TRY(launch_plan(ctx, pp, pl));
            kb_pair_expand_kernel<<<(n_pairs + 255) / 256, 256, 0, ctx->stream>>>(s16, n_pairs, (unsigned long long*)(P + pl.off_cnt[1]),
                                                                                (unsigned long long*)(P + pl.off_start[1]));
            CU(cudaGetLastError());
            ctx->launches++;
        } else if (l == 0 && have_hist1) {
            // level-1 offsets straight from K1's two-level histogram; its row sums are the level-0 counts
            KbPlanArgs p1{};
            p1.counts = (unsigned long long*)(P + pl.off_cnt[1]); p1.nc = pl.ncl[1];
            p1.start = (unsigned long long*)(P + pl.off_start[1]);
            p1.cursor = (unsigned long long*)(P + pl.off_cnt[1]);
            p1.tile0 = (uint32_t*)(P + pl.off_tile0[1]);
            p1.folded = a.cursor; p1.fold = 1u << pl.bits[1];
            TRY(launch_plan(ctx, p1, pl));
        }
        if (l == 1 && have_hist1) {
            kb_tilemap_kernel<<<(unsigned)std::min<uint32_t>((a.n_parents + 7) / 8, 4096), 256, 0, ctx->stream>>>(a.ptile0, a.n_parents, (uint32_t*)(P + pl.off_tilemap));
            CU(cudaGetLastError());
            ctx->launches++;
        } else if (l > 0) {
            kb_tilemap_kernel<<<(unsigned)std::min<uint32_t>((a.n_parents + 7) / 8, 4096), 256, 0, ctx->stream>>>(a.ptile0, a.n_parents, (uint32_t*)(P + pl.off_tilemap));
            CU(cudaGetLastError());
            ctx->launches++;
            if (!(l == l_begin && counts_given)) {           // (multi-GPU: the child counts of the first level came with the exchange)
                kb_part_hist_kernel<<<(unsigned)grid, KB_PT_THREADS, 0, ctx->stream>>>(a);
                CU(cudaGetLastError());
                ctx->launches++;
                ctx->alg_rec_bytes += 8;
            }
        }
        if (!(l == 1 && have_hist1)) {
            KbPlanArgs pa{};
            pa.counts = a.cursor; pa.nc = pl.ncl[l];
            pa.start = (unsigned long long*)(P + pl.off_start[l]);
            pa.cursor = a.cursor;
            pa.tile0 = (uint32_t*)(P + pl.off_tile0[l]);
            TRY(launch_plan(ctx, pa, pl));
        }
        prof_end(ctx);
        prof_begin(ctx, pnames[l]);
        a.pair_mode = (pair && l == 1) ? 1 : 0;
        if (a.pair_mode) {
            CU(cudaFuncSetAttribute(kb_part_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kb_part_kernel<2, true><<<(unsigned)grid, KB_PT_THREADS, smem, ctx->stream>>>(a);
        } else {
            kb_part_kernel<2><<<(unsigned)grid, KB_PT_THREADS, smem, ctx->stream>>>(a);
        }
        CU(cudaGetLastError());
        ctx->launches++;
        if (a.pair_mode) {
            kb_pair_fix_kernel<<<(pl.ncl[1] / 2 + 255) / 256, 256, 0, ctx->stream>>>(a.cursor, pl.ncl[1] / 2, (unsigned long long*)(P + pl.off_start[1]));
            CU(cudaGetLastError());
            ctx->launches++;
        }
        prof_end(ctx);
        ctx->alg_rec_bytes += 16;
        std::swap(cur, alt);
    }
    ctx->passes += l_end - l_begin;
    *parted = cur;
    *bstart = (const unsigned long long*)(P + pl.off_start[l_end - 1]);
    *n_buckets = pl.ncl[l_end - 1];
    return KB_OK;
}

struct HashStage {            // what run_group needs to run the bucket-hash kernels instead of the sorted-run kernels
    const PartPlan* pl;
    const unsigned long long* bstart;
    uint32_t n_buckets;
    const unsigned long long* brun = nullptr;   // lazy records: (start, length) per bucket instead of bstart
    uint32_t slots_log2 = 0;                     // != 0: table size of the exact pass (else the plan's)
    uint64_t n_kept = 0;                         // lazy records: elements the exact pass reads
    const unsigned long long* bend = nullptr;   // slab layout: fill level (absolute end) of every bucket
    uint64_t bcap = 0;                           // slab layout: bucket b starts at b * bcap
    int bb_hash = -1;                            // >= 0: key bits the buckets already separate, as the hash kernels see them (symmetric level 0:
                                                 // only the levels after it eat top bits of the mixed key); -1 = the plan's bucket bits
    int phase = 0;                               // 0 = everything; 1 = only the main kernel, on buckets [bucket0, bucket0 + n_buckets);
    uint32_t bucket0 = 0;                        // 2 = only what follows it (deferred buckets, group sizes) over all n_buckets
};

// group sizes of the survivors (kb_result_view.group_size): the rows do not need them, the alignment renderer does
static bool group_sizes_on(const kb_ctx* ctx) { return ctx->opt_want_records != 0 || ctx->opt_group_sizes != 0; }

static int launch_hash(kb_ctx* ctx, const KbGroupArgs& g, const HashStage& hs) {
    const KbLayout& lo = g.lo;
    KbHashArgs x{};
    x.g = g;
    x.bstart = hs.bstart; x.n_buckets = hs.n_buckets; x.slots_log2 = hs.slots_log2 ? hs.slots_log2 : hs.pl->slots_log2;
    x.bb = (uint32_t)(hs.bb_hash >= 0 ? hs.bb_hash : hs.pl->bb);
    x.ingroup64 = (uint64_t)g.ingroup[0] | ((uint64_t)g.ingroup[1] << 32);
    x.full64 = (uint64_t)g.full[0] | ((uint64_t)g.full[1] << 32);
    x.err = (unsigned long long*)ctx->small.p + SM_ERR;
    x.brun = hs.brun;
    x.bend = hs.bend; x.bcap = hs.bcap;
    const unsigned grid = (unsigned)std::min<uint32_t>(hs.n_buckets, 1u << 20);
    const bool warp_kernel = hs.pl->stream && (ctx->opt_hash_warp || hs.bcap);     // (the CTA-wide stream kernel needs contiguous buckets)
    if (hs.pl->stream) {
        TRY(ensure(ctx, ctx->deferred, (size_t)hs.n_buckets * 4 + 64));
        const int pwn = lo.n_files <= 64 ? 2 : (lo.n_files <= 128 ? 4 : 8);
        const bool spacer = lo.D == 1 && lo.FB == 54 && x.bb <= 22;
        if (warp_kernel && hs.phase != 2) {
            KbHWarpArgs xw{};
            xw.h = x;
            xw.h.bucket0 = hs.bucket0;
            xw.deferred = (uint32_t*)ctx->deferred.p;
            xw.n_deferred = (unsigned long long*)ctx->small.p + SM_NTAINT;   // zeroed with the result counters
            const bool packed = lo.D == 1 && lo.FB <= 54;
            bool shared = ctx->opt_hash_shared < 0 ? xw.h.slots_log2 >= 10 : ctx->opt_hash_shared != 0;
            if (!shared && xw.h.slots_log2 > 10) shared = true;                  // (a warp scans at most 32 x 32 slots)
            // warps per CTA: 8; the shared-table arrangement can run 10 (option hash_warps; 3 CTAs = 30 warps per SM)
            auto smem_nw = [&](uint32_t l2, uint32_t nw) {
                return (size_t)(shared ? kb_hash_cta_tbytes(l2, pwn, packed) : 0) + (size_t)nw * kb_hash_warp_wbytes(l2, pwn, packed, shared) + 16;
            };
            auto fit_per_sm = [&](size_t smem) { return (size_t)(228 * 1024) / (smem + 1024); };
            while (xw.h.slots_log2 > 4 && (smem_nw(xw.h.slots_log2, KB_HW_WARPS) > 224 * 1024 || xw.h.slots_log2 > 13)) xw.h.slots_log2--;
            uint32_t nw = KB_HW_WARPS;
            if (shared && ctx->opt_hash_warps == KB_HW_WARPS_WIDE) nw = KB_HW_WARPS_WIDE;   // measured on C2: 30 warps per SM 2.23 ms, 24 warps 2.11 ms
                                                                                             // (the kernel is bound by the shared-memory pipe, not by latency)
            xw.wbytes = kb_hash_warp_wbytes(xw.h.slots_log2, pwn, packed, shared);
            xw.tbytes = shared ? kb_hash_cta_tbytes(xw.h.slots_log2, pwn, packed) : 0;
            const size_t smem = smem_nw(xw.h.slots_log2, nw);
            const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(shared ? 4 : 3, fit_per_sm(smem)));
            const uint32_t units = shared ? hs.n_buckets : (hs.n_buckets + KB_HW_WARPS - 1) / KB_HW_WARPS;
            const unsigned wgrid = (unsigned)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)ctx->n_sm * per_sm, units));
#define KB_LAUNCH_WARP_(D1_, SP_, PW_, SH_, NW_)                                                                               \
            do {                                                                                                               \
                CU(cudaFuncSetAttribute(kb_hash_warp_kernel<D1_, SP_, PW_, SH_, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                kb_hash_warp_kernel<D1_, SP_, PW_, SH_, NW_><<<wgrid, 32 * NW_, smem, ctx->stream>>>(xw);                        \
            } while (0)
#define KB_LAUNCH_WARP(D1_, SP_, PW_)                                                                                          \
            do {                                                                                                               \
                if (shared && nw == KB_HW_WARPS_WIDE) KB_LAUNCH_WARP_(D1_, SP_, PW_, true, KB_HW_WARPS_WIDE);                  \
                else if (shared) KB_LAUNCH_WARP_(D1_, SP_, PW_, true, KB_HW_WARPS);                                            \
                else KB_LAUNCH_WARP_(D1_, SP_, PW_, false, KB_HW_WARPS);                                                       \
            } while (0)
            if (pwn == 2) { if (spacer) KB_LAUNCH_WARP(true, true, 2); else if (lo.D == 1) KB_LAUNCH_WARP(true, false, 2); else KB_LAUNCH_WARP(false, false, 2); }
            else if (pwn == 4) { if (spacer) KB_LAUNCH_WARP(true, true, 4); else if (lo.D == 1) KB_LAUNCH_WARP(true, false, 4); else KB_LAUNCH_WARP(false, false, 4); }
            else { if (spacer) KB_LAUNCH_WARP(true, true, 8); else if (lo.D == 1) KB_LAUNCH_WARP(true, false, 8); else KB_LAUNCH_WARP(false, false, 8); }
#undef KB_LAUNCH_WARP_
#undef KB_LAUNCH_WARP
            CU(cudaGetLastError());
            if (hs.phase == 1) { ctx->launches++; return KB_OK; }
        } else if (!warp_kernel) {
        KbHStreamArgs xs{};
        xs.h = x; xs.n_ptr = (const unsigned long long*)ctx->small.p + SM_NOUT;
        xs.deferred = (uint32_t*)ctx->deferred.p;
        xs.n_deferred = (unsigned long long*)ctx->small.p + SM_NTAINT;       // zeroed with the result counters
        const size_t smem = kb_hash_stream_smem(x.slots_log2, pwn);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (220 * 1024) / (smem + 1024)));
        const unsigned sgrid = (unsigned)std::min<uint64_t>((uint64_t)ctx->n_sm * per_sm, std::max<uint64_t>(1, g.n / 4096));
#define KB_LAUNCH_STREAM(D1_, SP_, PW_)                                                                                        \
        do {                                                                                                                   \
            CU(cudaFuncSetAttribute(kb_hash_stream_kernel<D1_, SP_, PW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            kb_hash_stream_kernel<D1_, SP_, PW_><<<sgrid, KB_HS_THREADS, smem, ctx->stream>>>(xs);                                \
        } while (0)
        if (pwn == 2) { if (spacer) KB_LAUNCH_STREAM(true, true, 2); else if (lo.D == 1) KB_LAUNCH_STREAM(true, false, 2); else KB_LAUNCH_STREAM(false, false, 2); }
        else if (pwn == 4) { if (spacer) KB_LAUNCH_STREAM(true, true, 4); else if (lo.D == 1) KB_LAUNCH_STREAM(true, false, 4); else KB_LAUNCH_STREAM(false, false, 4); }
        else { if (spacer) KB_LAUNCH_STREAM(true, true, 8); else if (lo.D == 1) KB_LAUNCH_STREAM(true, false, 8); else KB_LAUNCH_STREAM(false, false, 8); }
#undef KB_LAUNCH_STREAM
        CU(cudaGetLastError());
        }
        // fallback for the deferred buckets: the splitting kernels with a roomier table
        KbHashArgs fb = x;
        fb.list = (const uint32_t*)ctx->deferred.p; fb.n_list = (const unsigned long long*)ctx->small.p + SM_NTAINT;
        fb.abort_above = (ctx->replan_ok && hs.pl->bb < std::min(lo.FB, 24)) ? std::max<uint32_t>(hs.n_buckets / 8, 64) : 0;
        if (hs.pl->fast) {
            fb.slots_log2 = ctx->opt_hash_slots_log2 ? x.slots_log2 : std::max<uint32_t>(x.slots_log2, 11);
            const size_t fsmem = kb_hash_fast_smem(fb.slots_log2);
            if (lo.D == 1) {
                CU(cudaFuncSetAttribute(kb_hash_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                kb_hash_fast_kernel<true><<<(unsigned)ctx->n_sm * 2, KB_KH_THREADS, fsmem, ctx->stream>>>(fb);
            } else {
                CU(cudaFuncSetAttribute(kb_hash_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                kb_hash_fast_kernel<false><<<(unsigned)ctx->n_sm * 2, KB_KH_THREADS, fsmem, ctx->stream>>>(fb);
            }
        } else {                                               // > 64 files: the generic kernel
            if (!ctx->opt_hash_slots_log2) {
                uint32_t l2 = 11;
                while (l2 > 6 && ((size_t)1 << l2) * kb_hash_slot_bytes(lo) > 96 * 1024) l2--;
                fb.slots_log2 = l2;
            }
            const size_t fsmem = kb_hash_smem(lo, fb.slots_log2);
            CU(cudaFuncSetAttribute(kb_hash_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            kb_hash_kernel<1><<<(unsigned)ctx->n_sm * 2, KB_KH_THREADS, fsmem, ctx->stream>>>(fb);
        }
        CU(cudaGetLastError());
        ctx->launches += 2;
        if (group_sizes_on(ctx)) {                            // one more read of every survivor's bucket: only where the sizes are used
            KbHSizeArgs sz{};
            sz.ent = g.ent; sz.n_res = g.n_res; sz.cap = g.cap; sz.res_flank = g.res_flank; sz.res_run = g.res_run; sz.res_size = g.res_size; sz.lo = lo;
            kb_hsize_kernel<<<(unsigned)ctx->n_sm * 4, 256, 0, ctx->stream>>>(sz);
            CU(cudaGetLastError());
            ctx->launches++;
        }
        return KB_OK;
    }
    if (hs.pl->fast) {
        const size_t smem = kb_hash_fast_smem(x.slots_log2);
        if (lo.D == 1) {
            CU(cudaFuncSetAttribute(kb_hash_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kb_hash_fast_kernel<true><<<grid, KB_KH_THREADS, smem, ctx->stream>>>(x);
        } else {
            CU(cudaFuncSetAttribute(kb_hash_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kb_hash_fast_kernel<false><<<grid, KB_KH_THREADS, smem, ctx->stream>>>(x);
        }
    } else {
        const size_t smem = kb_hash_smem(lo, x.slots_log2);
        switch (lo.W) {
            case 1: CU(cudaFuncSetAttribute(kb_hash_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    kb_hash_kernel<1><<<grid, KB_KH_THREADS, smem, ctx->stream>>>(x); break;
            case 2: CU(cudaFuncSetAttribute(kb_hash_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    kb_hash_kernel<2><<<grid, KB_KH_THREADS, smem, ctx->stream>>>(x); break;
            case 4: CU(cudaFuncSetAttribute(kb_hash_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    kb_hash_kernel<4><<<grid, KB_KH_THREADS, smem, ctx->stream>>>(x); break;
            default: CU(cudaFuncSetAttribute(kb_hash_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    kb_hash_kernel<8><<<grid, KB_KH_THREADS, smem, ctx->stream>>>(x); break;
        }
    }
    CU(cudaGetLastError());
    ctx->launches++;
    return KB_OK;
}

static int ensure_results(kb_ctx* ctx, uint64_t cap) {
    const KbLayout& lo = ctx->lo;
    TRY(ensure(ctx, ctx->res_flank, cap * lo.FW * 8));
    TRY(ensure(ctx, ctx->res_in, cap * std::max(lo.MW, 1) * 4));
    TRY(ensure(ctx, ctx->res_out, cap * std::max(lo.MW, 1) * 4));
    TRY(ensure(ctx, ctx->res_size, cap * 4));
    TRY(ensure(ctx, ctx->res_run, cap * 16));
    ctx->result_cap = cap;
    return KB_OK;
}

#define KB_TAINT_CAP (1ull << 20)

static bool fast_group_ok(const kb_ctx* ctx) {
    const KbLayout& lo = ctx->lo;
    return ctx->opt_fast_group && lo.direct && lo.FB >= 1 && lo.MW <= 1 && lo.n_files <= 64;
}

// returns KB_OK and sets *fell_back when the taint list overflowed and the generic kernel must run instead
static int launch_group(kb_ctx* ctx, const KbGroupArgs& a, bool allow_fast) {
    const KbLayout& lo = a.lo;
    if (allow_fast && fast_group_ok(ctx)) {
        TRY(ensure(ctx, ctx->taint, KB_TAINT_CAP * 8));
        KbFastArgs x{};
        x.g = a;
        x.ingroup64 = (uint64_t)a.ingroup[0] | ((uint64_t)a.ingroup[1] << 32);
        x.full64 = (uint64_t)a.full[0] | ((uint64_t)a.full[1] << 32);
        x.taint = (unsigned long long*)ctx->taint.p;
        x.n_taint = (unsigned long long*)ctx->small.p + SM_NTAINT;
        x.taint_cap = KB_TAINT_CAP;
        const unsigned grid = (unsigned)((a.n + KB_K3F_TILE - 1) / KB_K3F_TILE);
        if (lo.D == 1) kb_group_fast_kernel<true><<<grid, KB_K3F_THREADS, 0, ctx->stream>>>(x);
        else kb_group_fast_kernel<false><<<grid, KB_K3F_THREADS, 0, ctx->stream>>>(x);
        CU(cudaGetLastError());
        kb_group_taint_kernel<<<(unsigned)ctx->n_sm * 16, 256, 0, ctx->stream>>>(x);   // exits at once when the list is empty
        CU(cudaGetLastError());
        ctx->launches += 2;
        return KB_OK;
    }
    const unsigned grid = (unsigned)((a.n + KB_K3_TILE - 1) / KB_K3_TILE);
    if (lo.direct) {
        if (lo.MW <= 1) kb_group_kernel<1, 1><<<grid, KB_K3_THREADS, 0, ctx->stream>>>(a);
        else kb_group_kernel<1, 4><<<grid, KB_K3_THREADS, 0, ctx->stream>>>(a);
    } else if (lo.W == 2) kb_group_kernel<2, KB_MAX_MW><<<grid, KB_K3_THREADS, 0, ctx->stream>>>(a);
    else if (lo.W == 4) kb_group_kernel<4, KB_MAX_MW><<<grid, KB_K3_THREADS, 0, ctx->stream>>>(a);
    else kb_group_kernel<8, KB_MAX_MW><<<grid, KB_K3_THREADS, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
    return KB_OK;
}

// CSV rows of the n_res survivors in the result table: order by flank words (few survivors: ranks by counting, one launch; else the
// chunked LSD sort of the indices), rendered into dev_dst (the rows section of the staging image that run_group downloads)
static int render_rows(kb_ctx* ctx, kb_result* res, uint64_t n_res, char* dev_dst, char* host_dst) {
    const KbLayout& lo = ctx->lo;
    if (!n_res || (lo.R == 0 && lo.D > 0)) return KB_OK;
    if (n_res >= (1ULL << 32)) return fail(ctx, KB_EUNSUPPORTED, "more than 2^32 survivors");
    TRY(ensure(ctx, ctx->rowkeyA, (size_t)(n_res + 2048) * 8));
    uint64_t* sorted = (uint64_t*)ctx->rowkeyA.p;
    const uint32_t* rank = nullptr;
    if (n_res <= KB_RANK_MAX && ctx->opt_rank_rows) {
        KbRankArgs rk{};
        rk.flank = (const uint64_t*)ctx->res_flank.p; rk.n = (uint32_t)n_res; rk.rank = (uint32_t*)ctx->rowkeyA.p;
        const unsigned rgrid = (unsigned)((n_res + KB_RANK_THREADS - 1) / KB_RANK_THREADS);
        const unsigned rows_y = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(32, (n_res + 255) / 256));      // CTA rows: slices of >= 256 survivors
        rk.slice = (uint32_t)((n_res + rows_y - 1) / rows_y);
        CU(cudaMemsetAsync(rk.rank, 0, (size_t)n_res * 4, ctx->stream));
        const dim3 g2(rgrid, rows_y);
        switch (lo.FW) {
            case 1: kb_rank_kernel<1><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 2: kb_rank_kernel<2><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 3: kb_rank_kernel<3><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 4: kb_rank_kernel<4><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 5: kb_rank_kernel<5><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 6: kb_rank_kernel<6><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 7: kb_rank_kernel<7><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            case 8: kb_rank_kernel<8><<<g2, KB_RANK_THREADS, 0, ctx->stream>>>(rk); break;
            default: return fail(ctx, KB_EINTERNAL, "flank wider than 8 words");
        }
        CU(cudaGetLastError());
        ctx->launches++;
        rank = rk.rank;
    } else {
        const long long prof = ctx->opt_profile;
        const int passes = ctx->passes;
        const uint64_t alg = ctx->alg_bytes;
        ctx->opt_profile = 0;                                    // (the sort below is bookkeeping on kilobytes, not a stage of the search)
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_res + 255) / 256, (uint64_t)ctx->n_sm * 8));
        kb_iota_kernel<<<grid, 256, 0, ctx->stream>>>((uint64_t*)ctx->rowkeyA.p, n_res);
        CU(cudaGetLastError());
        DevBuf* in = &ctx->rowkeyA; DevBuf* other = &ctx->rowkeyB;
        const uint32_t chunks = ((uint32_t)lo.FB + 31) / 32;
        for (int c = (int)chunks - 1; c >= 0; c--) {
            KbChunkKeyArgs ck{};
            ck.ent = (uint64_t*)in->p; ck.n = n_res; ck.recs = (const uint64_t*)ctx->res_flank.p; ck.W = (uint32_t)lo.FW;
            ck.bit_pos = 32u * (uint32_t)c; ck.nbits = std::min<uint32_t>(32, (uint32_t)lo.FB - ck.bit_pos);
            kb_chunk_key_kernel<<<grid, 256, 0, ctx->stream>>>(ck);
            CU(cudaGetLastError());
            CU(cudaMemsetAsync((uint64_t*)ctx->small.p + SM_HIST, 0, 9 * 256 * 8, ctx->stream));
            CU(cudaMemsetAsync((uint64_t*)ctx->small.p + SM_TICKET, 0, 8 * 8, ctx->stream));
            TRY(run_sort(ctx, *in, *other, n_res, 4, &sorted));
            if (sorted != (uint64_t*)in->p) std::swap(in, other);
            ctx->launches++;
        }
        ctx->opt_profile = prof; ctx->passes = passes; ctx->alg_bytes = alg;
        ctx->launches++;
    }
    const size_t bytes = (size_t)n_res * (size_t)res->row_bytes;
    KbRowsArgs ra{};
    ra.rank = rank; ra.order = sorted; ra.n = n_res; ra.flank = (const uint64_t*)ctx->res_flank.p;
    ra.in_mask = (const uint32_t*)ctx->res_in.p; ra.out_mask = (const uint32_t*)ctx->res_out.p;
    ra.L = lo.L; ra.D = lo.D; ra.R = lo.R; ra.FW = lo.FW; ra.MW = std::max(lo.MW, 1);
    ra.all_occurrences = ctx->opt_have_outgroup ? 0 : 1;
    ra.out = dev_dst;
    kb_rows_kernel<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>((bytes + 255) / 256, (uint64_t)ctx->n_sm * 16)), 256, 0, ctx->stream>>>(ra);
    CU(cudaGetLastError());
    ctx->launches++;
    res->rows = host_dst; res->rows_len = bytes;
    return KB_OK;                                            // (the caller downloads the image)
}

// K3 over sorted[0..n) + result download
static void fill_group_args(kb_ctx* ctx, KbGroupArgs& a, const uint64_t* sorted, uint64_t n) {
    const KbLayout& lo = ctx->lo;
    a.ent = sorted; a.n = n; a.recs = (const uint64_t*)ctx->recs.p; a.lo = lo;
    for (int f = 0; f < lo.n_files; f++) {
        a.full[f >> 5] |= 1u << (f & 31);
        if (ctx->is_ingroup[f]) a.ingroup[f >> 5] |= 1u << (f & 31);
    }
    a.n_res = (unsigned long long*)ctx->small.p + SM_NRES;
    a.cap = ctx->result_cap;
    a.res_flank = (uint64_t*)ctx->res_flank.p; a.res_in = (uint32_t*)ctx->res_in.p; a.res_out = (uint32_t*)ctx->res_out.p;
    a.res_size = (uint32_t*)ctx->res_size.p; a.res_run = (uint64_t*)ctx->res_run.p;
    a.stats = (unsigned long long*)ctx->small.p + SM_STATS;
}

// tail_only (pipelined multi-GPU search): the main kernel already ran (per group of buckets) on zeroed counters; only the deferred
// buckets, the group sizes and the download are left, and a survivor table that turned out too small cannot be repaired here.
static int run_group(kb_ctx* ctx, const uint64_t* sorted, uint64_t n, kb_result** out, const HashStage* hs = nullptr, bool tail_only = false) {
    // (hs != null: n is an upper bound; the exact count arrives with the first read-back)
    const KbLayout& lo = ctx->lo;
    kb_result* res = new (std::nothrow) kb_result();
    if (!res) return fail(ctx, KB_ENOMEM, "host allocation failed");
    memset(&res->v, 0, sizeof res->v);
    res->v.L = lo.L; res->v.D = lo.D; res->v.R = lo.R;
    res->v.flank_words = lo.FW; res->v.mask_words = lo.MW; res->v.record_words = lo.W;
    res->v.n_records = n;
    uint64_t n_res = 0;
    if (n > 0) {
        // (re)size for the current layout: the buffers only grow, so this is a no-op unless the configuration got wider
        { int rc = ensure_results(ctx, std::max<uint64_t>(ctx->result_cap, (uint64_t)ctx->opt_result_cap)); if (rc) { delete res; return rc; } }
        bool allow_fast = true;
        for (int attempt = 0; attempt < 4; attempt++) {
            KbGroupArgs a{};
            fill_group_args(ctx, a, sorted, n);
            cudaError_t e = tail_only ? cudaSuccess : cudaMemsetAsync((uint64_t*)ctx->small.p + SM_NRES, 0, 7 * 8, ctx->stream);
            if (e != cudaSuccess) { delete res; return fail(ctx, KB_ECUDA, cudaGetErrorString(e)); }
            prof_begin(ctx, hs ? "K3 bucket hash" : "K3 group");
            int rc = hs ? launch_hash(ctx, a, *hs) : launch_group(ctx, a, allow_fast);
            prof_end(ctx);
            if (rc) { delete res; return rc; }
            // one read-back: [0] K1's record count, [1] survivors, [2..5] stats, [6] taint / deferred count, [7] error flag
            e = cudaMemcpyAsync(ctx->h_pinned, (uint64_t*)ctx->small.p + SM_NOUT, 9 * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { delete res; return fail(ctx, KB_ECUDA, std::string("group pass: ") + cudaGetErrorString(e)); }
            n_res = ctx->h_pinned[1];
            for (int i = 0; i < 4; i++) res->v.stats[i] = ctx->h_pinned[2 + i];
            if (hs && attempt == 0) {                       // the search path ran without a host round trip: `n` was only a bound until now
                n = ctx->h_pinned[0];
                res->v.n_records = n;
                ctx->alg_bytes += n * ctx->alg_rec_bytes;
            }
            if (hs && hs->brun) ctx->alg_bytes += n * 8 + hs->n_kept * (uint64_t)(32 + 16 * lo.W + lo.k);   // K3a read, K3a/b/c on what is kept
            else ctx->alg_bytes += n * 8 * (lo.direct ? 1 : (1 + lo.W));
            if (hs && hs->bcap && ctx->h_pinned[8]) { delete res; return KB_SLABOVF; }   // a slab overflowed (records were dropped): exact path
            if (hs && ctx->h_pinned[7] == 2) { delete res; return KB_REPLAN; }     // too many deferred buckets: the caller re-plans
            if (hs && ctx->h_pinned[7]) { delete res; return fail(ctx, KB_EINTERNAL, "bucket hash: a bucket could not be resolved (hash table split limit)"); }
            if (hs && !hs->brun && !lo.direct && n >= (1ULL << 32)) { delete res; return fail(ctx, KB_EUNSUPPORTED, "more than 2^32 records per GPU in multi-word mode"); }
            if (!hs && allow_fast && fast_group_ok(ctx) && ctx->h_pinned[6] > KB_TAINT_CAP) { allow_fast = false; continue; }   // taint list overflow: generic kernel
            if (n_res <= ctx->result_cap) break;
            rc = ensure_results(ctx, n_res + n_res / 8 + 16);       // table too small: grow and re-run the pass
            if (rc) { delete res; return rc; }
            if (tail_only) { delete res; return KB_RESGROW; }
        }
        if (n_res > ctx->result_cap) { delete res; return fail(ctx, KB_EINTERNAL, "survivor table overflow"); }
    }
    res->v.n_groups = n_res;
    res->run_offset.assign(n_res + 1, 0);
    res->row_bytes = lo.L + lo.D + lo.R + 3;
    // pinned arena: flank | ingroup masks | outgroup masks | group sizes | runs (want_records, sorted path) | rows
    const size_t mw = (size_t)std::max(lo.MW, 1);
    const bool rows_on = ctx->opt_render_rows && !(lo.R == 0 && lo.D > 0);
    const bool need_runs = ctx->opt_want_records && !hs;
    const bool sizes_on = !hs || !hs->pl->stream || group_sizes_on(ctx);   // (only the streaming bucket hash needs an extra pass, kb_hsize_kernel, for them)
    auto up = [](size_t x) { return (x + 63) & ~(size_t)63; };
    const size_t o_flank = 0, o_in = o_flank + up(n_res * lo.FW * 8), o_out = o_in + up(n_res * mw * 4), o_size = o_out + up(n_res * mw * 4),
                 o_runs = o_size + up(n_res * 4), o_rows = o_runs + up(need_runs ? n_res * 16 : 0),
                 total_bytes = o_rows + up(rows_on ? n_res * (size_t)res->row_bytes : 0) + 64;
    if (total_bytes > ctx->h_arena_cap) {
        if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
        ctx->h_arena = nullptr; ctx->h_arena_cap = 0;
        const size_t want = total_bytes + total_bytes / 4 + (1 << 20);
        if (cudaHostAlloc((void**)&ctx->h_arena, want, cudaHostAllocDefault) != cudaSuccess) { delete res; return fail(ctx, KB_ENOMEM, "pinned result arena"); }
        ctx->h_arena_cap = want;
    }
    uint8_t* A = ctx->h_arena;
    const uint32_t* h_size = (const uint32_t*)(A + o_size);
    const uint64_t* runs = (const uint64_t*)(A + o_runs);
    if (n_res) {
        // the device builds an image of the arena (table columns + rows text), which then travels as ONE copy
        int rc = ensure(ctx, ctx->rowtext, total_bytes);
        uint8_t* img = (uint8_t*)ctx->rowtext.p;
        if (!rc && rows_on) rc = render_rows(ctx, res, n_res, (char*)(img + o_rows), (char*)(A + o_rows));
        if (rc) { delete res; return rc; }
        KbPackArgs pk{};
        pk.image = img;
        int ns = 0;
        auto seg = [&](const void* src, size_t off, size_t bytes) { if (bytes) { pk.src[ns] = (const uint32_t*)src; pk.dst_off[ns] = off; pk.words[ns] = bytes / 4; ns++; } };
        seg(ctx->res_flank.p, o_flank, n_res * lo.FW * 8);
        if (lo.MW) { seg(ctx->res_in.p, o_in, n_res * lo.MW * 4); seg(ctx->res_out.p, o_out, n_res * lo.MW * 4); }
        if (sizes_on) seg(ctx->res_size.p, o_size, n_res * 4);
        if (need_runs) seg(ctx->res_run.p, o_runs, n_res * 16);
        cudaError_t e = cudaSuccess;
        if (!lo.MW) e = cudaMemsetAsync(img + o_in, 0, o_size - o_in, ctx->stream);
        if (e == cudaSuccess) {
            const uint64_t most = n_res * std::max<uint64_t>(2 * (uint64_t)lo.FW, 4);
            kb_pack_kernel<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>((most + 255) / 256, (uint64_t)ctx->n_sm * 4)), 256, 0, ctx->stream>>>(pk);
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(A, img, total_bytes - 64, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { delete res; return fail(ctx, KB_ECUDA, std::string("result download: ") + cudaGetErrorString(e)); }
    }
    if (ctx->opt_want_records && n_res) {
        for (uint64_t g = 0; g < n_res; g++) res->run_offset[g + 1] = res->run_offset[g] + (hs ? (uint64_t)h_size[g] : runs[2 * g + 1]);
        const uint64_t total = res->run_offset[n_res];
        int rc = ensure(ctx, ctx->gather_off, n_res * 8);
        if (!rc) rc = ensure(ctx, ctx->gather_out, total * lo.W * 8);
        if (rc) { delete res; return rc; }
        res->records.resize(total * lo.W);
        cudaError_t e = cudaMemcpyAsync(ctx->gather_off.p, res->run_offset.data(), n_res * 8, cudaMemcpyHostToDevice, ctx->stream);
        KbGatherArgs g{};
        g.ent = sorted; g.recs = (const uint64_t*)ctx->recs.p; g.res_run = (const uint64_t*)ctx->res_run.p;
        g.off = (const uint64_t*)ctx->gather_off.p; g.n_groups = n_res; g.out = (uint64_t*)ctx->gather_out.p; g.lo = lo;
        const unsigned grid = (unsigned)((n_res + 7) / 8);
        if (e == cudaSuccess && hs) {
            KbHGatherArgs hg{};
            hg.ent = sorted; hg.recs = (const uint64_t*)ctx->recs.p; hg.res_run = (const uint64_t*)ctx->res_run.p;
            hg.res_flank = (const uint64_t*)ctx->res_flank.p; hg.off = (const uint64_t*)ctx->gather_off.p;
            hg.n_groups = n_res; hg.out = (uint64_t*)ctx->gather_out.p; hg.lo = lo;
            switch (lo.W) {
                case 1: kb_hgather_kernel<1><<<grid, 256, 0, ctx->stream>>>(hg); break;
                case 2: kb_hgather_kernel<2><<<grid, 256, 0, ctx->stream>>>(hg); break;
                case 4: kb_hgather_kernel<4><<<grid, 256, 0, ctx->stream>>>(hg); break;
                default: kb_hgather_kernel<8><<<grid, 256, 0, ctx->stream>>>(hg); break;
            }
            e = cudaGetLastError();
            ctx->launches++;
        } else if (e == cudaSuccess) {
            switch (lo.W) {
                case 1: kb_gather_kernel<1><<<grid, 256, 0, ctx->stream>>>(g); break;
                case 2: kb_gather_kernel<2><<<grid, 256, 0, ctx->stream>>>(g); break;
                case 4: kb_gather_kernel<4><<<grid, 256, 0, ctx->stream>>>(g); break;
                default: kb_gather_kernel<8><<<grid, 256, 0, ctx->stream>>>(g); break;
            }
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(res->records.data(), ctx->gather_out.p, total * lo.W * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { delete res; return fail(ctx, KB_ECUDA, std::string("record gather: ") + cudaGetErrorString(e)); }
        res->v.n_run_records = total;
    }
    res->v.flank = (const uint64_t*)(A + o_flank);
    res->v.in_mask = (const uint32_t*)(A + o_in);
    res->v.out_mask = (const uint32_t*)(A + o_out);
    res->v.group_size = sizes_on ? h_size : nullptr;
    res->v.run_offset = res->run_offset.data();
    res->v.records = res->records.data();
    *out = res;
    return KB_OK;
}

// Lazy records (kb_prefilter.cuh): K3a drops the occurrences whose hash-table entry misses a file, K3b builds the records of the rest.
// On return *kept = the compacted elements [hash32 | index] and hs carries their per-bucket runs.
static int run_prefilter(kb_ctx* ctx, const PartPlan& pl, uint64_t* parted, uint64_t n_bound, HashStage* hs, uint64_t** kept) {
    const KbLayout& lo = ctx->lo;
    DevBuf& other = parted == (uint64_t*)ctx->entA.p ? ctx->entB : ctx->entA;
    TRY(ensure(ctx, other, (size_t)(n_bound + 2048) * 8));
    TRY(ensure(ctx, ctx->brun, (size_t)hs->n_buckets * 16 + 16));
    unsigned long long* n_kept_dev = (unsigned long long*)ctx->small.p + SM_NTAINT;      // zero since prepare_small
    KbPrefilterArgs a{};
    a.ent = parted; a.bstart = hs->bstart; a.n_buckets = hs->n_buckets; a.bb = (uint32_t)pl.bb;
    a.tb = kb_prefilter_tb(lo.PW, a.bb);
    a.file_starts = (const uint64_t*)ctx->d_file_starts.p; a.file_gid = (const uint32_t*)ctx->d_file_gid.p;
    a.n_local_files = (int)ctx->file_gid.size();
    a.blk_shift = 0;
    while ((ctx->n_bases >> a.blk_shift) >= KB_PF_BLOCKS) a.blk_shift++;
    for (int f = 0; f < lo.n_files; f++) a.full[f >> 5] |= 1u << (f & 31);
    a.out = (uint64_t*)other.p; a.n_out = n_kept_dev; a.brun = (unsigned long long*)ctx->brun.p;
    const size_t smem = kb_prefilter_smem(lo.PW, a.tb);
    const unsigned pgrid = (unsigned)std::min<uint32_t>(hs->n_buckets, (uint32_t)ctx->n_sm * 6);
    prof_begin(ctx, "K3a hash prefilter");
    switch (kb_prefilter_pwn(lo.PW)) {
        case 2: CU(cudaFuncSetAttribute(kb_prefilter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kb_prefilter_kernel<2><<<pgrid, KB_PF_THREADS, smem, ctx->stream>>>(a); break;
        case 4: CU(cudaFuncSetAttribute(kb_prefilter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kb_prefilter_kernel<4><<<pgrid, KB_PF_THREADS, smem, ctx->stream>>>(a); break;
        default: CU(cudaFuncSetAttribute(kb_prefilter_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kb_prefilter_kernel<8><<<pgrid, KB_PF_THREADS, smem, ctx->stream>>>(a); break;
    }
    CU(cudaGetLastError());
    prof_end(ctx);
    ctx->launches++;
    CU(cudaMemcpyAsync(ctx->h_pinned + 9, n_kept_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const uint64_t n_kept = ctx->h_pinned[9];
    if (n_kept >= (1ULL << 32)) return fail(ctx, KB_EUNSUPPORTED, "more than 2^32 candidate records per GPU in multi-word mode");
    TRY(ensure(ctx, ctx->recs, std::max<uint64_t>(n_kept, 1) * 8 * lo.W));
    if (n_kept) {
        KbMatArgs m{};
        m.ent = (uint64_t*)other.p; m.n = n_kept; m.bases = (const uint8_t*)ctx->bases.p;
        m.file_starts = a.file_starts; m.file_gid = a.file_gid; m.n_local_files = a.n_local_files;
        m.lo = lo; m.recs = (uint64_t*)ctx->recs.p;
        const unsigned grid = (unsigned)std::min<uint64_t>((n_kept + 255) / 256, (uint64_t)ctx->n_sm * 16);
        prof_begin(ctx, "K3b build records");
        switch (lo.W) {
            case 2: kb_materialize_kernel<2><<<grid, 256, 0, ctx->stream>>>(m); break;
            case 4: kb_materialize_kernel<4><<<grid, 256, 0, ctx->stream>>>(m); break;
            default: kb_materialize_kernel<8><<<grid, 256, 0, ctx->stream>>>(m); break;
        }
        CU(cudaGetLastError());
        prof_end(ctx);
        ctx->launches++;
    }
    hs->brun = (const unsigned long long*)ctx->brun.p;
    hs->n_kept = n_kept;
    if (!ctx->opt_hash_slots_log2) {
        // exact pass: every kept table entry holds all files, so a bucket keeps about (kept records / files) keys — usually a
        // handful; a table sized for them (not for the whole bucket) is cleared in no time, and an overflow only splits the bucket
        const uint64_t keys = n_kept / std::max<uint64_t>(1, (uint64_t)hs->n_buckets * (uint64_t)std::max(lo.n_files, 1)) + 1;
        uint32_t l2 = 6;
        while (l2 < hs->pl->slots_log2 && ((uint64_t)1 << l2) < 4 * keys) l2++;
        hs->slots_log2 = l2;
    }
    *kept = (uint64_t*)other.p;
    return KB_OK;
}

// ---- slab search path (one-word records): K1 fused with partition level 0 (kb_extract_part.cuh), further levels and the bucket
//      hash on fixed-capacity slabs.  No histogram, no exact offsets: a level's cursors ARE its bucket table. ------------------------

// capacity of one of `nc` slabs that share n_est records: mean + 6 sigma (a key's occurrences move together) + a little
static uint64_t slab_capacity(const kb_ctx* ctx, uint64_t n_est, uint64_t nc, bool is_parent) {
    if (ctx->opt_slab_cap) return ((uint64_t)ctx->opt_slab_cap + 1) & ~1ULL;
    const double mu = (double)n_est / (double)nc;
    const double m = 2.0 * std::max(ctx->lo.n_files, 4);
    const uint64_t cap = (uint64_t)(mu + 6.0 * std::sqrt(m * mu) + 0.02 * mu + 64.0);
    // a level whose slabs are the parents of another level: a multiple of the partition tile, so that tiles map to parents by
    // division (no tile map, no dependent loads at CTA start) — where that costs at most a few percent of memory
    if (is_parent && cap >= 16 * (uint64_t)KB_PT_TILE) return (cap + KB_PT_TILE - 1) / KB_PT_TILE * KB_PT_TILE;
    return (cap + 1) & ~1ULL;                                   // even: every slab starts 16-byte aligned
}

// strand-symmetric level 0 (window items): the core must carry enough bits for the level-0 digit, and level 1 must exist to expand
static bool sym_ok(const kb_ctx* ctx, const PartPlan& pl, uint64_t* core_mask, int n_shards = 1) {
    const KbLayout& lo = ctx->lo;
    int core = 0;
    const uint64_t m = kb_core_mask(lo.L, lo.D, lo.R, &core);
    if (core_mask) *core_mask = m;
    const bool want = ctx->opt_sym < 0 ? n_shards >= 4 : ctx->opt_sym != 0;
    return want && lo.direct && lo.mix && pl.levels >= 2 && 2 * core >= pl.bits[0] + 10;
}

static SlabPlan make_slab_plan(const kb_ctx* ctx, const PartPlan& pl, uint64_t n_est) {
    SlabPlan sp;
    sp.levels = pl.levels;
    sp.sym = sym_ok(ctx, pl, &sp.core_mask);
    uint32_t maxnc = 0;
    size_t off = 0;
    for (int l = 0; l < pl.levels; l++) {
        sp.bits[l] = pl.bits[l];
        sp.nc[l] = pl.nc[l];
        sp.cap[l] = slab_capacity(ctx, (l == 0 && sp.sym) ? n_est / 2 : n_est, pl.nc[l], l + 1 < pl.levels);   // (symmetric level 0: one item per window)
        maxnc = std::max(maxnc, sp.nc[l]);
        sp.off_cur[l] = off; off += (size_t)sp.nc[l] * 8;
    }
    sp.off_counts = off; off += (size_t)maxnc * 8;
    sp.off_start = off; off += ((size_t)maxnc + 1) * 8;
    sp.off_part = off; off += ((size_t)maxnc / KB_PLAN_BLOCK + 2) * 16;
    sp.off_tab0 = off; off += (size_t)2 * KB_XP_MAXR * 8;       // level-0 slab ends | destination element addresses
    for (int l = 0; l < pl.levels; l++) { sp.off_tile0[l] = off; off += (((size_t)sp.nc[l] + 2) * 4 + 7) & ~(size_t)7; }
    sp.max_tiles = n_est / KB_PT_TILE + maxnc + 2;
    sp.off_tilemap = off; off += ((size_t)sp.max_tiles * 4 + 7) & ~(size_t)7;
    sp.off_snap = off; off += (size_t)(KB_MAX_BATCHES + 2) * sp.nc[0] * 8;      // level-0 cursors after every batch of input files
    sp.bytes = off;
    return sp;
}

static bool slab_ok(const kb_ctx* ctx, const PartPlan& pl) {
    const KbLayout& lo = ctx->lo;
    return ctx->opt_slab && !ctx->slab_off && ctx->opt_group_algo && lo.direct && pl.stream && pl.levels >= 1 && pl.bits[0] >= 1 && ctx->own_count < 0;
}

// K1 + level 0 over tiles [tile0, tile0 + n_tiles): cursor / limit / destination tables as in KbXPartArgs (device pointers).
// after_batch(i, n_batches, tiles of the batch) runs after the launch of every batch of arrived files (host buffers in flight).
template <class F>
static int run_extract_part(kb_ctx* ctx, uint32_t tile0, uint32_t n_tiles, uint64_t pos_lo, uint64_t pos_hi, uint32_t bits,
                            unsigned long long* cursor, const unsigned long long* limit, const unsigned long long* out_elems,
                            const std::vector<KBatch>& batches, F after_batch, bool sym = false, uint64_t core_mask = 0) {
    const KbLayout& lo = ctx->lo;
    KbXPartArgs a{};
    a.bases = (const uint8_t*)ctx->bases.p; a.n_bases = ctx->n_bases;
    a.file_starts = (const uint64_t*)ctx->d_file_starts.p; a.file_gid = (const uint32_t*)ctx->d_file_gid.p;
    a.n_local_files = (int)ctx->file_gid.size();
    a.soft_omit = ctx->soft_mode;
    a.lo = lo;
    a.pos_lo = pos_lo; a.pos_hi = pos_hi;
    a.bits = bits;
    a.cursor = cursor; a.limit = limit; a.out_elems = out_elems;
    a.n_out = (unsigned long long*)ctx->small.p + SM_NOUT;
    a.ovf = (unsigned long long*)ctx->small.p + SM_OVF;
    const bool spacer = lo.L == 25 && lo.D == 1 && lo.R == 2 && lo.mix;
    const size_t smem = sym ? kb_xsym_smem() : kb_xpart_smem();
    if (sym) CU(cudaFuncSetAttribute(kb_extract_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (spacer) CU(cudaFuncSetAttribute(kb_extract_part_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(cudaFuncSetAttribute(kb_extract_part_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t t0 = tile0;
    int bi = 0;
    for (auto& bt : batches) {
        if (bt.second) CU(cudaStreamWaitEvent(ctx->stream, bt.second, 0));
        TRY(deline_upto_file(ctx, bt.last_file));
        const uint32_t nt = bt.first > t0 ? bt.first - t0 : 0;
        if (nt) {
            a.tile0 = t0; a.n_tiles = nt;
            if (sym) {
                KbXSymArgs as{};
                as.x = a; as.core_mask = core_mask;
                kb_extract_items_kernel<<<(nt + KB_XS_TPC - 1) / KB_XS_TPC, KB_XP_THREADS, smem, ctx->stream>>>(as);
            } else if (spacer) kb_extract_part_kernel<true><<<nt, KB_XP_THREADS, smem, ctx->stream>>>(a);
            else kb_extract_part_kernel<false><<<nt, KB_XP_THREADS, smem, ctx->stream>>>(a);
            CU(cudaGetLastError());
            ctx->launches++;
            t0 = bt.first;
        }
        TRY(after_batch(bi, (int)batches.size(), nt));
        bi++;
    }
    ctx->alg_bytes += std::min<uint64_t>(pos_hi, ctx->n_bases) - pos_lo;
    ctx->alg_rec_bytes += sym ? 4 : 8;                                        // (one 8-byte item per window = per two records)
    return KB_OK;
}

// separator padding past the data + file table: what both K1 variants need before their first launch
static int prepare_extract(kb_ctx* ctx) {
    const size_t plen = padded_len(ctx->n_bases);
    TRY(ensure(ctx, ctx->bases, plen, true));
    CU(cudaMemsetAsync((uint8_t*)ctx->bases.p + ctx->n_bases, '\n', plen - ctx->n_bases, ctx->stream));
    return upload_file_table(ctx);
}

// One slab level l >= 1: reads the slabs of level l - 1 in `in` (fill levels pend, optionally only the segments from pbegin on),
// writes the slabs of level l into `out`.  rec_bound = records the pass can meet at most (sizes the grid in tile-map mode).
static int launch_slab_level(kb_ctx* ctx, const SlabPlan& sp, const PartPlan& pl, int l, const uint64_t* in, uint64_t* out,
                             const unsigned long long* pend, const unsigned long long* pbegin, uint32_t n_parents, const uint32_t* prow,
                             uint64_t rec_bound, uint32_t psel_n = 0, uint32_t psel_j0 = 0, uint32_t psel_dps = 0, uint32_t psel_skip = 0xFFFFFFFFu) {
    uint8_t* P = (uint8_t*)ctx->plan.p;
    const size_t smem = kb_part_smem();
    const bool expand = sp.sym && l == 1;                      // the parents hold window items: form the records here
    const KbLayout& lo = ctx->lo;
    const bool spacer = lo.L == 25 && lo.D == 1 && lo.R == 2 && lo.mix;
    if (expand && spacer) CU(cudaFuncSetAttribute(kb_part_expand_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (expand) CU(cudaFuncSetAttribute(kb_part_expand_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(cudaFuncSetAttribute(kb_part_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int shift = 64;
    for (int i = sp.sym ? 1 : 0; i <= l; i++) shift -= sp.bits[i];   // (symmetric level 0 eats no bits of the mixed key)
    KbPartArgs a{};
    a.in = in; a.out = out;
    a.pend = pend; a.pbegin = pbegin; a.pcap = sp.cap[l - 1];
    a.ptile0 = (const uint32_t*)(P + sp.off_tile0[l - 1]);
    a.tile_parent = (const uint32_t*)(P + sp.off_tilemap);
    a.prow = prow;
    a.n_parents = n_parents;
    a.psel_n = psel_n; a.psel_j0 = psel_j0; a.psel_dps = psel_dps; a.psel_skip = psel_skip;
    a.shift = (uint32_t)shift; a.bits = (uint32_t)sp.bits[l];
    a.cursor = (unsigned long long*)(P + sp.off_cur[l]);
    a.ccap = sp.cap[l];
    a.ovf = (unsigned long long*)ctx->small.p + SM_OVF;
    const bool by_division = !pbegin && sp.cap[l - 1] % KB_PT_TILE == 0;
    uint64_t grid = (uint64_t)n_parents * (sp.cap[l - 1] / KB_PT_TILE);
    if (by_division) { a.ptile0 = nullptr; a.tile_parent = nullptr; }
    else if (n_parents <= KB_PLAN_BLOCK) {
        // few parents (the level-0 slabs of a batch of files): counts, tile prefix and tile map in one single-CTA launch
        kb_slab_plan_fused_kernel<<<1, KB_PLAN_BLOCK, 0, ctx->stream>>>(pend, pbegin, n_parents, a.pcap, (unsigned long long*)(P + sp.off_start),
                                                                        (uint32_t*)(P + sp.off_tile0[l - 1]), (uint32_t*)(P + sp.off_tilemap));
        CU(cudaGetLastError());
        ctx->launches++;
        grid = rec_bound / KB_PT_TILE + n_parents + 1;
    } else {
        unsigned long long* counts = (unsigned long long*)(P + sp.off_counts);
        kb_slab_counts_kernel<<<(unsigned)std::min<uint32_t>((n_parents + 255) / 256, 1024), 256, 0, ctx->stream>>>(pend, pbegin, n_parents, a.pcap, counts);
        CU(cudaGetLastError());
        KbPlanArgs pa{};
        pa.counts = counts; pa.nc = n_parents; pa.start = (unsigned long long*)(P + sp.off_start); pa.cursor = nullptr;
        pa.tile0 = (uint32_t*)(P + sp.off_tile0[l - 1]); pa.part = (unsigned long long*)(P + sp.off_part);
        TRY(launch_plan(ctx, pa, pl));
        kb_tilemap_kernel<<<(unsigned)std::min<uint32_t>((n_parents + 7) / 8, 4096), 256, 0, ctx->stream>>>(a.ptile0, n_parents, (uint32_t*)(P + sp.off_tilemap));
        CU(cudaGetLastError());
        ctx->launches += 2;
        grid = rec_bound / KB_PT_TILE + n_parents + 1;
    }
    if (expand && spacer) kb_part_expand_kernel<true><<<(unsigned)grid, KB_PT_THREADS, smem, ctx->stream>>>(a, lo);
    else if (expand) kb_part_expand_kernel<false><<<(unsigned)grid, KB_PT_THREADS, smem, ctx->stream>>>(a, lo);
    else kb_part_kernel<2, false, true><<<(unsigned)grid, KB_PT_THREADS, smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
    return KB_OK;
}

// partition levels [l_begin, levels) on slabs; level l reads buf[(l - 1) & 1] and writes buf[l & 1]
static int run_slab_levels(kb_ctx* ctx, const SlabPlan& sp, const PartPlan& pl, int l_begin, uint64_t* bufs[2], uint64_t n_est) {
    uint8_t* P = (uint8_t*)ctx->plan.p;
    static const char* pnames[3] = {"K2 partition 0", "K2 partition 1", "K2 partition 2"};
    for (int l = l_begin; l < sp.levels; l++) {
        prof_begin(ctx, pnames[l]);
        TRY(launch_slab_level(ctx, sp, pl, l, bufs[(l - 1) & 1], bufs[l & 1], (const unsigned long long*)(P + sp.off_cur[l - 1]), nullptr,
                              sp.nc[l - 1], nullptr, n_est));
        prof_end(ctx);
        ctx->alg_rec_bytes += (sp.sym && l == 1) ? 12 : 16;
    }
    return KB_OK;
}

static int search_slab(kb_ctx* ctx, const PartPlan& pl, kb_result** out) {
    const uint64_t n_est = 2 * ctx->n_bases + 64;
    const SlabPlan sp = make_slab_plan(ctx, pl, n_est);
    size_t need[2] = {0, 0};
    for (int l = 0; l < sp.levels; l++) need[l & 1] = std::max(need[l & 1], (size_t)sp.nc[l] * sp.cap[l] + 4096);
    TRY(ensure(ctx, ctx->entA, need[0] * 8));
    if (need[1]) TRY(ensure(ctx, ctx->entB, need[1] * 8));
    TRY(ensure(ctx, ctx->plan, sp.bytes + 64));
    uint8_t* P = (uint8_t*)ctx->plan.p;
    uint64_t* bufs[2] = {(uint64_t*)ctx->entA.p, (uint64_t*)ctx->entB.p};
    for (int l = 0; l < sp.levels; l++) {
        kb_slab_init_kernel<<<(unsigned)std::min<uint32_t>((sp.nc[l] + 255) / 256, 1024), 256, 0, ctx->stream>>>((unsigned long long*)(P + sp.off_cur[l]), sp.nc[l], sp.cap[l]);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    // level-0 tables: slab ends, destination (this GPU's buffer for every digit)
    std::vector<uint64_t>& t0 = ctx->scatter_host;
    t0.assign((size_t)2 * KB_XP_MAXR, 0);
    for (uint32_t d = 0; d < sp.nc[0]; d++) { t0[d] = (uint64_t)(d + 1) * sp.cap[0]; t0[KB_XP_MAXR + d] = (uint64_t)(reinterpret_cast<uintptr_t>(bufs[0]) >> 3); }
    CU(cudaMemcpyAsync(P + sp.off_tab0, t0.data(), t0.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t n_tiles = (uint32_t)((ctx->n_bases + KB_K1_TB - 1) / KB_K1_TB);
    TRY(prepare_extract(ctx));
    const auto batches = extract_batches(ctx, 0, n_tiles);
    // Host buffers still arriving: level 1 runs per batch of files too, on the segment the batch appended to every level-0 slab
    // (cursor snapshots), so that after the last copy only the last batch's share of K1 / level 0 / level 1 is left
    const bool per_batch = ctx->opt_batch_level0 && sp.levels >= 2 && batches.size() > 1 && batches.size() <= KB_MAX_BATCHES;
    unsigned long long* cur0 = (unsigned long long*)(P + sp.off_cur[0]);
    unsigned long long* snap = (unsigned long long*)(P + sp.off_snap);
    if (per_batch) CU(cudaMemcpyAsync(snap, cur0, (size_t)sp.nc[0] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    prof_begin(ctx, per_batch ? "K1 extract + partition 0 + partition 1 per batch" : "K1 extract + partition 0");
    TRY(run_extract_part(ctx, 0, n_tiles, 0, ctx->n_bases, (uint32_t)sp.bits[0], cur0,
                         (const unsigned long long*)(P + sp.off_tab0), (const unsigned long long*)(P + sp.off_tab0) + KB_XP_MAXR, batches,
                         [&](int bi, int, uint32_t nt) -> int {
                             if (!per_batch) return KB_OK;
                             unsigned long long* s0 = snap + (size_t)bi * sp.nc[0];
                             unsigned long long* s1 = s0 + sp.nc[0];
                             CU(cudaMemcpyAsync(s1, cur0, (size_t)sp.nc[0] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                             if (!nt) return KB_OK;
                             return launch_slab_level(ctx, sp, pl, 1, bufs[0], bufs[1], s1, s0, sp.nc[0], nullptr, 2ull * nt * KB_K1_TB);
                         }, sp.sym, sp.core_mask));
    prof_end(ctx);
    if (per_batch) ctx->alg_rec_bytes += sp.sym ? 12 : 16;
    TRY(run_slab_levels(ctx, sp, pl, per_batch ? 2 : 1, bufs, n_est));
    ctx->passes = sp.levels;
    HashStage hs{};
    hs.pl = &pl;
    hs.n_buckets = sp.nc[sp.levels - 1];
    hs.bend = (const unsigned long long*)(P + sp.off_cur[sp.levels - 1]);
    hs.bcap = sp.cap[sp.levels - 1];
    if (sp.sym) hs.bb_hash = pl.bb - sp.bits[0];
    int rc = run_group(ctx, bufs[(sp.levels - 1) & 1], n_est, out, &hs);
    prof_collect(ctx);
    return rc;
}

// every added file id must exist in the current configuration: an id >= n_files would set a presence bit outside the "all files"
// mask and silently empty the result
static int check_file_ids(kb_ctx* ctx) {
    for (uint32_t g : ctx->file_gid)
        if ((int)g >= ctx->lo.n_files)
            return fail(ctx, KB_EINVAL, "a sequence was added with file id " + std::to_string(g) + " but kb_configure was told of " + std::to_string(ctx->lo.n_files) + " files");
    return KB_OK;
}

static void begin_search(kb_ctx* ctx) {
    ctx->lazy_now = false;
    ctx->launches = 0; ctx->alg_bytes = 0; ctx->alg_rec_bytes = 0; ctx->passes = 0;
    for (auto& e : ctx->prof_events) { cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second); }
    ctx->prof_events.clear();
}

static int search_once(kb_ctx* ctx, kb_result** out);

extern "C" {

int kb_search(kb_ctx* ctx, kb_result** out) {
    if (!ctx || !out) return KB_EINVAL;
    *out = nullptr;
    if (!ctx->configured) return fail(ctx, KB_EINVAL, "kb_configure has not been called");
    CU(cudaSetDevice(ctx->device));
    // The plan sizes the buckets from the record count, assuming that a key occurs in a good part of the files.  Divergent genomes
    // have far more distinct keys: the stream kernel then defers most buckets, and splitting each of them in the fallback kernel
    // costs more than a finer partition.  So a search that defers more than 1/8 of its buckets is abandoned and repeated with
    // two more bucket bits (the sequences are still resident); the context remembers the bits for the next search.
    for (int attempt = 0; ; attempt++) {
        ctx->replan_ok = attempt < 3 && ctx->opt_bucket_bits < 0 && ctx->opt_group_algo && ctx->lo.direct;
        const int rc = search_once(ctx, out);
        ctx->replan_ok = false;
        if (rc == KB_SLABOVF) { ctx->slab_off = true; attempt--; continue; }     // a slab overflowed: the exact path from now on
        if (rc != KB_REPLAN) return rc;
        ctx->bb_extra += 2;
    }
}

}  // extern "C"

static int search_once(kb_ctx* ctx, kb_result** out) {
    TRY(check_file_ids(ctx));
    begin_search(ctx);
    TRY(prepare_small(ctx));
    const KbLayout& lo = ctx->lo;
    uint64_t n = 0;
    const uint32_t n_tiles = (uint32_t)((ctx->n_bases + KB_K1_TB - 1) / KB_K1_TB);
    ctx->lazy_now = ctx->opt_group_algo && ctx->opt_lazy_records && !lo.direct && padded_len(ctx->n_bases) < (1ULL << 32);
    if (ctx->opt_group_algo) {
        const PartPlan pl = make_plan(ctx, 2 * ctx->n_bases + 64);
        if (slab_ok(ctx, pl)) return search_slab(ctx, pl, out);
        TRY(ensure(ctx, ctx->plan, pl.bytes + 64));
        if (pl.levels) CU(cudaMemsetAsync(ctx->plan.p, 0, pl.bytes, ctx->stream));
        // K1 counts the children of the first level — or of the first TWO levels at once (packed shared-memory histogram), which
        // saves the pass that would re-read every record just to count level 1
        const bool fused2 = ctx->opt_fused_hist && pl.levels >= 2 && pl.bits[0] + pl.bits[1] >= 10 && pl.bits[0] + pl.bits[1] <= 16;
        // 17 bits in two levels: K1 still counts 16 — the sizes of sibling pairs of level-1 children, which is all the pass needs when
        // the two siblings fill their common range from both ends (KbPartArgs::pair_mode)
        const bool pair = ctx->opt_fused_hist && ctx->opt_pair_hist && pl.levels == 2 && pl.bits[0] + pl.bits[1] == 17;
        const int hl = fused2 ? 1 : 0;
        const uint32_t hbits = pair ? 16u : (fused2 ? (uint32_t)(pl.bits[0] + pl.bits[1]) : (uint32_t)pl.bits[0]);
        unsigned long long* h0 = pl.levels ? (unsigned long long*)((uint8_t*)ctx->plan.p + (pair ? pl.off_pair : pl.off_cnt[hl])) : nullptr;
        BatchL0 bl;
        if (fused2 && ctx->opt_batch_level0 && pl.levels == 2 && lo.direct) {
            // carve the batch tables: [B][nc1] two-level counts | [B*nc0] level-0 counts | [B*nc0+1] offsets (x2) | [B][nc0] cursors | roots | tiles | rows
            bl.nc0 = pl.nc[0]; bl.nc1 = pl.nc[1]; bl.bits0 = (uint32_t)pl.bits[0]; bl.bits1 = (uint32_t)pl.bits[1];
            const size_t B = KB_MAX_BATCHES, np = B * bl.nc0;
            const size_t words = B * bl.nc1 + np + 2 * (np + 1) + np + 2 * B + B + (np + 2) / 2 + (np + 1) / 2 + 8;
            TRY(ensure(ctx, ctx->batchbuf, words * 8));
            TRY(ensure(ctx, ctx->entA, (size_t)(2 * ctx->n_bases + 64 + 2048) * 8));      // (what run_extract will ask for)
            TRY(ensure(ctx, ctx->entB, (size_t)(2 * ctx->n_bases + 64 + 2048) * 8));
            CU(cudaMemsetAsync(ctx->batchbuf.p, 0, words * 8, ctx->stream));
            unsigned long long* q = (unsigned long long*)ctx->batchbuf.p;
            bl.hist2 = q; q += B * bl.nc1;
            bl.counts_all = q; q += np;
            bl.start_all = q; q += np + 1;
            bl.scratch_start = q; q += np + 1;
            bl.cursors = q; q += np;
            bl.roots = q; q += 2 * B;
            bl.roottiles = (uint32_t*)q; q += B;
            bl.ptile0_all = (uint32_t*)q; q += (np + 2) / 2;
            bl.prow = (uint32_t*)q;
            bl.total2 = (unsigned long long*)((uint8_t*)ctx->plan.p + pl.off_cnt[1]);
            bl.out = (uint64_t*)ctx->entB.p;
            bl.enabled = true;
        }
        TRY(run_extract(ctx, lo, 0, n_tiles, 0, ctx->n_bases, &n, h0, 64 - hbits, hbits, true, bl.enabled ? &bl : nullptr));
        if (bl.n_batches > 0) {
            // level 0 is done batch by batch: the (batch, digit) pieces are the parents of level 1 (rows = digits)
            uint8_t* P = (uint8_t*)ctx->plan.p;
            const uint32_t np = (uint32_t)bl.n_batches * bl.nc0;
            KbPlanArgs pp{};
            pp.counts = bl.counts_all; pp.nc = np; pp.start = bl.scratch_start; pp.cursor = nullptr; pp.tile0 = bl.ptile0_all;
            TRY(launch_plan(ctx, pp, pl));
            KbPlanArgs p1{};
            p1.counts = bl.total2; p1.nc = pl.ncl[1];
            p1.start = (unsigned long long*)(P + pl.off_start[1]); p1.cursor = bl.total2; p1.tile0 = (uint32_t*)(P + pl.off_tile0[1]);
            TRY(launch_plan(ctx, p1, pl));
            std::vector<uint32_t>& rows = ctx->batch_rows_host;
            rows.resize(np);
            for (uint32_t i = 0; i < np; i++) rows[i] = i % bl.nc0;
            CU(cudaMemcpyAsync(bl.prow, rows.data(), (size_t)np * 4, cudaMemcpyHostToDevice, ctx->stream));
            CustomParents cp{};
            cp.pstart = bl.start_all; cp.ptile0 = bl.ptile0_all; cp.prow = bl.prow; cp.n = np;
            cp.max_tiles = n / KB_PT_TILE + np + 2;
            ctx->alg_rec_bytes += 16;
            ctx->passes += 1;
            uint64_t* parted = nullptr;
            HashStage hs{};
            hs.pl = &pl;
            TRY(run_partition(ctx, pl, ctx->entB, ctx->entA, n, 1, pl.levels, &cp, &parted, &hs.bstart, &hs.n_buckets, true));
            int rc = run_group(ctx, parted, n, out, &hs);
            prof_collect(ctx);
            return rc;
        }
        if (!lo.direct && !ctx->lazy_now && n >= (1ULL << 32) + 64) return fail(ctx, KB_EUNSUPPORTED, "more than 2^32 records per GPU in multi-word mode");
        uint64_t* parted = nullptr;
        HashStage hs{};
        hs.pl = &pl;
        TRY(run_partition(ctx, pl, ctx->entA, ctx->entB, n, 0, pl.levels, nullptr, &parted, &hs.bstart, &hs.n_buckets, fused2 || pair, false, pair));
        if (ctx->lazy_now && n > 0) TRY(run_prefilter(ctx, pl, parted, n, &hs, &parted));
        int rc = run_group(ctx, parted, n, out, &hs);
        prof_collect(ctx);
        return rc;
    }
    TRY(run_extract(ctx, lo, 0, n_tiles, 0, ctx->n_bases, &n));
    if (!lo.direct && n >= (1ULL << 32)) return fail(ctx, KB_EUNSUPPORTED, "more than 2^32 records per GPU in multi-word mode");
    uint64_t* sorted = nullptr;
    TRY(run_sort(ctx, ctx->entA, ctx->entB, n, lo.P, &sorted));
    int rc = run_group(ctx, sorted, n, out);
    prof_collect(ctx);
    return rc;
}

extern "C" {

// ---- multi-GPU: level 0 of the partition decides the owner shard (contiguous digit ranges), the exchange moves every
//      level-0 bucket to its owner, levels >= 1 and the bucket hash run there ------------------------------------------
static uint32_t shard_first_digit(uint32_t shard, uint32_t n_shards, uint32_t n_digits) {
    return (uint32_t)(((uint64_t)shard * n_digits) / n_shards);
}

int kb_shard_plan(kb_ctx* ctx, int n_shards, int shard_index, uint64_t total_bases, int* n_digits) {
    if (!ctx || n_shards < 1 || n_shards > 256 || shard_index < 0 || shard_index >= n_shards) return KB_EINVAL;
    if (!ctx->configured) return fail(ctx, KB_EINVAL, "kb_configure has not been called");
    const KbLayout& lo = ctx->lo;
    if (!lo.direct && !(ctx->opt_lazy_records && ctx->opt_group_algo))
        return fail(ctx, KB_EUNSUPPORTED, "multi-GPU sharding of multi-word records needs lazy_records = 1 and group_algo = 1");
    int min_bits0 = 0;
    while ((1 << min_bits0) < n_shards) min_bits0++;
    if (lo.FB < min_bits0 + 1) return fail(ctx, KB_EUNSUPPORTED, "flank key too short to shard over this many GPUs");
    // every rank derives the same plan from the same numbers (records over all ranks ~ 2 * total bases).  Level 0 is the
    // exchange: 4 digits per shard keep the digit runs of a tile long (efficient peer stores) and the piece table small.
    // Measured at N = 2 (0.2 Gbp per GPU): 8 + 9 bits in two levels (256-byte runs, 450 GB/s of peer stores) beats 3 + 7 + 7 in
    // three (8 KB runs, 570 GB/s, but one more pass): stay with two levels while the bucket bits allow it.
    int bits0 = (int)ctx->opt_shard_bits0;
    if (bits0 <= 0) {
        const PartPlan probe = make_plan(ctx, 2 * total_bases + 64, 0, false);
        bits0 = (probe.bb >= min_bits0 + 9 && probe.bb <= 17) ? probe.bb - 9 : min_bits0 + 2;
    }
    bits0 = std::max(std::max(min_bits0, 1), std::min(std::min(bits0, 9), lo.FB - 1));
    PartPlan pl = make_plan(ctx, 2 * total_bases + 64, bits0, false);
    if (pl.levels < 2 || pl.levels > 3 || pl.bits[0] < min_bits0 || pl.bb > lo.FB) return fail(ctx, KB_EINTERNAL, "shard plan");
    ctx->shard_plan = pl;
    ctx->shard_n = n_shards;
    ctx->shard_index = shard_index;
    if (n_digits) *n_digits = (int)pl.nc[0];
    return KB_OK;
}

int kb_sequence_buffer(kb_ctx* ctx, void** device_bytes, uint64_t* n_bytes) {
    if (!ctx || !device_bytes || !n_bytes) return KB_EINVAL;
    CU(cudaSetDevice(ctx->device));
    TRY(deline_pending(ctx, ctx->pending_fasta.size()));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *device_bytes = ctx->bases.p;
    *n_bytes = ctx->n_bases;
    return KB_OK;
}

int kb_shard_own_files(kb_ctx* ctx, int first_local, int n_local) {
    if (!ctx) return KB_EINVAL;
    const int nf = (int)ctx->file_starts.size();
    if (n_local >= 0 && (first_local < 0 || first_local + n_local > nf)) return fail(ctx, KB_EINVAL, "own file range outside the added sequences");
    ctx->own_first = n_local < 0 ? 0 : first_local;
    ctx->own_count = n_local;
    return KB_OK;
}

// positions / K1 tiles of this rank's own files, and the lazy-records switch of the shard calls
static int shard_extract_range(kb_ctx* ctx, uint64_t* pos_lo, uint64_t* pos_hi, uint32_t* tile0, uint32_t* n_tiles) {
    const int nf = (int)ctx->file_starts.size();
    *pos_lo = 0; *pos_hi = ctx->n_bases;
    if (ctx->own_count >= 0) {
        *pos_lo = ctx->own_first < nf ? ctx->file_starts[ctx->own_first] : ctx->n_bases;
        *pos_hi = ctx->own_first + ctx->own_count < nf ? ctx->file_starts[ctx->own_first + ctx->own_count] : ctx->n_bases;
    }
    *tile0 = (uint32_t)(*pos_lo / KB_K1_TB);
    *n_tiles = (uint32_t)((*pos_hi + KB_K1_TB - 1) / KB_K1_TB) - *tile0;
    ctx->lazy_now = !ctx->lo.direct;
    if (ctx->lazy_now && padded_len(ctx->n_bases) >= (1ULL << 32))
        return fail(ctx, KB_EUNSUPPORTED, "multi-word records on several GPUs: more than 2^32 sequence bytes over all ranks");
    return KB_OK;
}

int kb_shard_extract(kb_ctx* ctx, void** records, uint64_t* shard_counts, uint64_t* digit_counts) {
    if (!ctx || !records || !shard_counts || !digit_counts) return KB_EINVAL;
    if (!ctx->configured || ctx->shard_n < 1) return fail(ctx, KB_EINVAL, "kb_configure / kb_shard_plan have not been called");
    const KbLayout& lo = ctx->lo;
    const PartPlan& pl = ctx->shard_plan;
    CU(cudaSetDevice(ctx->device));
    TRY(check_file_ids(ctx));
    begin_search(ctx);
    TRY(prepare_small(ctx));
    // the plan buffer also has to hold this rank's tile map: its own records may outnumber the per-shard estimate
    const size_t tilemap_extra = (size_t)((2 * ctx->n_bases + 64) / KB_PT_TILE + 2) * 4;
    TRY(ensure(ctx, ctx->plan, pl.bytes + tilemap_extra + 64));
    CU(cudaMemsetAsync(ctx->plan.p, 0, pl.bytes, ctx->stream));
    uint64_t n = 0, pos_lo = 0, pos_hi = 0;
    uint32_t tile0 = 0, n_tiles = 0;
    TRY(shard_extract_range(ctx, &pos_lo, &pos_hi, &tile0, &n_tiles));
    unsigned long long* h0 = (unsigned long long*)((uint8_t*)ctx->plan.p + pl.off_cnt[0]);
    TRY(run_extract(ctx, lo, tile0, n_tiles, pos_lo, pos_hi, &n, h0, (uint32_t)(64 - pl.bits[0]), (uint32_t)pl.bits[0], true));
    uint64_t* parted = nullptr;
    const unsigned long long* start0 = nullptr;
    uint32_t nd = 0;
    TRY(run_partition(ctx, pl, ctx->entA, ctx->entB, n, 0, 1, nullptr, &parted, &start0, &nd));
    // level-0 bucket offsets -> per-digit and per-shard counts
    std::vector<uint64_t> st((size_t)nd + 1);
    CU(cudaMemcpyAsync(st.data(), start0, ((size_t)nd + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t d = 0; d < nd; d++) digit_counts[d] = st[d + 1] - st[d];
    for (int sh = 0; sh < ctx->shard_n; sh++) {
        const uint32_t d0 = shard_first_digit((uint32_t)sh, (uint32_t)ctx->shard_n, nd), d1 = shard_first_digit((uint32_t)sh + 1, (uint32_t)ctx->shard_n, nd);
        shard_counts[sh] = st[d1] - st[d0];
    }
    const uint64_t n_local = st[nd];
    ctx->alg_bytes += n_local * ctx->alg_rec_bytes;
    ctx->alg_rec_bytes = 0;
    *records = parted;
    ctx->shard_send_in_B = (parted == (uint64_t*)ctx->entB.p) ? 1 : 0;
    ctx->shard_direct = 0;
    ctx->shard_n_records = n_local;
    prof_collect(ctx);
    return KB_OK;
}

int kb_shard_ipc_export(kb_ctx* ctx, uint64_t capacity_records, uint8_t* handle64) {
    if (!ctx || !handle64) return KB_EINVAL;
    CU(cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    const size_t bytes = (size_t)(capacity_records + 2048) * 8;
    if (bytes > ctx->recvbuf.cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->recvbuf.p) CU(cudaFree(ctx->recvbuf.p));
        ctx->recvbuf.p = nullptr; ctx->recvbuf.cap = 0;
        CU(cudaMalloc(&ctx->recvbuf.p, bytes));            // a plain cudaMalloc allocation: exportable
        ctx->recvbuf.cap = bytes;
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->recvbuf.p));
    memcpy(handle64, &h, 64);
    return KB_OK;
}

// Unmap every peer's receive buffer.  Exported memory must not be freed while another process still maps it, so when a buffer has
// to grow ALL ranks close first (then a barrier), and only then the owners reallocate and export again.
int kb_shard_ipc_close(kb_ctx* ctx) {
    if (!ctx) return KB_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < ctx->peer_ptr.size(); r++) {
        if (!ctx->peer_ptr[r] || (int)r == ctx->shard_index) continue;
        if (cudaIpcCloseMemHandle(ctx->peer_ptr[r]) != cudaSuccess) cudaGetLastError();     // (do not leave a latent error behind)
    }
    ctx->peer_ptr.clear();
    return KB_OK;
}

int kb_shard_ipc_import(kb_ctx* ctx, int n_ranks, const uint8_t* handles) {
    if (!ctx || !handles || n_ranks < 1 || n_ranks != ctx->shard_n) return KB_EINVAL;
    TRY(kb_shard_ipc_close(ctx));
    ctx->peer_ptr.assign((size_t)n_ranks, nullptr);
    for (int r = 0; r < n_ranks; r++) {
        if (r == ctx->shard_index) { ctx->peer_ptr[r] = ctx->recvbuf.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_ptr[r] = p;
    }
    return KB_OK;
}

int kb_shard_count(kb_ctx* ctx, uint64_t* digit_counts) {
    if (!ctx || !digit_counts) return KB_EINVAL;
    if (!ctx->configured || ctx->shard_n < 1) return fail(ctx, KB_EINVAL, "kb_configure / kb_shard_plan have not been called");
    const KbLayout& lo = ctx->lo;
    const PartPlan& pl = ctx->shard_plan;
    CU(cudaSetDevice(ctx->device));
    TRY(check_file_ids(ctx));
    begin_search(ctx);
    TRY(prepare_small(ctx));
    const size_t tilemap_extra = (size_t)((2 * ctx->n_bases + 64) / KB_PT_TILE + 2) * 4;
    TRY(ensure(ctx, ctx->plan, pl.bytes + tilemap_extra + 64));
    CU(cudaMemsetAsync(ctx->plan.p, 0, pl.bytes, ctx->stream));
    uint64_t n = 0, pos_lo = 0, pos_hi = 0;
    uint32_t tile0 = 0, n_tiles = 0;
    TRY(shard_extract_range(ctx, &pos_lo, &pos_hi, &tile0, &n_tiles));
    // K1 counts the level-0 digits — or, when the first two levels have 10..16 bits together, the children of level 1 as well
    // (packed shared-memory histogram): the owners then get the level-1 counts with the digit counts and skip the histogram pass
    const int hb2 = pl.levels >= 2 ? pl.bits[0] + pl.bits[1] : 0;
    const bool two = ctx->opt_fused_hist && hb2 >= 10 && hb2 <= 16;
    const int hl = two ? 1 : 0;
    const uint32_t hbits = two ? (uint32_t)hb2 : (uint32_t)pl.bits[0];
    unsigned long long* h0 = (unsigned long long*)((uint8_t*)ctx->plan.p + pl.off_cnt[hl]);
    TRY(run_extract(ctx, lo, tile0, n_tiles, pos_lo, pos_hi, &n, h0, 64 - hbits, hbits, true));
    std::vector<uint64_t>& st = ctx->scatter_host;
    st.assign(pl.nc[hl], 0);
    CU(cudaMemcpyAsync(st.data(), h0, (size_t)pl.nc[hl] * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    uint64_t n_local = 0;
    const uint32_t fold = two ? (1u << pl.bits[1]) : 1u;
    for (uint32_t d = 0; d < pl.nc[0]; d++) {
        uint64_t c = 0;
        for (uint32_t j = 0; j < fold; j++) c += st[(size_t)d * fold + j];
        digit_counts[d] = c; n_local += c;
    }
    if (two) ctx->shard_child_host = st; else ctx->shard_child_host.clear();
    ctx->shard_child_set.clear();
    ctx->alg_bytes += n_local * ctx->alg_rec_bytes;
    ctx->alg_rec_bytes = 0;
    ctx->shard_n_records = n_local;
    prof_collect(ctx);
    return KB_OK;
}

int kb_shard_child_counts(kb_ctx* ctx, uint64_t* counts, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return KB_EINVAL;
    *n = ctx->shard_child_host.size();
    if (counts && cap >= *n && *n) memcpy(counts, ctx->shard_child_host.data(), *n * 8);
    return KB_OK;
}

int kb_shard_set_child_counts(kb_ctx* ctx, const uint64_t* counts, uint64_t n) {
    if (!ctx || (n && !counts)) return KB_EINVAL;
    ctx->shard_child_set.assign(counts, counts + n);
    return KB_OK;
}

int kb_shard_scatter(kb_ctx* ctx, const uint64_t* piece_base) {
    if (!ctx || !piece_base) return KB_EINVAL;
    if (!ctx->configured || ctx->shard_n < 1) return fail(ctx, KB_EINVAL, "kb_configure / kb_shard_plan have not been called");
    if ((int)ctx->peer_ptr.size() != ctx->shard_n) return fail(ctx, KB_EINVAL, "kb_shard_ipc_import has not been called");
    const PartPlan& pl = ctx->shard_plan;
    CU(cudaSetDevice(ctx->device));
    const uint32_t nd = pl.nc[0];
    const uint64_t n = ctx->shard_n_records;
    uint8_t* P = (uint8_t*)ctx->plan.p;
    // per digit: cursor = start of this rank's piece in the owner's buffer; destination = the owner's buffer
    std::vector<uint64_t>& hp = ctx->scatter_host;
    hp.assign((size_t)nd * 2, 0);
    for (int sh = 0; sh < ctx->shard_n; sh++) {
        const uint32_t d0 = shard_first_digit((uint32_t)sh, (uint32_t)ctx->shard_n, nd), d1 = shard_first_digit((uint32_t)sh + 1, (uint32_t)ctx->shard_n, nd);
        for (uint32_t d = d0; d < d1; d++) { hp[d] = piece_base[d]; hp[nd + d] = (uint64_t)(reinterpret_cast<uintptr_t>(ctx->peer_ptr[sh]) >> 3); }
    }
    TRY(ensure(ctx, ctx->shard_tab, ((size_t)nd * 2 + 8) * 8));
    unsigned long long* cursor = (unsigned long long*)(P + pl.off_cnt[0]);
    CU(cudaMemcpyAsync(cursor, hp.data(), (size_t)nd * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->shard_tab.p, hp.data() + nd, (size_t)nd * 8, cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long* root = (unsigned long long*)ctx->small.p + SM_ROOT;
    uint32_t* roottile = (uint32_t*)((uint64_t*)ctx->small.p + SM_ROOTTILE);
    kb_root_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long*)ctx->small.p + SM_NOUT, root, roottile);
    CU(cudaGetLastError());
    KbPartArgs a{};
    a.in = (const uint64_t*)ctx->entA.p; a.out = nullptr;
    a.pstart = root; a.ptile0 = roottile; a.tile_parent = nullptr; a.n_parents = 1;
    a.shift = (uint32_t)(64 - pl.bits[0]); a.bits = (uint32_t)pl.bits[0];
    a.cursor = cursor; a.hist = cursor;
    a.out_elems = (const unsigned long long*)ctx->shard_tab.p;
    const size_t smem = kb_part_smem();
    CU(cudaFuncSetAttribute(kb_part_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(ctx, "K2 partition 0 + exchange (peer stores)");
    if (n) kb_part_kernel<2><<<(unsigned)((n + KB_PT_TILE - 1) / KB_PT_TILE), KB_PT_THREADS, smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    prof_end(ctx);
    ctx->launches += 2;
    ctx->alg_bytes += n * 16;
    ctx->passes += 1;
    ctx->shard_direct = 1;
    prof_collect(ctx);
    return KB_OK;
}

int kb_shard_recv_buffer(kb_ctx* ctx, uint64_t n_records, void** buffer) {
    if (!ctx || !buffer) return KB_EINVAL;
    CU(cudaSetDevice(ctx->device));
    // the partitioned records live in one ping-pong buffer: receive into the other one, which then is the input of the
    // next partition level — no staging copy
    DevBuf& dst = ctx->shard_send_in_B ? ctx->entA : ctx->entB;
    TRY(ensure(ctx, dst, (n_records + 2048) * 8));
    *buffer = dst.p;
    return KB_OK;
}

int kb_shard_search(kb_ctx* ctx, uint64_t n_records, const uint64_t* piece_counts, kb_result** out) {
    if (!ctx || !out || (n_records && !piece_counts)) return KB_EINVAL;
    *out = nullptr;
    if (!ctx->configured || ctx->shard_n < 1) return fail(ctx, KB_EINVAL, "kb_configure / kb_shard_plan have not been called");
    CU(cudaSetDevice(ctx->device));
    begin_search(ctx);
    TRY(prepare_small(ctx));
    PartPlan pl = ctx->shard_plan;
    // the received records (see kb_shard_recv_buffer) are this level's input; the send buffer is free again
    DevBuf& in = ctx->shard_direct ? ctx->recvbuf : (ctx->shard_send_in_B ? ctx->entA : ctx->entB);
    DevBuf& other = ctx->shard_direct ? ctx->entA : (ctx->shard_send_in_B ? ctx->entB : ctx->entA);
    if (in.cap < (n_records + 2048) * 8) return fail(ctx, KB_EINVAL, "the receive buffer is smaller than the number of received records");
    const uint32_t nd = pl.nc[0];
    const uint32_t d_lo = shard_first_digit((uint32_t)ctx->shard_index, (uint32_t)ctx->shard_n, nd);
    const uint32_t dps = shard_first_digit((uint32_t)ctx->shard_index + 1, (uint32_t)ctx->shard_n, nd) - d_lo;
    // pieces = (source rank, level-0 digit) in arrival order: parents of level 1; pieces of one digit share a cursor row
    const uint32_t np = (uint32_t)ctx->shard_n * dps;
    std::vector<uint64_t>& hp = ctx->shard_tab_host;
    hp.assign(((size_t)np + 1) * 2 + 2, 0);                 // pstart [np + 1] u64 | ptile0 [np + 1] u32 | prow [np] u32
    uint64_t* h_start = hp.data();
    uint32_t* h_tile0 = (uint32_t*)(hp.data() + np + 1);
    uint32_t* h_row = h_tile0 + np + 1;
    uint64_t run = 0, trun = 0;
    for (uint32_t pi = 0; pi < np; pi++) {
        h_start[pi] = run; h_tile0[pi] = (uint32_t)trun; h_row[pi] = pi % dps;
        run += piece_counts[pi]; trun += (piece_counts[pi] + KB_PT_TILE - 1) / KB_PT_TILE;
    }
    h_start[np] = run; h_tile0[np] = (uint32_t)trun;
    if (run != n_records) return fail(ctx, KB_EINVAL, "piece counts do not add up to the number of received records");
    TRY(ensure(ctx, ctx->shard_tab, hp.size() * 8));
    CU(cudaMemcpyAsync(ctx->shard_tab.p, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h_pinned[10] = n_records;                           // the "record counter" the stream kernel and the root table read
    CU(cudaMemcpyAsync((uint64_t*)ctx->small.p + SM_NOUT, ctx->h_pinned + 10, 8, cudaMemcpyHostToDevice, ctx->stream));
    // local child counts: dps rows at level 0, then the usual fan-out
    { uint32_t c = dps; pl.ncl[0] = c; for (int l = 1; l < pl.levels; l++) { c <<= pl.bits[l]; pl.ncl[l] = c; } }
    const size_t tilemap_extra = (size_t)(trun + np + 2) * 4;
    TRY(ensure(ctx, ctx->plan, pl.bytes + tilemap_extra + 64));
    CU(cudaMemsetAsync(ctx->plan.p, 0, pl.bytes, ctx->stream));
    CustomParents cp{};
    cp.pstart = (const unsigned long long*)ctx->shard_tab.p;
    cp.ptile0 = (const uint32_t*)((uint64_t*)ctx->shard_tab.p + np + 1);
    cp.prow = cp.ptile0 + np + 1;
    cp.n = np;
    cp.max_tiles = trun + 1;
    uint64_t* parted = nullptr;
    HashStage hs{};
    hs.pl = &pl;
    // level-1 child counts that came with the exchange (K1's two-level histograms, summed over the source ranks by the host layer)
    const bool counts_given = pl.levels >= 2 && ctx->shard_child_set.size() == (size_t)pl.ncl[1] && n_records > 0;
    if (counts_given) {
        uint64_t total = 0;
        for (uint64_t c : ctx->shard_child_set) total += c;
        if (total != n_records) return fail(ctx, KB_EINVAL, "level-1 child counts do not add up to the number of received records");
        CU(cudaMemcpyAsync((uint8_t*)ctx->plan.p + pl.off_cnt[1], ctx->shard_child_set.data(), (size_t)pl.ncl[1] * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    TRY(run_partition(ctx, pl, in, other, n_records, 1, pl.levels, &cp, &parted, &hs.bstart, &hs.n_buckets, false, counts_given));
    ctx->shard_child_set.clear();
    ctx->lazy_now = !ctx->lo.direct;
    if (ctx->lazy_now && n_records > 0) TRY(run_prefilter(ctx, pl, parted, n_records, &hs, &parted));
    int rc = run_group(ctx, parted, n_records, out, &hs);
    prof_collect(ctx);
    return rc;
}

// ---- multi-GPU on slabs: K1 + level 0 store straight into the owners' receive slabs (the exchange IS K1's store phase), the
//      cursors are all-gathered on the device, levels >= 1 and the bucket hash run on the owner -----------------------------------
__global__ void __launch_bounds__(256) kb_shard_pend_kernel(const unsigned long long* gathered, uint32_t n_ranks, uint32_t nd0, uint32_t d_lo, uint32_t dps,
                                                            unsigned long long* pend, uint32_t* prow) {
    for (uint32_t p = blockIdx.x * 256 + threadIdx.x; p < n_ranks * dps; p += gridDim.x * 256) {
        const uint32_t src = p / dps, j = p - src * dps;
        pend[p] = gathered[(size_t)src * nd0 + d_lo + j];              // the source's cursor of (its slab of) digit d_lo + j in this buffer
        prow[p] = j;
    }
}

int kb_shard_slab_plan(kb_ctx* ctx, int n_shards, int shard_index, uint64_t total_bases, uint64_t max_rank_bases, int* n_digits,
                       uint64_t* recv_capacity_records, int* max_groups) {
    if (!ctx || n_shards < 1 || n_shards > 256 || shard_index < 0 || shard_index >= n_shards || !recv_capacity_records) return KB_EINVAL;
    if (!ctx->configured) return fail(ctx, KB_EINVAL, "kb_configure has not been called");
    const KbLayout& lo = ctx->lo;
    ctx->shard_slab = false;
    const uint64_t n_est = 2 * total_bases + 64;
    const PartPlan probe = make_plan(ctx, n_est, 0, false);
    if (!(ctx->opt_slab && ctx->opt_group_algo && lo.direct && probe.stream))
        return fail(ctx, KB_EUNSUPPORTED, "the slab exchange needs one-word records and the bucket-hash path (slab = 1, group_algo = 1)");
    int min_bits0 = 0;
    while ((1 << min_bits0) < n_shards) min_bits0++;
    const int keybits = lo.FB;
    int bb = std::min(probe.bb + (int)ctx->opt_shard_bb_extra, std::min(keybits, 24));
    bb = std::max(bb, min_bits0 + 1);
    if (keybits < min_bits0 + 1) return fail(ctx, KB_EUNSUPPORTED, "flank key too short to shard over this many GPUs");
    int bits0 = ctx->opt_shard_bits0 > 0 ? (int)ctx->opt_shard_bits0 : (bb + 1) / 2;
    bits0 = std::max(std::max(min_bits0, 1), std::min(std::min(bits0, 9), bb - 1));
    if (bb - bits0 > 18) bb = bits0 + 18;
    // make_plan with a forced level 0: (bits0, then the rest in levels of <= 9 bits); the option bucket_bits, when set, still wins
    const long long save_bb = ctx->opt_bucket_bits;
    ctx->opt_bucket_bits = ctx->opt_bucket_bits >= 0 ? ctx->opt_bucket_bits : bb;
    PartPlan pl = make_plan(ctx, n_est, bits0, false);
    ctx->opt_bucket_bits = save_bb;
    if (pl.levels < 2 || pl.levels > 3 || pl.bits[0] < min_bits0) return fail(ctx, KB_EINTERNAL, "shard slab plan");
    const uint32_t nd0 = pl.nc[0];
    const uint32_t d_lo = shard_first_digit((uint32_t)shard_index, (uint32_t)n_shards, nd0);
    const uint32_t dps = shard_first_digit((uint32_t)shard_index + 1, (uint32_t)n_shards, nd0) - d_lo;
    SlabPlan sp;
    sp.levels = pl.levels;
    sp.sym = sym_ok(ctx, pl, &sp.core_mask, n_shards);
    size_t off = 0;
    uint32_t maxnc = std::max<uint32_t>(nd0, (uint32_t)n_shards * dps);
    uint32_t loc = dps;
    for (int l = 0; l < pl.levels; l++) {
        sp.bits[l] = pl.bits[l];
        if (l == 0) { sp.nc[0] = nd0; sp.cap[0] = slab_capacity(ctx, (sp.sym ? 1 : 2) * max_rank_bases + 64, nd0, true); }   // (items / records of the largest rank)
        else { loc <<= pl.bits[l]; sp.nc[l] = loc; sp.cap[l] = slab_capacity(ctx, n_est, pl.nc[l], l + 1 < pl.levels); }
        pl.ncl[l] = l == 0 ? dps : loc;
        maxnc = std::max(maxnc, sp.nc[l]);
        sp.off_cur[l] = off; off += (size_t)sp.nc[l] * 8;
    }
    sp.off_counts = off; off += (size_t)maxnc * 8;
    sp.off_start = off; off += ((size_t)maxnc + 1) * 8;
    sp.off_part = off; off += ((size_t)maxnc / KB_PLAN_BLOCK + 2) * 16;
    sp.off_tab0 = off; off += (size_t)2 * KB_XP_MAXR * 8;
    for (int l = 0; l < pl.levels; l++) { sp.off_tile0[l] = off; off += (((size_t)std::max<uint32_t>(sp.nc[l], (uint32_t)n_shards * dps) + 2) * 4 + 7) & ~(size_t)7; }
    sp.max_tiles = (2 * total_bases / std::max(1, n_shards) * 2 + 64) / KB_PT_TILE + maxnc + (uint64_t)n_shards * dps + 2;
    sp.off_tilemap = off; off += ((size_t)sp.max_tiles * 4 + 7) & ~(size_t)7;
    sp.off_snap = off; off += ((size_t)n_shards * dps) * 8 + (((size_t)n_shards * dps * 4 + 7) & ~(size_t)7);   // parent fill levels | cursor rows
    sp.bytes = off;
    ctx->shard_sp = sp;
    ctx->shard_plan = pl;
    ctx->shard_n = n_shards;
    ctx->shard_index = shard_index;
    ctx->shard_total_bases = total_bases;
    ctx->shard_slab = true;
    if (n_digits) *n_digits = (int)nd0;
    *recv_capacity_records = (uint64_t)n_shards * dps * sp.cap[0] + 4096;
    // digit groups of the pipelined exchange: tile-aligned parent slabs (kb_shard_slab_level); three-level plans put level 2 into a buffer of its own
    if (max_groups) *max_groups = (sp.cap[0] % KB_PT_TILE == 0 && (sp.levels == 2 || (sp.levels == 3 && sp.cap[1] % KB_PT_TILE == 0)))
                                      ? (int)std::max<uint32_t>(1, nd0 / (uint32_t)n_shards) : 1;
    return KB_OK;
}

int kb_shard_slab_extract(kb_ctx* ctx, void** cursors_dev) {
    if (!ctx || !cursors_dev) return KB_EINVAL;
    if (!ctx->configured || !ctx->shard_slab) return fail(ctx, KB_EINVAL, "kb_shard_slab_plan has not been called");
    if ((int)ctx->peer_ptr.size() != ctx->shard_n) return fail(ctx, KB_EINVAL, "kb_shard_ipc_import has not been called");
    const SlabPlan& sp = ctx->shard_sp;
    const uint32_t nd0 = sp.nc[0];
    const int N = ctx->shard_n, me = ctx->shard_index;
    CU(cudaSetDevice(ctx->device));
    TRY(check_file_ids(ctx));
    begin_search(ctx);
    ctx->shard_own_done = false;
    TRY(prepare_small(ctx));
    TRY(ensure(ctx, ctx->plan, sp.bytes + 64));
    uint8_t* P = (uint8_t*)ctx->plan.p;
    // this GPU's buffers: entA = staging of the slabs bound for the other owners (slab d at d * cap0, copied out by kb_shard_slab_send),
    // children levels: level 1 (which reads the receive buffer) writes entB, level 2 entC — not entA: the copies of later digit groups still read it
    TRY(ensure_results(ctx, std::max<uint64_t>(ctx->result_cap, (uint64_t)ctx->opt_result_cap)));
    TRY(ensure(ctx, ctx->deferred, (size_t)sp.nc[sp.levels - 1] * 4 + 64));       // (the per-group launches must not reallocate it)
    size_t need[3] = {(size_t)nd0 * sp.cap[0] + 4096, 0, 0};
    for (int l = 1; l < sp.levels; l++) need[l] = (size_t)sp.nc[l] * sp.cap[l] + 4096;
    TRY(ensure(ctx, ctx->entA, need[0] * 8));
    if (need[1]) TRY(ensure(ctx, ctx->entB, need[1] * 8));
    if (need[2]) TRY(ensure(ctx, ctx->entC, need[2] * 8));
    for (int l = 1; l < sp.levels; l++) {
        kb_slab_init_kernel<<<(unsigned)std::min<uint32_t>((sp.nc[l] + 255) / 256, 1024), 256, 0, ctx->stream>>>((unsigned long long*)(P + sp.off_cur[l]), sp.nc[l], sp.cap[l]);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    // level-0 tables.  Own digits: straight into this rank's slabs of its receive buffer (slab (me, j)); the other owners' digits:
    // into the local staging buffer, from where whole digit groups travel as bulk peer copies (kb_shard_slab_send).  The cursor of
    // digit d counts in the coordinates of the OWNER's receive buffer either way, so the all-gathered cursors are the fill levels
    // the owner needs.
    std::vector<uint64_t>& t0 = ctx->scatter_host;
    t0.assign((size_t)3 * KB_XP_MAXR, 0);
    for (int o = 0; o < N; o++) {
        const uint32_t d0 = shard_first_digit((uint32_t)o, (uint32_t)N, nd0), d1 = shard_first_digit((uint32_t)o + 1, (uint32_t)N, nd0);
        const uint64_t dps_o = d1 - d0;
        for (uint32_t d = d0; d < d1; d++) {
            const uint64_t first = ((uint64_t)me * dps_o + (d - d0)) * sp.cap[0];                        // slab (me, d - d0) of owner o
            t0[d] = first + sp.cap[0];                                                                   // slab end
            t0[2 * KB_XP_MAXR + d] = first;                                                              // cursor start
            // element address that cursor value 0 stands for: the receive buffer itself, or the staging buffer shifted so that
            // the slab lands at d * cap0
            if (o == me) t0[KB_XP_MAXR + d] = (uint64_t)(reinterpret_cast<uintptr_t>(ctx->recvbuf.p) >> 3);
            else t0[KB_XP_MAXR + d] = (uint64_t)(reinterpret_cast<uintptr_t>(ctx->entA.p) >> 3) + (uint64_t)d * sp.cap[0] - first;
        }
    }
    if ((uint64_t)N * (shard_first_digit((uint32_t)me + 1, (uint32_t)N, nd0) - shard_first_digit((uint32_t)me, (uint32_t)N, nd0)) * sp.cap[0] + 2048 > ctx->recvbuf.cap / 8)
        return fail(ctx, KB_EINVAL, "the receive buffer is smaller than kb_shard_slab_plan asked for");
    CU(cudaMemcpyAsync(P + sp.off_tab0, t0.data(), (size_t)2 * KB_XP_MAXR * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(P + sp.off_cur[0], t0.data() + 2 * KB_XP_MAXR, (size_t)nd0 * 8, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t pos_lo = 0, pos_hi = 0;
    uint32_t tile0 = 0, n_tiles = 0;
    TRY(shard_extract_range(ctx, &pos_lo, &pos_hi, &tile0, &n_tiles));
    TRY(prepare_extract(ctx));
    const auto batches = extract_batches(ctx, tile0, n_tiles);
    prof_begin(ctx, "K1 extract + partition 0");
    TRY(run_extract_part(ctx, tile0, n_tiles, pos_lo, pos_hi, (uint32_t)sp.bits[0], (unsigned long long*)(P + sp.off_cur[0]),
                         (const unsigned long long*)(P + sp.off_tab0), (const unsigned long long*)(P + sp.off_tab0) + KB_XP_MAXR, batches,
                         [](int, int, uint32_t) -> int { return KB_OK; }, sp.sym, sp.core_mask));
    prof_end(ctx);
    *cursors_dev = P + sp.off_cur[0];
    return KB_OK;
}

int kb_shard_slab_send(kb_ctx* ctx, int group, int n_groups, int part, int n_parts, void* cuda_stream) {
    const bool by_peer = n_parts < 0;                         // n_parts < 0: this call serves the peers k with (k - 1) % |n_parts| == part, whole copies
    if (by_peer) n_parts = -n_parts;
    if (!ctx || n_groups < 1 || group < 0 || group >= n_groups || n_parts < 1 || part < 0 || part >= n_parts) return KB_EINVAL;
    if (!ctx->configured || !ctx->shard_slab) return fail(ctx, KB_EINVAL, "kb_shard_slab_plan has not been called");
    if ((int)ctx->peer_ptr.size() != ctx->shard_n) return fail(ctx, KB_EINVAL, "kb_shard_ipc_import has not been called");
    const SlabPlan& sp = ctx->shard_sp;
    const uint32_t nd0 = sp.nc[0];
    const uint32_t N = (uint32_t)ctx->shard_n, me = (uint32_t)ctx->shard_index;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(cudaSetDevice(ctx->device));
    // peer k of this call = rank (me + k) % N: at any time every receiver has one sender
    for (uint32_t k = 1; k < N; k++) {
        if (by_peer && (int)((k - 1) % (uint32_t)n_parts) != part) continue;
        const uint32_t o = (me + k) % N;
        const uint32_t d0 = shard_first_digit(o, N, nd0), dps_o = shard_first_digit(o + 1, N, nd0) - d0;
        const uint32_t j0 = (uint32_t)(((uint64_t)group * dps_o) / (uint32_t)n_groups), j1 = (uint32_t)(((uint64_t)(group + 1) * dps_o) / (uint32_t)n_groups);
        if (j1 <= j0) continue;
        const uint64_t* src = (const uint64_t*)ctx->entA.p + (uint64_t)(d0 + j0) * sp.cap[0];
        uint64_t* dst = (uint64_t*)ctx->peer_ptr[o] + ((uint64_t)me * dps_o + j0) * sp.cap[0];
        // whole slabs (the fill levels travel separately); this call moves byte range `part` of `n_parts` of every copy, so that
        // several streams = several copy engines share one group
        const uint64_t total = (uint64_t)(j1 - j0) * sp.cap[0];
        const uint64_t e0 = by_peer ? 0 : ((total * (uint64_t)part / (uint64_t)n_parts) & ~1ULL);
        const uint64_t e1 = (by_peer || part + 1 == n_parts) ? total : ((total * (uint64_t)(part + 1) / (uint64_t)n_parts) & ~1ULL);
        if (e1 > e0) CU(cudaMemcpyAsync(dst + e0, src + e0, (size_t)(e1 - e0) * 8, cudaMemcpyDeviceToDevice, st));
    }
    return KB_OK;
}

int kb_shard_slab_buffers(kb_ctx* ctx, void** staging, void** receive, uint64_t* slab_records, int* window_items) {
    if (!ctx || !staging || !receive || !slab_records) return KB_EINVAL;
    if (!ctx->configured || !ctx->shard_slab) return fail(ctx, KB_EINVAL, "kb_shard_slab_plan has not been called");
    *staging = ctx->entA.p; *receive = ctx->recvbuf.p; *slab_records = ctx->shard_sp.cap[0];
    if (window_items) *window_items = ctx->shard_sp.sym ? 1 : 0;
    return KB_OK;
}

int kb_shard_slab_level(kb_ctx* ctx, const void* gathered_cursors_dev, int group, int n_groups) {
    if (!ctx || !gathered_cursors_dev || n_groups < 1 || group < 0 || group >= n_groups) return KB_EINVAL;
    if (!ctx->configured || !ctx->shard_slab) return fail(ctx, KB_EINVAL, "kb_shard_slab_plan has not been called");
    const SlabPlan& sp = ctx->shard_sp;
    PartPlan pl = ctx->shard_plan;
    const uint32_t nd0 = sp.nc[0];
    const uint32_t N = (uint32_t)ctx->shard_n, me = (uint32_t)ctx->shard_index;
    const uint32_t d_lo = shard_first_digit(me, N, nd0), dps = shard_first_digit(me + 1, N, nd0) - d_lo;
    const uint32_t np = N * dps;
    const bool three = sp.levels == 3;
    if (n_groups > 1 && (sp.levels < 2 || sp.levels > 3 || sp.cap[0] % KB_PT_TILE != 0 || (three && sp.cap[1] % KB_PT_TILE != 0)))
        return fail(ctx, KB_EINVAL, "digit groups need tile-aligned parent slabs (use one group)");
    CU(cudaSetDevice(ctx->device));
    uint8_t* P = (uint8_t*)ctx->plan.p;
    unsigned long long* pend = (unsigned long long*)(P + sp.off_snap);
    uint32_t* prow = (uint32_t*)(pend + np);
    if (group == 0 && !ctx->shard_own_done) {
        kb_shard_pend_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>((const unsigned long long*)gathered_cursors_dev, N, nd0, d_lo, dps, pend, prow);
        CU(cudaGetLastError());
        CU(cudaMemsetAsync((uint64_t*)ctx->small.p + SM_NRES, 0, 7 * 8, ctx->stream));
        ctx->launches++;
    }
    const uint32_t j0 = (uint32_t)(((uint64_t)group * dps) / (uint32_t)n_groups), j1 = (uint32_t)(((uint64_t)(group + 1) * dps) / (uint32_t)n_groups);
    if (j1 <= j0) return KB_OK;
    const uint64_t n_est = 2 * ctx->shard_total_bases + 64;
    uint64_t* bufs[2] = {(uint64_t*)ctx->entA.p, (uint64_t*)ctx->entB.p};
    static const char* n1[8] = {"K2 partition 1 (group 0)", "K2 partition 1 (group 1)", "K2 partition 1 (group 2)", "K2 partition 1 (group 3)",
                                "K2 partition 1 (group 4)", "K2 partition 1 (group 5)", "K2 partition 1 (group 6)", "K2 partition 1 (group 7+)"};
    static const char* n3[8] = {"K3 bucket hash (group 0)", "K3 bucket hash (group 1)", "K3 bucket hash (group 2)", "K3 bucket hash (group 3)",
                                "K3 bucket hash (group 4)", "K3 bucket hash (group 5)", "K3 bucket hash (group 6)", "K3 bucket hash (group 7+)"};
    prof_begin(ctx, n_groups == 1 ? "K2 partition 1" : n1[std::min(group, 7)]);
    if (n_groups == 1) {
        TRY(launch_slab_level(ctx, sp, pl, 1, (const uint64_t*)ctx->recvbuf.p, bufs[1], pend, nullptr, np, prow, std::min<uint64_t>(n_est, (uint64_t)np * sp.cap[0])));
    } else {
        // (this rank's own slabs went through kb_shard_slab_own already: the group passes then cover the other sources only)
        if (ctx->shard_own_done) { if (N > 1) TRY(launch_slab_level(ctx, sp, pl, 1, (const uint64_t*)ctx->recvbuf.p, bufs[1], pend, nullptr, (N - 1) * (j1 - j0), prow, 0, j1 - j0, j0, dps, me)); }
        else TRY(launch_slab_level(ctx, sp, pl, 1, (const uint64_t*)ctx->recvbuf.p, bufs[1], pend, nullptr, N * (j1 - j0), prow, 0, j1 - j0, j0, dps));
    }
    prof_end(ctx);
    if (group == 0) ctx->alg_rec_bytes += sp.sym ? 12 : 16;
    if (three) {
        // level 2 on the level-1 slabs of this group's digits ([j0 << bits1, j1 << bits1): complete once level 1 has seen every source)
        static const char* n2[8] = {"K2 partition 2 (group 0)", "K2 partition 2 (group 1)", "K2 partition 2 (group 2)", "K2 partition 2 (group 3)",
                                    "K2 partition 2 (group 4)", "K2 partition 2 (group 5)", "K2 partition 2 (group 6)", "K2 partition 2 (group 7+)"};
        prof_begin(ctx, n_groups == 1 ? "K2 partition 2" : n2[std::min(group, 7)]);
        const unsigned long long* pend1 = (const unsigned long long*)(P + sp.off_cur[1]);
        if (n_groups == 1) {
            TRY(launch_slab_level(ctx, sp, pl, 2, bufs[1], (uint64_t*)ctx->entC.p, pend1, nullptr, sp.nc[1], nullptr, std::min<uint64_t>(n_est, (uint64_t)sp.nc[1] * sp.cap[1])));
        } else {
            const uint32_t np2 = (j1 - j0) << sp.bits[1];
            TRY(launch_slab_level(ctx, sp, pl, 2, bufs[1], (uint64_t*)ctx->entC.p, pend1, nullptr, np2, nullptr, 0, np2, j0 << sp.bits[1], 0));
        }
        prof_end(ctx);
        if (group == 0) ctx->alg_rec_bytes += 16;
    }
    // bucket hash on the children of this group's digits: buckets [j0 << bits, j1 << bits)
    const int last = sp.levels - 1;
    uint32_t cb = 0;
    for (int l = 1; l <= last; l++) cb += (uint32_t)sp.bits[l];
    HashStage hs{};
    hs.pl = &pl;
    hs.n_buckets = n_groups == 1 ? sp.nc[last] : (j1 - j0) << cb;
    hs.bucket0 = n_groups == 1 ? 0u : j0 << cb;
    hs.bend = (const unsigned long long*)(P + sp.off_cur[last]);
    hs.bcap = sp.cap[last];
    if (sp.sym) hs.bb_hash = pl.bb - sp.bits[0];
    hs.phase = 1;
    KbGroupArgs a{};
    fill_group_args(ctx, a, three ? (uint64_t*)ctx->entC.p : bufs[1], n_est);
    prof_begin(ctx, n_groups == 1 ? "K3 bucket hash" : n3[std::min(group, 7)]);
    TRY(launch_hash(ctx, a, hs));
    prof_end(ctx);
    return KB_OK;
}

int kb_shard_slab_own(kb_ctx* ctx, const void* gathered_cursors_dev) {
    if (!ctx || !gathered_cursors_dev) return KB_EINVAL;
    if (!ctx->configured || !ctx->shard_slab) return fail(ctx, KB_EINVAL, "kb_shard_slab_plan has not been called");
    const SlabPlan& sp = ctx->shard_sp;
    PartPlan pl = ctx->shard_plan;
    if (sp.levels < 2 || sp.levels > 3 || sp.cap[0] % KB_PT_TILE != 0) return fail(ctx, KB_EINVAL, "needs tile-aligned level-0 slabs (as digit groups do)");
    const uint32_t nd0 = sp.nc[0];
    const uint32_t N = (uint32_t)ctx->shard_n, me = (uint32_t)ctx->shard_index;
    const uint32_t d_lo = shard_first_digit(me, N, nd0), dps = shard_first_digit(me + 1, N, nd0) - d_lo;
    const uint32_t np = N * dps;
    CU(cudaSetDevice(ctx->device));
    uint8_t* P = (uint8_t*)ctx->plan.p;
    unsigned long long* pend = (unsigned long long*)(P + sp.off_snap);
    uint32_t* prow = (uint32_t*)(pend + np);
    kb_shard_pend_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>((const unsigned long long*)gathered_cursors_dev, N, nd0, d_lo, dps, pend, prow);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync((uint64_t*)ctx->small.p + SM_NRES, 0, 7 * 8, ctx->stream));
    ctx->launches++;
    uint64_t* bufs[2] = {(uint64_t*)ctx->entA.p, (uint64_t*)ctx->entB.p};
    prof_begin(ctx, "K2 partition 1 (own slabs)");
    // parents = slab (me, j) for every digit j of this shard: q -> me * dps + q
    if (dps) TRY(launch_slab_level(ctx, sp, pl, 1, (const uint64_t*)ctx->recvbuf.p, bufs[1], pend, nullptr, dps, prow, 0, dps, me * dps, dps));
    prof_end(ctx);
    ctx->shard_own_done = true;
    return KB_OK;
}

int kb_shard_slab_finish(kb_ctx* ctx, int* status, kb_result** out) {
    if (!ctx || !status || !out) return KB_EINVAL;
    *out = nullptr; *status = 0;
    if (!ctx->configured || !ctx->shard_slab) return fail(ctx, KB_EINVAL, "kb_shard_slab_plan has not been called");
    const SlabPlan& sp = ctx->shard_sp;
    PartPlan pl = ctx->shard_plan;
    CU(cudaSetDevice(ctx->device));
    uint8_t* P = (uint8_t*)ctx->plan.p;
    const uint64_t n_est = 2 * ctx->shard_total_bases + 64;
    uint64_t* bufs[2] = {(uint64_t*)ctx->entA.p, (uint64_t*)ctx->entB.p};
    const int last = sp.levels - 1;
    ctx->passes = sp.levels;
    HashStage hs{};
    hs.pl = &pl;
    hs.n_buckets = sp.nc[last];
    hs.bend = (const unsigned long long*)(P + sp.off_cur[last]);
    hs.bcap = sp.cap[last];
    if (sp.sym) hs.bb_hash = pl.bb - sp.bits[0];
    hs.phase = 2;
    ctx->replan_ok = ctx->opt_bucket_bits < 0 && pl.bb < std::min(ctx->lo.FB, 24);
    int rc = run_group(ctx, sp.levels == 3 ? (uint64_t*)ctx->entC.p : bufs[last & 1], n_est, out, &hs, true);
    ctx->replan_ok = false;
    prof_collect(ctx);
    if (rc == KB_REPLAN) { *status = 1; return KB_OK; }
    if (rc == KB_SLABOVF) { *status = 2; return KB_OK; }
    if (rc == KB_RESGROW) { *status = 3; return KB_OK; }
    return rc;
}

int kb_result_get(const kb_result* res, kb_result_view* view) {
    if (!res || !view) return KB_EINVAL;
    *view = res->v;
    return KB_OK;
}
int kb_result_rows(const kb_result* res, const char** text, uint64_t* n_bytes, int* row_bytes) {
    if (!res || !text || !n_bytes) return KB_EINVAL;
    *text = res->rows;
    *n_bytes = res->rows_len;
    if (row_bytes) *row_bytes = res->row_bytes;
    return KB_OK;
}
void kb_result_free(kb_result* res) { delete res; }

int kb_last_profile(const kb_ctx* ctx, const char** names, float* ms, int cap) {
    if (!ctx) return KB_EINVAL;
    int n = (int)ctx->profile.size();
    for (int i = 0; i < n && i < cap; i++) { if (names) names[i] = ctx->profile[i].first.c_str(); if (ms) ms[i] = ctx->profile[i].second; }
    return n;
}

int kb_last_counters(const kb_ctx* ctx, uint64_t* kernel_launches, uint64_t* algorithmic_bytes, int* radix_passes) {
    if (!ctx) return KB_EINVAL;
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (algorithmic_bytes) *algorithmic_bytes = ctx->alg_bytes;
    if (radix_passes) *radix_passes = ctx->passes;
    return KB_OK;
}

int kb_extract_sorted(kb_ctx* ctx, int local_index, kb_table** out) {
    if (!ctx || !out) return KB_EINVAL;
    *out = nullptr;
    if (!ctx->configured) return fail(ctx, KB_EINVAL, "kb_configure has not been called");
    if (local_index < 0 || local_index >= (int)ctx->file_starts.size()) return fail(ctx, KB_EINVAL, "no such sequence");
    CU(cudaSetDevice(ctx->device));
    begin_search(ctx);
    TRY(prepare_small(ctx));
    KbLayout lo = ctx->lo;
    lo.mix = 0;                                  // plain keys: the sorted order is the reference's LC_ALL=C order
    lo.P = (2 * lo.k + 7) / 8;                   // order by every base: left, right, then middle
    const uint64_t pos_lo = ctx->file_starts[local_index];
    const uint64_t pos_hi = local_index + 1 < (int)ctx->file_starts.size() ? ctx->file_starts[local_index + 1] : ctx->n_bases;
    const uint32_t tile0 = (uint32_t)(pos_lo / KB_K1_TB);
    const uint32_t tile1 = (uint32_t)((pos_hi + KB_K1_TB - 1) / KB_K1_TB);
    if (ctx->opt_strands && !lo.direct) return fail(ctx, KB_EUNSUPPORTED, "strands 1 / 2 (no complements, canonicals): one-word records only (k <= 28)");
    uint64_t n = 0;
    ctx->strands_now = (int)ctx->opt_strands;
    const int rc_x = run_extract(ctx, lo, tile0, tile1 - tile0, pos_lo, pos_hi, &n);
    ctx->strands_now = 0;
    TRY(rc_x);
    uint64_t* sorted = nullptr;
    if (!lo.direct) {
        // multi-word records: K1 wrote them in extraction order; sort their indices by 32-bit chunks of the base bits, least
        // significant chunk first (stable LSD), then write the records out in that order
        if (n >= (1ULL << 32)) return fail(ctx, KB_EUNSUPPORTED, "more than 2^32 records in a multi-word table");
        const uint32_t key_bits = 2 * (uint32_t)lo.k, chunks = (key_bits + 31) / 32;
        DevBuf* in = &ctx->entA; DevBuf* other = &ctx->entB;
        sorted = (uint64_t*)in->p;
        const unsigned cgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->n_sm * 16));
        for (int c = (int)chunks - 1; c >= 0 && n > 0; c--) {
            KbChunkKeyArgs ck{};
            ck.ent = (uint64_t*)in->p; ck.n = n; ck.recs = (const uint64_t*)ctx->recs.p; ck.W = (uint32_t)lo.W;
            ck.bit_pos = 32u * (uint32_t)c; ck.nbits = std::min<uint32_t>(32, key_bits - ck.bit_pos);
            kb_chunk_key_kernel<<<cgrid, 256, 0, ctx->stream>>>(ck);
            CU(cudaGetLastError());
            ctx->launches++;
            CU(cudaMemsetAsync((uint64_t*)ctx->small.p + SM_HIST, 0, 9 * 256 * 8, ctx->stream));     // run_sort accumulates its histograms
            CU(cudaMemsetAsync((uint64_t*)ctx->small.p + SM_TICKET, 0, 8 * 8, ctx->stream));
            TRY(run_sort(ctx, *in, *other, n, 4, &sorted));
            if (sorted != (uint64_t*)in->p) std::swap(in, other);
        }
        kb_table* t = new (std::nothrow) kb_table();
        if (!t) return fail(ctx, KB_ENOMEM, "host allocation failed");
        t->records.resize(n * lo.W);
        t->record_words = lo.W;
        if (n) {
            int rc = ensure(ctx, ctx->gather_out, n * lo.W * 8);
            if (rc) { delete t; return rc; }
            KbTableGatherArgs g{};
            g.ent = sorted; g.n = n; g.recs = (const uint64_t*)ctx->recs.p; g.W = (uint32_t)lo.W; g.out = (uint64_t*)ctx->gather_out.p;
            kb_table_gather_kernel<<<cgrid * 4, 256, 0, ctx->stream>>>(g);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(t->records.data(), ctx->gather_out.p, n * lo.W * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { delete t; return fail(ctx, KB_ECUDA, std::string("table download: ") + cudaGetErrorString(e)); }
            ctx->launches++;
        }
        prof_collect(ctx);
        *out = t;
        return KB_OK;
    }
    TRY(run_sort(ctx, ctx->entA, ctx->entB, n, lo.P, &sorted));
    if (ctx->opt_strands && n) {                 // every record was written twice: one of each sorted pair stays
        uint64_t* other = sorted == (uint64_t*)ctx->entA.p ? (uint64_t*)ctx->entB.p : (uint64_t*)ctx->entA.p;
        n /= 2;
        kb_every_second_kernel<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->n_sm * 16)), 256, 0, ctx->stream>>>(sorted, other, n);
        CU(cudaGetLastError());
        ctx->launches++;
        sorted = other;
    }
    kb_table* t = new (std::nothrow) kb_table();
    if (!t) return fail(ctx, KB_ENOMEM, "host allocation failed");
    t->records.resize(n);
    t->record_words = 1;
    if (n) {
        cudaError_t e = cudaMemcpyAsync(t->records.data(), sorted, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { delete t; return fail(ctx, KB_ECUDA, std::string("table download: ") + cudaGetErrorString(e)); }
    }
    prof_collect(ctx);
    *out = t;
    return KB_OK;
}

int kb_table_get(const kb_table* t, const uint64_t** records, uint64_t* n_records, int* record_words) {
    if (!t) return KB_EINVAL;
    if (records) *records = t->records.data();
    if (n_records) *n_records = t->records.size() / (size_t)std::max(t->record_words, 1);
    if (record_words) *record_words = t->record_words;
    return KB_OK;
}
void kb_table_free(kb_table* t) { delete t; }

}  // extern "C"
