/*
 * krisp_b200.h — C ABI of libkrisp_b200.so: the B200 (sm_100a) implementation of krisp_fasta's
 * diagnostic-region search hot path (k-mer extraction -> grouping by conserved flanks: radix
 * partition + per-bucket hash, a full radix sort only for the kstream tables -> intersection over
 * all files -> diagnostic filter).
 *
 * The reference (grunwaldlab/krisp 0.1.6) is pure Python and has no FFI for this path; the
 * boundary it crosses today is its Python stage functions and the text k-mer files between them.
 * Each entry point below names the reference interface it replaces (paths relative to
 * src/krisp/ in the reference).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions: plain C types only; every call returns KB_OK (0) or a negative KB_E* code and
 * leaves a message retrievable with kb_last_error(); no exceptions or aborts cross the ABI.
 * One kb_ctx drives one GPU (one process per GPU; multi-GPU = one ctx per rank, see kb_shard_*).
 * A ctx is not thread-safe.  The library owns all device memory and the host result buffers
 * until the matching *_free / kb_destroy; input pointers are borrowed for the duration of the
 * call.  Output *content* is deterministic.  The survivor TABLE comes in no particular order (the
 * host canonical-sorts where it compares sets); the CSV rows of kb_result_rows are in ascending
 * (left, right) order, the reference's order with --cores 1.
 */
#ifndef KRISP_B200_H
#define KRISP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_OK 0
#define KB_EINVAL -1        /* bad argument / unsupported parameter combination */
#define KB_ECUDA -2         /* CUDA runtime error (message has the CUDA error string) */
#define KB_ENOMEM -3        /* device or host allocation failed */
#define KB_EUNSUPPORTED -4  /* valid in the reference, not supported here (documented in DESIGN.md) */
#define KB_EINTERNAL -5     /* invariant violated */

typedef struct kb_ctx kb_ctx;
typedef struct kb_result kb_result;
typedef struct kb_table kb_table;

/* Library / build identification: "krisp_b200 <version> sm_100a". */
const char* kb_version(void);

/* Create a context on CUDA device `device` (per-process ordinal). */
int kb_create(int device, kb_ctx** out);
void kb_destroy(kb_ctx* ctx);
const char* kb_last_error(const kb_ctx* ctx);

/* Launch all work on this cudaStream_t (e.g. torch's current stream, so that library work orders with NCCL).
 * NULL = the (legacy) default stream, which is also torch's default; KB_STREAM_OWN = the ctx's own
 * non-blocking stream (the initial setting). */
#define KB_STREAM_OWN ((void*)(intptr_t)-1)
int kb_set_stream(kb_ctx* ctx, void* cuda_stream);

/*
 * Search parameters.  Replaces the argument deduction of krisp_fasta.py:178-213 (done by the host
 * layer) and the kstream options fixed by extractSortedKmers (krisp_fasta.py:16-43):
 *   L, D, R      --conserved-left / --diagnostic / --conserved-right (k = L+D+R <= 252)
 *   soft_mode    0 = map soft-masked bases to upper case (default), 1 = --omit-soft
 *   n_files      number of input files over ALL GPUs (ingroup + outgroup), ids 0..n_files-1 (<= 256)
 *   is_ingroup   n_files bytes; 1 iff simplename(file) is in the ingroup label set
 *                (membership is by label: krisp_fasta.py:267, shared.py:58-73)
 */
int kb_configure(kb_ctx* ctx, int L, int D, int R, int soft_mode, int n_files, const uint8_t* is_ingroup);

/*
 * Tuning / diagnostics knobs:
 *   "group_algo"    1 (default) = radix partition + per-bucket hash aggregation; 0 = radix sort + segmented pass
 *   "bucket_bits"   top bits of the (mixed) flank key / flank hash the partition separates (-1 = from the input size)
 *   "hash_slots_log2" log2 of the shared-memory hash table size (0 = default); "hash_stream" 1 = TMA-fed persistent kernel
 *   "sort_bits"     (group_algo 0) bits of the key the radix sort orders by (default 32; fewer bits = fewer passes,
 *                   residual collisions are resolved exactly in the group pass)
 *   "mix"           1 (default) = store the flank key mixed by a bijection (uniform digits)
 *   "want_records"  1 = also return every record of the surviving groups' runs (for --out_align)
 *   "profile"       1 = time each stage with CUDA events (kb_last_profile)
 *   "result_cap"    initial capacity of the survivor table (it grows and the group pass re-runs on overflow)
 *   "slab"          1 (default) = one-word records: K1 fused with partition level 0 into fixed-capacity slabs (kb_extract_part.cuh);
 *                   a slab overflow (very repetitive input) repeats the search on the exact, histogram-based path.  "slab_cap" forces
 *                   the slab capacity (tests).  "hash_warp" 1 (default) = bucket hash with per-warp streaming (kb_hash_warp.cuh),
 *                   "hash_shared" 1 / 0 / -1 = one table per CTA / per warp / by table size
 *   "sym"           1 = slab path with a strand-symmetric level 0 (0 = records; -1, the default = only on >= 4 GPUs, where it halves the exchange): both records of a window share the level-0 digit, so level 0
 *                   (and the multi-GPU exchange) moves one 8-byte window item per window; records are formed by level 1
 *                   (csrc/kb_extract_sym.cuh).  Needs >= 2 partition levels and a core of enough bases, else the record path is used
 *   "group_sizes"   1 = fill kb_result_view.group_size (one more read of every survivor's bucket; implied by want_records).
 *                   "rank_rows" 1 (default) = up to 8192 survivors are ordered by counting ranks (one launch) instead of the chunked radix sort.
 *                   "hash_warps" 8 / 10 = warps per CTA of the shared-table bucket hash (0, the default: 8)
 *   "batch_level0"  1 (default) = with host buffers in flight K1 (+ partition levels 0 and 1) run per batch of arrived files
 *   "shard_bb_extra" bucket bits added to the sharded slab plan (kb_shard_slab_search status 1)
 */
int kb_set_option(kb_ctx* ctx, const char* name, long long value);

/*
 * Sequence ingest.  Replaces the per-record strings produced by kstream._parse_FASTA
 * (kstream/kstream.py:556-583).  One call per input file: `bytes` = the bases of the file as ASCII,
 * one separator byte (any byte that is not a letter, e.g. '\n') between FASTA records; headers and
 * line breaks already removed.  `on_device` != 0 means `bytes` is a device pointer.  The copy is
 * asynchronous on the ctx stream; a host buffer must stay valid until the next synchronising call
 * (kb_search / kb_synchronize) and should be pinned for the copy to be truly asynchronous.
 */
int kb_clear_sequences(kb_ctx* ctx);
int kb_reserve(kb_ctx* ctx, uint64_t total_bytes);
int kb_add_sequence(kb_ctx* ctx, int file_id, const uint8_t* bytes, uint64_t n_bytes, int on_device);
int kb_synchronize(kb_ctx* ctx);

/*
 * GPU-side FASTA de-lining.  Replaces the host loop of kstream._parse_FASTA (kstream/kstream.py:556-583; plain input:
 * _parse_seqs :539-554): `bytes` = the whole decompressed file as read (host pointer, borrowed for the call).  The first
 * line is the reference's FASTA probe (kstream.py:510-537: a '>' in it means FASTA; it is consumed either way, :450); header
 * lines, line breaks and CRs before them are removed on the device and every header becomes one record separator.
 * kb_fasta_flags: bit 0 = a 'U'/'u' was seen (RNA, kstream.py:481), bit 1 = whitespace that str.strip() would remove but the
 * device path does not — since the last kb_clear_sequences; when a bit is set the caller re-ingests with its host parser
 * (kb_add_sequence).  kb_get_sequence reads a file's bytes back as K1 will see them (tests / debugging; out == NULL: size only).
 */
int kb_add_fasta(kb_ctx* ctx, int file_id, const uint8_t* bytes, uint64_t n_bytes);
int kb_fasta_flags(kb_ctx* ctx, unsigned int* flags);
int kb_get_sequence(kb_ctx* ctx, int local_index, uint8_t* out, uint64_t cap, uint64_t* n_bytes);

/*
 * The whole search on one GPU.  Replaces, in one call:
 *   extractSortedKmers  krisp_fasta.py:16  (kstream.write + GNU sort, kstream.py:250-325, :83-119)
 *   mergeFiles          intersectAmplicons.py:232 (tree of intersectSortedStreams, shared.py:321)
 *   filterAlignments    filterAlignments.py:31  (ingroupUniqueColumns, Amplicon.py:495)
 * Blocking.  The result holds the surviving groups: flank words, per-column ingroup / outgroup
 * base sets (what consensus(), Amplicon.py:550, needs), and optionally every record of the
 * surviving groups' runs (what render_alignment, Amplicon.py:598, needs).
 */
int kb_search(kb_ctx* ctx, kb_result** out);

/*
 * Multi-GPU (one ctx per rank).  Every rule of the search is local to one (left,right) key, so records are sharded by
 * the TOP bits of the mixed flank key: level 0 of the radix partition (2^bits0 "digits") is done where the records are
 * extracted, every digit is owned by one shard (contiguous digit ranges), the exchange moves each level-0 bucket to its
 * owner, and the remaining partition levels + the bucket hash run there.  The reference has no counterpart (single host,
 * krisp_fasta.py:86-123); the file fan-out mirrors sortedKmersParallel: files are independent extraction units.
 *   kb_shard_plan         all ranks call it with the same n_shards and total_bases (bases over ALL ranks) so that they
 *                         derive the same plan; shard_index = this rank.  *n_digits = level-0 fan-out.
 *   kb_shard_extract      K1 on this rank's files + partition level 0.  *records = device pointer to the records
 *                         (8 bytes each) grouped by digit, hence by destination shard; shard_counts[s] = records bound
 *                         for shard s; digit_counts[d] = records of digit d (n_digits entries; host arrays).
 *   kb_shard_recv_buffer  device buffer for `n_records` incoming records; the host layer fills it with an all-to-all
 *                         (torch.distributed / NCCL) in source-rank order.
 *   kb_shard_search       remaining partition levels + bucket hash over the received records.  piece_counts[src * dps + j]
 *                         = records received from rank `src` with digit (first digit of this shard + j), dps = digits of
 *                         this shard = first(shard_index + 1) - first(shard_index), first(s) = s * n_digits / n_shards.
 */
int kb_shard_plan(kb_ctx* ctx, int n_shards, int shard_index, uint64_t total_bases, int* n_digits);
/*
 * Multi-word records (k > 28) on several GPUs: the 8-byte elements that travel carry (flank hash, strand, window start), and
 * the owner rebuilds the records of the occurrences that are left after the hash filter from the sequence bytes — so every
 * rank holds ALL sequences (the host layer all-gathers them, krisp_b200/sharded.py:replicate_sequences, and adds them in the
 * same order everywhere) and extracts only its own files:
 *   kb_sequence_buffer    device pointer / size of this rank's concatenated sequence bytes (every file followed by one separator)
 *   kb_shard_own_files    K1 of the kb_shard_* calls covers only local files [first_local, first_local + n_local)
 *                         (n_local < 0: all; reset by kb_clear_sequences)
 */
/*
 * Optional: level-1 child counts travel with the digit counts, and the owner skips its histogram pass over the received records.
 *   kb_shard_child_counts      after kb_shard_count: this rank's records per child of level 1 over the WHOLE key space
 *                              (*n = n_digits << bits of level 1 entries, child = digit << bits1 | next digit; *n = 0 when K1 did not
 *                              count them: the first two levels must have 10..16 bits together; counts == NULL: size only)
 *   kb_shard_set_child_counts  before kb_shard_search: the counts of THIS shard's digits summed over all source ranks
 *                              (dps << bits1 entries, in digit order); cleared by kb_shard_search and kb_shard_count
 */
int kb_shard_child_counts(kb_ctx* ctx, uint64_t* counts, uint64_t cap, uint64_t* n);
int kb_shard_set_child_counts(kb_ctx* ctx, const uint64_t* counts, uint64_t n);
int kb_sequence_buffer(kb_ctx* ctx, void** device_bytes, uint64_t* n_bytes);
int kb_shard_own_files(kb_ctx* ctx, int first_local, int n_local);

/*
 * Fused partition + exchange over NVLink peer memory (the GPUs of one box; one process per GPU).  Instead of partitioning
 * locally and calling an all-to-all, partition level 0 stores every digit's run straight into the owner's receive buffer
 * (coalesced peer stores), so the transfer overlaps the partition tile by tile:
 *   kb_shard_ipc_export   (re)allocate this rank's receive buffer for `capacity_records` and export its CUDA IPC handle
 *                         (64 bytes); ranks exchange the handles (torch.distributed all_gather) ...
 *   kb_shard_ipc_import   ... and map each other's buffers (handles[r] = rank r's; the own entry is not opened).
 *   kb_shard_count        K1 on this rank's files (level-0 histogram fused): digit_counts[d], n_digits entries.  The host
 *                         layer all-gathers them; that tells every rank where its piece of every digit starts in the
 *                         owner's buffer (pieces are laid out by source rank, then digit).
 *   kb_shard_scatter      partition level 0 with piece_base[d] = element offset of this rank's piece of digit d in its
 *                         owner's receive buffer.  After a barrier across ranks, kb_shard_search runs on the receive buffer.
 */
int kb_shard_ipc_export(kb_ctx* ctx, uint64_t capacity_records, uint8_t* handle64);
int kb_shard_ipc_import(kb_ctx* ctx, int n_ranks, const uint8_t* handles);
/* Unmap the peers' buffers.  When a receive buffer has to grow: every rank closes, barrier, then export / import again (exported
 * memory must not be freed while another process maps it). */
int kb_shard_ipc_close(kb_ctx* ctx);
int kb_shard_count(kb_ctx* ctx, uint64_t* digit_counts);
int kb_shard_scatter(kb_ctx* ctx, const uint64_t* piece_base);
int kb_shard_extract(kb_ctx* ctx, void** records, uint64_t* shard_counts, uint64_t* digit_counts);
int kb_shard_recv_buffer(kb_ctx* ctx, uint64_t n_records, void** buffer);
int kb_shard_search(kb_ctx* ctx, uint64_t n_records, const uint64_t* piece_counts, kb_result** out);

/*
 * Multi-GPU on slabs (one-word records; the default exchange): K1 is fused with partition level 0 (csrc/kb_extract_part.cuh) and
 * leaves every level-0 digit's records in a fixed-capacity slab — slab (source rank, digit) of the OWNER's receive buffer.  No
 * count exchange before the data and no separate partition pass: the slabs of a group of digits travel as bulk peer copies over
 * NVLink while the owner already runs level 1 + the bucket hash on the previous group.  The fill levels (cursors) stay on the
 * device and are all-gathered there (torch.distributed / NCCL on device tensors).
 *   kb_shard_slab_plan     same call on every rank (total_bases = bases over all ranks, max_rank_bases = the largest rank's share:
 *                          it sizes the slabs).  *recv_capacity_records = what kb_shard_ipc_export must allocate; *max_groups = into how
 *                          many digit groups the exchange may be cut (1 = no pipelining: small inputs, three-level plans).
 *                          KB_EUNSUPPORTED for multi-word records / when the slab path is switched off: use kb_shard_count ... instead.
 *   kb_shard_slab_extract  K1 + level 0 on this rank's files: the slabs of its own digits go straight into its receive buffer, the
 *                          others into a local staging buffer.  *cursors_dev = device array of n_digits u64 (this rank's fill level
 *                          of its slab of every digit, in the coordinates of the owner's buffer) for the all-gather.
 *   kb_shard_slab_send     enqueue, on `cuda_stream`, the bulk peer copies (copy engines over NVLink) of digit group `group` of
 *                          `n_groups`: for every other owner the staged slabs of the group's digits as ONE contiguous copy (of which
 *                          this call moves byte range `part` of `n_parts`; n_parts < 0: the whole copies of the peers k with
 *                          (k - 1) % |n_parts| == part — several streams keep several copy engines busy).  The
 *                          host layer runs the groups in order on a copy stream and puts a tiny collective behind each group, after
 *                          which every rank's copies of that group have landed.
 *   kb_shard_slab_own      optional, before the first kb_shard_slab_level of a pipelined search (n_groups > 1): partition level 1 on the
 *                          slabs this rank filled itself — they are complete when K1 ends, so this runs while the first digit group
 *                          is still on the wire; kb_shard_slab_level then covers the other source ranks only.
 *   kb_shard_slab_level    gathered_cursors_dev = device array [n_ranks][n_digits] (all-gather result, rank-major).  Partition
 *                          level 1 + the bucket hash on the slabs of digit group `group` in this rank's receive buffer — group g
 *                          is processed while group g + 1 is still in flight.  (Three-level plans and small inputs: one group.)
 *   kb_shard_slab_finish   deferred buckets, group sizes, rows, download.  *status: 0 = result in *out, 1 = the plan was too coarse
 *                          for this input (option "shard_bb_extra" + 2 on EVERY rank, plan again), 2 = a slab overflowed (repetitive
 *                          input: use the exact exchange, kb_shard_count ...), 3 = the survivor table was too small and has been
 *                          grown (search again).  The host layer all-reduces the status so that all ranks take the same decision
 *                          (krisp_b200/sharded.py:slab_search).
 */
int kb_shard_slab_plan(kb_ctx* ctx, int n_shards, int shard_index, uint64_t total_bases, uint64_t max_rank_bases, int* n_digits,
                       uint64_t* recv_capacity_records, int* max_groups);
int kb_shard_slab_extract(kb_ctx* ctx, void** cursors_dev);
int kb_shard_slab_send(kb_ctx* ctx, int group, int n_groups, int part, int n_parts, void* cuda_stream);
/* For a host layer that moves the digit groups itself (e.g. one NCCL all-to-all per group instead of kb_shard_slab_send), valid after
 * kb_shard_slab_extract: the staging buffer (slab of digit d at d * slab_records elements), this rank's receive buffer (slab
 * (source rank s, own digit j) at (s * own_digits + j) * slab_records) and the slab capacity in 8-byte elements.  *window_items = 1:
 * the elements are window items, one per k-mer window (strand-symmetric level 0, csrc/kb_extract_sym.cuh: half the bytes travel);
 * 0: records, two per window. */
int kb_shard_slab_buffers(kb_ctx* ctx, void** staging, void** receive, uint64_t* slab_records, int* window_items);
int kb_shard_slab_own(kb_ctx* ctx, const void* gathered_cursors_dev);
int kb_shard_slab_level(kb_ctx* ctx, const void* gathered_cursors_dev, int group, int n_groups);
int kb_shard_slab_finish(kb_ctx* ctx, int* status, kb_result** out);

/* Result accessors: borrowed pointers.  The survivor table (flank, masks, group sizes) and the rows live in the context's pinned
 * result arena — valid until kb_result_free OR the next search on the same context, whichever comes first (copy what must outlive
 * it); run_offset / records (want_records) belong to the result and live until kb_result_free. */
typedef struct {
    uint64_t n_groups;          /* surviving (left,right) groups                                  */
    uint64_t n_records;         /* valid k-mer occurrences processed (both strands, all files)    */
    uint64_t n_run_records;     /* records returned in `records` (0 unless want_records)          */
    int32_t L, D, R;
    int32_t flank_words;        /* 64-bit words per flank: bits MSB-first, left then right        */
    int32_t mask_words;         /* 32-bit words per side: column c = nibble (7 - c%8) of word c/8 */
    int32_t record_words;       /* 64-bit words per record: [left][right][mid][pad][file id : 8]  */
    int32_t reserved;
    const uint64_t* flank;      /* [n_groups][flank_words]                                        */
    const uint32_t* in_mask;    /* [n_groups][mask_words]  bit0 A, bit1 C, bit2 G, bit3 T         */
    const uint32_t* out_mask;   /* [n_groups][mask_words]                                         */
    const uint32_t* group_size; /* [n_groups] records in the group; NULL unless option "group_sizes"
                                   or "want_records" is set (the rows do not need them)           */
    const uint64_t* run_offset; /* [n_groups + 1] range of the group's run in `records`           */
    const uint64_t* records;    /* [n_run_records][record_words]; a run may hold records of other
                                   flank keys too (prefix sort): filter by the flank bits          */
    uint64_t stats[4];          /* groups, buckets (or queued runs), groups present in every file, bucket splits (or mixed runs) */
} kb_result_view;
int kb_result_get(const kb_result* res, kb_result_view* view);
/*
 * The survivors as CSV rows `left,consensus,right\n` (render_csv, Amplicon.py:663-671; consensus :550-558), rendered on the device in
 * ascending (left, right) order — the row order of the reference with --cores 1 (outputAlignments.py:101-162).  Fixed width:
 * *row_bytes = L + D + R + 3, *n_bytes = n_groups * row_bytes (no header, not NUL-terminated).  Option "render_rows" (default 1)
 * switches it off; option "have_outgroup" (default 1) = an --outgroup was given: the consensus is the ingroup's, else every
 * occurrence's (krisp_fasta.py:282-283).  Empty when R == 0 < D (the reference prints no rows then, kstream.py:824-830).
 */
int kb_result_rows(const kb_result* res, const char** text, uint64_t* n_bytes, int* row_bytes);
void kb_result_free(kb_result* res);

/* Per-stage device times of the last search in ms (needs option "profile"=1): fills up to `cap`
 * entries of names/ms, returns the number of stages. */
int kb_last_profile(const kb_ctx* ctx, const char** names, float* ms, int cap);

/* Counters of the last search: kernels launched, algorithmic bytes the kernels read+wrote, radix passes. */
int kb_last_counters(const kb_ctx* ctx, uint64_t* kernel_launches, uint64_t* algorithmic_bytes, int* radix_passes);

/*
 * kstream path: one file's k-mers as a packed table sorted like the reference's `*.{k}mers` file
 * (LC_ALL=C order on left, right, then middle).  Replaces kstream.write + sortInPlace
 * (kstream/kstream.py:250-325, :83-119); the count equals kstream.write's return value.
 * `local_index` = order of the kb_add_sequence call.  Records are one 64-bit word each for 2k + 8 <= 64, else W words
 * (kb_table_get: record_words).  Option "strands" selects the stream the table holds: 0 (default) = every window and its reverse
 * complement (--complements, kstream.py:644-677), 1 = the windows only (kstream's default), 2 = the alphabetically first of the
 * two (--canonicals, :679-694); 1 and 2 need one-word records (k <= 28).  A layout (L, D, R) = (k, 0, 0) gives the unsplit table
 * (whole k-mers in LC_ALL=C order).
 */
int kb_extract_sorted(kb_ctx* ctx, int local_index, kb_table** out);
int kb_table_get(const kb_table* t, const uint64_t** records, uint64_t* n_records, int* record_words);
void kb_table_free(kb_table* t);

#ifdef __cplusplus
}
#endif
#endif
