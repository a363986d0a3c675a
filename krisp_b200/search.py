"""Host-side driver of the B200 search: the Python mirror of krisp_fasta's stage functions.

One :class:`Searcher` owns one ``kb_ctx`` (one GPU).  ``search_files`` replaces, in one device
pass, the reference's stage sequence in ``krisp_fasta.main`` (krisp_fasta/krisp_fasta.py:236-270):
``sortedKmersSerial/Parallel`` -> ``mergeFiles`` -> ``filterAlignments``; what comes back is what
``render_output`` (outputAlignments.py:101) needs: the surviving groups, their per-column base
sets, and on request their records.

No CPU fallback: everything that touches sequence data runs in libkrisp_b200.so.
"""
import ctypes

import weakref

import numpy as np

from . import _lib, ingest
from .names import simplename

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
# 4-bit base set (bit0 A, bit1 C, bit2 G, bit3 T) -> IUPAC letter; the inverse of
# Bio.Data.IUPACData.ambiguous_dna_values as Amplicon.py:10-12 builds it (4 bases -> N)
_IUPAC = np.frombuffer(b"?ACMGRSVTWYHKDBN", dtype=np.uint8)


class SearchResult:
    """Survivor table of one search.

    Packed, as the library returns it: ``flank_words`` [n_groups, FW] (MSB-first flank bits), ``in_words`` / ``out_words``
    [n_groups, MW] (column sets), ``group_size`` [n_groups] (None unless option ``group_sizes`` or want_records asked for
    it: the rows do not need it), and with want_records ``run_offset`` [n_groups + 1] into
    ``records`` [n, W].  Decoded lazily on first access: ``left`` [n_groups, L] / ``right`` [n_groups, R] ASCII,
    ``in_mask`` / ``out_mask`` [n_groups, D] 4-bit base sets over the ingroup-labelled / the other occurrences.
    """

    def __init__(self, L, D, R, n_records=0, left=None, right=None, in_mask=None, out_mask=None, group_size=None,
                 flank_words=None, run_offset=None, records=None, stats=None, profile=None, have_outgroup=True,
                 in_words=None, out_words=None):
        self.L, self.D, self.R, self.n_records = L, D, R, n_records
        self._left, self._right, self._in_mask, self._out_mask = left, right, in_mask, out_mask
        self._own, self._views = {}, {}     # packed arrays: owned copies / views of library memory not copied yet (see _collect)
        self._n = None                      # number of groups, known without touching an array
        self.flank_word_count = self.mask_word_count = None
        self.in_words, self.out_words = in_words, out_words
        self.group_size, self.flank_words, self.run_offset, self.records = group_size, flank_words, run_offset, records
        self.stats, self.profile, self.have_outgroup = stats or {}, profile or [], have_outgroup
        self._rows_blob, self._rows_ref, self.row_bytes = None, None, 0   # CSV rows rendered on the device, ascending (left, right) (kb_result_rows)
        self._borrow = None      # (free function, kb_result handle): the packed arrays are VIEWS of library memory until _detach()

    def _packed(name):                      # noqa: N805 — property factory for the packed arrays
        def get(self):
            v = self._views.pop(name, None)
            if v is not None:                # first use: one memcpy out of the library's memory; callers only ever see owned arrays
                self._own[name] = np.array(v, copy=True)
            return self._own.get(name)

        def put(self, value):
            self._views.pop(name, None)
            self._own[name] = value
        return property(get, put)

    flank_words, in_words, out_words = _packed("flank_words"), _packed("in_words"), _packed("out_words")
    group_size, run_offset, records = _packed("group_size"), _packed("run_offset"), _packed("records")
    del _packed

    def _detach(self):
        """Copy what is still a view of the library's result arena / kb_result and release the handle.  The Searcher calls this on
        a still-living result right before its next search overwrites the arena; arrays nobody looked at are never copied when the
        result is dropped before that."""
        if self._borrow is None:
            return
        for name in list(self._views):
            self._own[name] = np.array(self._views.pop(name), copy=True)
        if self._rows_blob is None and self._rows_ref is not None:
            self._rows_blob = ctypes.string_at(*self._rows_ref)
        self._rows_ref = None
        free, handle = self._borrow
        self._borrow = None
        free(handle)

    def __del__(self):
        try:
            if self._borrow is not None:
                free, handle = self._borrow
                self._borrow = None
                free(handle)
        except Exception:
            pass

    @property
    def rows_blob(self):
        if self._rows_blob is None and self._rows_ref is not None:
            self._rows_blob = ctypes.string_at(*self._rows_ref)      # (one memcpy out of the pinned arena, on first use)
        return self._rows_blob

    @rows_blob.setter
    def rows_blob(self, value):
        self._rows_blob, self._rows_ref = value, None

    def csv_rows_bytes(self):
        """csv_rows_text() as bytes (what a caller writes to the CSV file; no decoding)."""
        blob = self.rows_blob
        if blob is not None and (blob or self.n_groups == 0 or (self.R == 0 and self.D > 0)):
            return blob
        return self.csv_rows_text().encode("ascii")

    @property
    def n_groups(self):
        if self._n is not None:
            return self._n
        if self.flank_words is not None:
            return int(self.flank_words.shape[0])
        return 0 if self._left is None else int(self._left.shape[0])

    def _decode_flanks(self):
        both = _decode_bases(self.flank_words, 0, self.L + self.R)      # left and right are adjacent in the flank bits
        if self._left is None:
            self._left = both[:, :self.L]
        if self._right is None:
            self._right = both[:, self.L:]

    @property
    def left(self):
        if self._left is None:
            self._decode_flanks()
        return self._left

    @left.setter
    def left(self, value):
        self._left = value

    @property
    def right(self):
        if self._right is None:
            self._decode_flanks()
        return self._right

    @right.setter
    def right(self, value):
        self._right = value

    @property
    def in_mask(self):
        if self._in_mask is None:
            self._in_mask = _decode_masks(self.in_words, self.D)
        return self._in_mask

    @in_mask.setter
    def in_mask(self, value):
        self._in_mask = value

    @property
    def out_mask(self):
        if self._out_mask is None:
            self._out_mask = _decode_masks(self.out_words, self.D)
        return self._out_mask

    @out_mask.setter
    def out_mask(self, value):
        self._out_mask = value

    def csv_rows_text(self):
        """The CSV body exactly as ``krisp_fasta --cores 1`` prints it after the header: one ``left,consensus,right`` line per
        region in ascending (left, right) order — rendered and ordered on the device (kb_result_rows), no host work.
        Falls back to the host decoder for results assembled by hand."""
        if self.rows_blob is not None and (self.rows_blob or self.n_groups == 0 or (self.R == 0 and self.D > 0)):
            return self.rows_blob.decode("ascii")
        from . import render
        return "".join(r + "\n" for r in render.csv_rows(self))

    def rows(self):
        """CSV rows ``left,consensus,right`` (render_csv, Amplicon.py:663-671), canonically sorted.

        Consensus column = IUPAC letter of the ingroup base set when an outgroup was given, of every
        occurrence otherwise (krisp_fasta.py:282-283, Amplicon.py:550-558; after the diagnostic
        filter every ingroup sequence is ingroup-only, so the per-column OR is the same set).
        """
        n = self.n_groups
        if n == 0:
            return []
        cons = self.in_mask if self.have_outgroup else (self.in_mask | self.out_mask)
        width = self.L + 1 + self.D + 1 + self.R
        m = np.empty((n, width + 1), dtype=np.uint8)               # row = text + "\n"
        m[:, :self.L] = self.left
        m[:, self.L] = ord(",")
        m[:, self.L + 1:self.L + 1 + self.D] = _IUPAC[cons]
        m[:, self.L + 1 + self.D] = ord(",")
        m[:, self.L + 2 + self.D:width] = self.right
        m[:, width] = ord("\n")
        # Canonical (str) order.  The flank words sort by (left, right); the row text sorts by (left, consensus, right): the two
        # differ only inside runs of equal `left`, which are rare and are re-sorted as text.
        fw = self.flank_words
        if fw is None or fw.shape[0] != n:                         # (results assembled by hand: no packed flanks)
            return sorted(m.tobytes().decode("ascii").split("\n")[:-1])
        order = np.lexsort(tuple(fw[:, j] for j in range(fw.shape[1] - 1, -1, -1))) if fw.shape[1] > 1 else np.argsort(fw[:, 0], kind="stable")
        ms = m[order]
        rows = ms.tobytes().decode("ascii").split("\n")[:-1]
        tie = np.flatnonzero(np.all(ms[1:, :self.L] == ms[:-1, :self.L], axis=1)) if n > 1 else ()
        i = 0
        while i < len(tie):
            a = int(tie[i])
            while i + 1 < len(tie) and tie[i + 1] == tie[i] + 1:
                i += 1
            b = int(tie[i]) + 2
            rows[a:b] = sorted(rows[a:b])
            i += 1
        return rows


_BYTE2LETTERS = _ACGT[(np.arange(256, dtype=np.uint16)[:, None] >> np.array([6, 4, 2, 0], dtype=np.uint16)[None, :]) & 3]   # [256, 4]


def _decode_bases(words, first_bit, n_bases):
    """[n, W] uint64 MSB-first bit strings -> [n, n_bases] ASCII letters starting at bit `first_bit`."""
    n = words.shape[0]
    if n == 0 or n_bases == 0:
        return np.zeros((n, n_bases), dtype=np.uint8)
    if first_bit % 2 == 0:
        # one table look-up per byte: 4 letters at a time
        by = np.ascontiguousarray(words.astype(">u8")).view(np.uint8).reshape(n, -1)
        b0, b1 = first_bit // 8, (first_bit + 2 * n_bases + 7) // 8
        letters = _BYTE2LETTERS[by[:, b0:b1]].reshape(n, -1)
        off = (first_bit % 8) // 2
        return letters[:, off:off + n_bases]
    bits = np.unpackbits(np.ascontiguousarray(words.astype(">u8")).view(np.uint8).reshape(n, -1), axis=1)
    sel = bits[:, first_bit:first_bit + 2 * n_bases].reshape(n, n_bases, 2)
    return _ACGT[sel[:, :, 0] * 2 + sel[:, :, 1]]


def _decode_masks(words, D):
    """[n, MW] uint32 (column c = nibble 7 - c%8 of word c/8) -> [n, D] 4-bit sets."""
    n = words.shape[0]
    if D == 0 or n == 0:
        return np.zeros((n, D), dtype=np.uint8)
    shifts = (28 - 4 * np.arange(8)).astype(np.uint32)
    nib = (words[:, :, None] >> shifts[None, None, :]) & np.uint32(0xF)
    return nib.reshape(n, -1)[:, :D].astype(np.uint8)


class Searcher:
    """One GPU context.  Not thread-safe (like the library)."""

    def __init__(self, device=0, stream=None):
        self._L = _lib.load()
        self._ctx = ctypes.c_void_p()
        rc = self._L.kb_create(int(device), ctypes.byref(self._ctx))
        if rc != _lib.KB_OK:
            msg = self._L.kb_last_error(self._ctx)
            self._L.kb_destroy(self._ctx)
            self._ctx = None
            raise _lib.KrispB200Error(rc, msg.decode() if msg else "kb_create failed")
        if stream is not None:
            # a raw cudaStream_t handle; 0 is the default stream (torch's default), not "none"
            self._check(self._L.kb_set_stream(self._ctx, ctypes.c_void_p(int(stream))))
        self._keep = []          # host buffers that must outlive the async copies
        self._last = None        # weak reference to the last SearchResult while its arrays are views of the result arena
        self.bases_added = 0     # bytes handed to add_sequence since the last clear_sequences
        self.added_ids = []      # global file ids in the order they were added
        self.lo = None

    def close(self):
        if self._ctx:
            self._detach_last()                    # (kb_destroy frees the arena a living result may still look at)
            self._L.kb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.check(self._ctx, rc)

    # ---- configuration -------------------------------------------------------------------------
    def configure(self, L, D, R, is_ingroup, omit_soft=False):
        arr = np.ascontiguousarray(np.asarray(is_ingroup, dtype=np.uint8))
        self._check(self._L.kb_configure(self._ctx, int(L), int(D), int(R), int(bool(omit_soft)), int(arr.size),
                                         arr.ctypes.data))
        self.lo = (int(L), int(D), int(R))
        self.n_files = int(arr.size)

    def set_option(self, name, value):
        self._check(self._L.kb_set_option(self._ctx, name.encode(), int(value)))

    # ---- sequences -------------------------------------------------------------------------------
    def clear_sequences(self):
        self._check(self._L.kb_clear_sequences(self._ctx))
        self._keep = []
        self.bases_added = 0
        self.added_ids = []

    def reserve(self, total_bytes):
        self._check(self._L.kb_reserve(self._ctx, int(total_bytes)))

    def add_sequence(self, file_id, data):
        """`data`: numpy uint8 array (host), or a (device_pointer, n_bytes) tuple for device-resident bases."""
        if isinstance(data, tuple):
            ptr, n = data
            self.bases_added += int(n)
            self.added_ids.append(int(file_id))
            self._check(self._L.kb_add_sequence(self._ctx, int(file_id), ctypes.c_void_p(int(ptr)), int(n), 1))
            return
        arr = np.ascontiguousarray(data, dtype=np.uint8)
        self.bases_added += int(arr.size)
        self.added_ids.append(int(file_id))
        self._keep.append(arr)
        self._check(self._L.kb_add_sequence(self._ctx, int(file_id), ctypes.c_void_p(arr.ctypes.data), int(arr.size), 0))

    def add_fasta(self, file_id, raw):
        """`raw`: the decompressed file content (bytes or a uint8 array, ideally pinned): headers and line breaks are removed on the
        device (kb_add_fasta).  Asynchronous: the copy runs on the library's copy stream, the de-lining kernels are enqueued by the
        search, per batch of arrived files, right before K1 needs them."""
        arr = raw if isinstance(raw, np.ndarray) else np.frombuffer(raw, dtype=np.uint8)
        self.bases_added += int(arr.size)
        self.added_ids.append(int(file_id))
        self._keep.append(arr)                       # the copy is asynchronous: the bytes must outlive it (pinned memory = no staging)
        self._check(self._L.kb_add_fasta(self._ctx, int(file_id), ctypes.c_void_p(arr.ctypes.data if arr.size else 0), int(arr.size)))

    def fasta_flags(self):
        """Bit 0: RNA letters seen, bit 1: whitespace only the host parser handles — since the last clear_sequences."""
        f = ctypes.c_uint()
        self._check(self._L.kb_fasta_flags(self._ctx, ctypes.byref(f)))
        return int(f.value)

    def get_sequence(self, local_index):
        """A file's bytes as K1 will see them (tests / debugging)."""
        n = ctypes.c_uint64()
        self._check(self._L.kb_get_sequence(self._ctx, int(local_index), None, 0, ctypes.byref(n)))
        out = np.zeros(int(n.value), dtype=np.uint8)
        if out.size:
            self._check(self._L.kb_get_sequence(self._ctx, int(local_index), ctypes.c_void_p(out.ctypes.data), int(out.size), ctypes.byref(n)))
        return out

    def sequence_buffer(self):
        """(device pointer, bytes) of this context's concatenated sequences (every file followed by one separator)."""
        ptr, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._check(self._L.kb_sequence_buffer(self._ctx, ctypes.byref(ptr), ctypes.byref(n)))
        return int(ptr.value or 0), int(n.value)

    def sequence_sizes(self):
        """Bytes of every added file in the sequence buffer (separator included), in the order they were added."""
        out = []
        for i in range(len(self.added_ids)):
            n = ctypes.c_uint64()
            self._check(self._L.kb_get_sequence(self._ctx, i, None, 0, ctypes.byref(n)))
            out.append(int(n.value))
        return out

    def shard_own_files(self, first_local, n_local):
        self._check(self._L.kb_shard_own_files(self._ctx, int(first_local), int(n_local)))

    def synchronize(self):
        self._check(self._L.kb_synchronize(self._ctx))
        self._keep = []

    # ---- search ------------------------------------------------------------------------------------
    def _detach_last(self):
        """The previous result (if somebody still holds it) copies its arrays out of the arena before the next search reuses it."""
        last = self._last() if self._last is not None else None
        if last is not None:
            last._detach()
        self._last = None

    def _collect(self, res_ptr, have_outgroup):
        """kb_result -> SearchResult whose packed arrays are views of the library's memory (the pinned result arena; run offsets and
        records inside the kb_result): nothing is copied unless the result outlives the next search (_detach_last)."""
        L, D, R = self.lo
        view = _lib.ResultView()
        ok = False
        try:
            self._check(self._L.kb_result_get(res_ptr, ctypes.byref(view)))
            n, FW, MW, W = int(view.n_groups), int(view.flank_words), int(view.mask_words), int(view.record_words)

            def arr(ptr, count, dtype):
                if count == 0 or not ptr:
                    return np.zeros(0, dtype=dtype)
                nbytes = count * np.dtype(dtype).itemsize
                return np.frombuffer((ctypes.c_ubyte * nbytes).from_address(ctypes.addressof(ptr.contents)), dtype=dtype)

            out = SearchResult(L=L, D=D, R=R, n_records=int(view.n_records), have_outgroup=have_outgroup)
            out._n, out.flank_word_count, out.mask_word_count = n, FW, MW
            out._views = {"flank_words": arr(view.flank, n * FW, np.uint64).reshape(n, FW),
                          "in_words": arr(view.in_mask, n * MW, np.uint32).reshape(n, MW),
                          "out_words": arr(view.out_mask, n * MW, np.uint32).reshape(n, MW),
                          "run_offset": arr(view.run_offset, n + 1, np.uint64)}
            if view.group_size:                                        # (option group_sizes / want_records; else None)
                out._views["group_size"] = arr(view.group_size, n, np.uint32)
            nrr = int(view.n_run_records)
            out._views["records"] = arr(view.records, nrr * W, np.uint64).reshape(nrr, W)
            out.stats = dict(zip(("runs", "queued_runs", "groups_in_every_file", "mixed_runs"), [int(x) for x in view.stats]))
            text, nb, rb = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_int()
            self._check(self._L.kb_result_rows(res_ptr, ctypes.byref(text), ctypes.byref(nb), ctypes.byref(rb)))
            if nb.value:
                out._rows_ref = (text.value, int(nb.value))
            else:
                out.rows_blob = b""
            out.row_bytes = int(rb.value)
            free = self._L.kb_result_free
            out._borrow = (free, ctypes.c_void_p(res_ptr.value))
            self._last = weakref.ref(out)
            ok = True
        finally:
            if not ok:
                self._L.kb_result_free(res_ptr)
        out.profile = self.last_profile()
        return out

    def search(self, have_outgroup=True):
        """Run K1 -> K2 -> K3 on the sequences added so far."""
        res = ctypes.c_void_p()
        self._detach_last()
        self._set_have_outgroup(have_outgroup)
        self._check(self._L.kb_search(self._ctx, ctypes.byref(res)))
        self._keep = []
        return self._collect(res, have_outgroup)

    def _set_have_outgroup(self, have_outgroup):
        if getattr(self, "_have_outgroup", None) != bool(have_outgroup):       # (the option re-derives the plan: only on change)
            self.set_option("have_outgroup", 1 if have_outgroup else 0)
            self._have_outgroup = bool(have_outgroup)

    # ---- multi-GPU pieces (see krisp_b200/sharded.py) -------------------------------------------------
    def shard_plan(self, n_shards, shard_index, total_bases):
        """Same call on every rank (same n_shards / total_bases): fixes the partition plan.  Returns the level-0 fan-out."""
        nd = ctypes.c_int()
        self._check(self._L.kb_shard_plan(self._ctx, int(n_shards), int(shard_index), int(total_bases), ctypes.byref(nd)))
        self._shard = (int(n_shards), int(shard_index), int(nd.value))
        return int(nd.value)

    def shard_extract(self):
        """K1 + partition level 0 -> (device pointer of the records grouped by digit, counts per shard, counts per digit)."""
        n_shards, _, nd = self._shard
        rec = ctypes.c_void_p()
        counts = (ctypes.c_uint64 * n_shards)()
        digits = (ctypes.c_uint64 * nd)()
        self._check(self._L.kb_shard_extract(self._ctx, ctypes.byref(rec), counts, digits))
        self._keep = []
        return rec.value, [int(c) for c in counts], [int(c) for c in digits]

    # fused partition + exchange over peer memory (see include/krisp_b200.h)
    def shard_ipc_export(self, capacity_records):
        buf = ctypes.create_string_buffer(64)
        self._check(self._L.kb_shard_ipc_export(self._ctx, int(capacity_records), buf))
        return buf.raw

    def shard_ipc_import(self, handles):
        blob = b"".join(handles)
        self._check(self._L.kb_shard_ipc_import(self._ctx, len(handles), blob))

    def shard_ipc_close(self):
        self._check(self._L.kb_shard_ipc_close(self._ctx))

    def shard_count(self):
        n_shards, _, nd = self._shard
        digits = (ctypes.c_uint64 * nd)()
        self._check(self._L.kb_shard_count(self._ctx, digits))
        self._keep = []
        return [int(c) for c in digits]

    def shard_child_counts(self):
        """This rank's records per level-1 child over the whole key space (after shard_count), or None when K1 did not count them."""
        n = ctypes.c_uint64()
        self._check(self._L.kb_shard_child_counts(self._ctx, None, 0, ctypes.byref(n)))
        if n.value == 0:
            return None
        out = np.zeros(int(n.value), dtype=np.uint64)
        self._check(self._L.kb_shard_child_counts(self._ctx, ctypes.c_void_p(out.ctypes.data), int(out.size), ctypes.byref(n)))
        return out

    def shard_set_child_counts(self, counts):
        arr = np.ascontiguousarray(counts, dtype=np.uint64)
        self._check(self._L.kb_shard_set_child_counts(self._ctx, ctypes.c_void_p(arr.ctypes.data), int(arr.size)))

    def shard_scatter(self, piece_base):
        arr = (ctypes.c_uint64 * len(piece_base))(*[int(c) for c in piece_base])
        self._check(self._L.kb_shard_scatter(self._ctx, arr))

    # slab exchange (K1 fused with partition level 0, peer stores into the owners' slabs; see include/krisp_b200.h)
    def shard_slab_plan(self, n_shards, shard_index, total_bases, max_rank_bases):
        """-> (level-0 fan-out, receive-buffer capacity in records, most digit groups the exchange may be cut into); raises
        UnsupportedError where the slab exchange does not apply."""
        nd, cap, mg = ctypes.c_int(), ctypes.c_uint64(), ctypes.c_int()
        self._check(self._L.kb_shard_slab_plan(self._ctx, int(n_shards), int(shard_index), int(total_bases), int(max_rank_bases),
                                               ctypes.byref(nd), ctypes.byref(cap), ctypes.byref(mg)))
        self._shard = (int(n_shards), int(shard_index), int(nd.value))
        return int(nd.value), int(cap.value), int(mg.value)

    def shard_slab_extract(self):
        """K1 + level 0 + peer stores; -> device pointer of this rank's n_digits slab cursors (u64)."""
        ptr = ctypes.c_void_p()
        self._check(self._L.kb_shard_slab_extract(self._ctx, ctypes.byref(ptr)))
        self._keep = []
        return int(ptr.value)

    def shard_slab_send(self, group, n_groups, stream, part=0, n_parts=1):
        """Bulk peer copies of one digit group (byte range `part` of `n_parts` of each) on the raw cudaStream_t `stream`."""
        self._check(self._L.kb_shard_slab_send(self._ctx, int(group), int(n_groups), int(part), int(n_parts), ctypes.c_void_p(int(stream))))

    def shard_slab_buffers(self):
        """(staging device pointer, receive-buffer device pointer, slab capacity in 8-byte elements, elements are window items)
        after shard_slab_extract."""
        a, b, c, w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_int()
        self._check(self._L.kb_shard_slab_buffers(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(w)))
        return int(a.value or 0), int(b.value or 0), int(c.value), bool(w.value)

    def shard_slab_own(self, gathered_ptr):
        """Partition level 1 on the slabs this rank filled itself (complete when K1 ends), ahead of the first digit group."""
        self._check(self._L.kb_shard_slab_own(self._ctx, ctypes.c_void_p(int(gathered_ptr))))

    def shard_slab_level(self, gathered_ptr, group, n_groups):
        """Partition level 1 + bucket hash on one digit group of the receive buffer."""
        self._check(self._L.kb_shard_slab_level(self._ctx, ctypes.c_void_p(int(gathered_ptr)), int(group), int(n_groups)))

    def shard_slab_finish(self, have_outgroup=True):
        """-> (SearchResult or None, status): 1 = plan too coarse, 2 = slab overflow, 3 = survivor table grown (see the header)."""
        res, status = ctypes.c_void_p(), ctypes.c_int()
        self._detach_last()
        self._check(self._L.kb_shard_slab_finish(self._ctx, ctypes.byref(status), ctypes.byref(res)))
        if status.value != 0 or not res.value:
            return None, int(status.value)
        return self._collect(res, have_outgroup), 0

    def set_stream(self, stream):
        """Launch on this raw cudaStream_t from now on (0 = the default stream)."""
        self._check(self._L.kb_set_stream(self._ctx, ctypes.c_void_p(int(stream))))

    def shard_recv_buffer(self, n_records):
        buf = ctypes.c_void_p()
        self._check(self._L.kb_shard_recv_buffer(self._ctx, int(n_records), ctypes.byref(buf)))
        return buf.value

    def shard_search(self, n_records, piece_counts, have_outgroup=True):
        """piece_counts: flat [source rank][digit of this shard] record counts, in arrival order."""
        arr = (ctypes.c_uint64 * max(1, len(piece_counts)))(*[int(c) for c in piece_counts])
        res = ctypes.c_void_p()
        self._detach_last()
        self._set_have_outgroup(have_outgroup)
        self._check(self._L.kb_shard_search(self._ctx, int(n_records), arr, ctypes.byref(res)))
        return self._collect(res, have_outgroup)

    # ---- kstream path ------------------------------------------------------------------------------
    def extract_sorted(self, local_index):
        """One file's k-mer table as packed records in the reference's sorted order: [n] uint64, or [n, W] for k > 28."""
        tab = ctypes.c_void_p()
        self._check(self._L.kb_extract_sorted(self._ctx, int(local_index), ctypes.byref(tab)))
        try:
            ptr, n, w = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_uint64(), ctypes.c_int()
            self._check(self._L.kb_table_get(tab, ctypes.byref(ptr), ctypes.byref(n), ctypes.byref(w)))
            if n.value == 0:
                return np.zeros(0 if w.value == 1 else (0, w.value), dtype=np.uint64)
            flat = np.ctypeslib.as_array(ptr, shape=(n.value * w.value,)).copy()
            return flat if w.value == 1 else flat.reshape(n.value, w.value)
        finally:
            self._L.kb_table_free(tab)

    # ---- bookkeeping --------------------------------------------------------------------------------
    def last_profile(self):
        names = (ctypes.c_char_p * 32)()
        ms = (ctypes.c_float * 32)()
        n = self._L.kb_last_profile(self._ctx, names, ms, 32)
        return [(names[i].decode(), float(ms[i])) for i in range(max(0, min(n, 32)))]

    def last_counters(self):
        a, b, p = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
        self._check(self._L.kb_last_counters(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(p)))
        return {"kernel_launches": int(a.value), "algorithmic_bytes": int(b.value), "radix_passes": int(p.value)}


def labels_for(ingroup_files, outgroup_files):
    """(labels, is_ingroup) per input file, in the reference's file order ``files + outgroup``
    (krisp_fasta.py:238).  Membership is by label string (krisp_fasta.py:267); a single input file is
    never merged, so its lines carry no label and read back as ``merged_file`` (intersectAmplicons.py:310)."""
    files = list(ingroup_files) + list(outgroup_files)
    labels = ["merged_file"] if len(files) == 1 else [simplename(f) for f in files]
    ingroup = {simplename(f) for f in ingroup_files}
    return labels, [1 if lab in ingroup else 0 for lab in labels]


class RNAInputError(ValueError):
    pass


def search_files(ingroup_files, outgroup_files, L, D, R, omit_soft=False, want_records=False, searcher=None,
                 options=None):
    """The diagnostic-region search of ``krisp_fasta <ingroup> --outgroup <outgroup>`` -> :class:`SearchResult`."""
    files = list(ingroup_files) + list(outgroup_files)
    labels, is_in = labels_for(ingroup_files, outgroup_files)
    have_out = len(outgroup_files) > 0
    if R == 0 and D > 0:
        # kstream's split [L, -0] treats -0 as a positive split: the middle lands in the right-hand field
        # and the diagnostic region is empty, so the filter keeps nothing (kstream.py:824-830)
        z = np.zeros((0, 0), dtype=np.uint8)
        return SearchResult(L=L, D=D, R=R, left=np.zeros((0, L), np.uint8), right=z, in_mask=np.zeros((0, D), np.uint8),
                            out_mask=np.zeros((0, D), np.uint8), group_size=np.zeros(0, np.uint32), have_outgroup=have_out,
                            flank_words=np.zeros((0, 1), np.uint64))
    own = searcher is None
    s = searcher or Searcher()
    try:
        s.configure(L, D, R, is_in, omit_soft)
        s.set_option("want_records", 1 if want_records else 0)
        for k, v in (options or {}).items():
            s.set_option(k, v)
        s.clear_sequences()
        # de-lining on the device; inputs it does not reproduce exactly (RNA, stray whitespace) raise a flag and are
        # re-ingested with the host restatement of the reference parser
        raws = [ingest.read_bytes(f) for f in files]
        s.reserve(sum(len(r) + 1 for r in raws))
        for i, raw in enumerate(raws):
            s.add_fasta(i, raw)
        res = s.search(have_outgroup=have_out)
        # de-lining happened on the device, under the search; inputs it does not reproduce byte for byte (RNA, stray whitespace) have
        # raised a flag by now and are re-ingested with the host restatement of the reference parser — rare, so checked afterwards
        if s.fasta_flags():
            s.clear_sequences()
            packed = []
            for f, raw in zip(files, raws):
                arr = ingest.pack_bytes(raw)
                if ingest.detect_rna(arr):
                    raise RNAInputError(f"{f}: RNA input — the reference's krisp_fasta output is undefined for it "
                                        "(U-tables never intersect DNA tables and the renderer raises KeyError)")
                packed.append(arr)
            s.reserve(sum(a.size + 1 for a in packed))
            for i, arr in enumerate(packed):
                s.add_sequence(i, arr)
            res = s.search(have_outgroup=have_out)
        del raws
        res.labels = labels
        res.is_ingroup = is_in
        return res
    finally:
        if own:
            s.close()
