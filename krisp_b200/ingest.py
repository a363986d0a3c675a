"""Host-side sequence ingest: file -> the byte layout K1 consumes.

Restates what kstream does before any k-mer exists (reference kstream/kstream.py):
``_read_file`` :458-479 (``.gz`` / ``.bz2`` by extension), ``_detect_FASTA`` :510-537 (the FIRST
line alone decides FASTA vs plain, and is then dropped because the re-chained iterator is
discarded at :450), ``_parse_FASTA`` :556-583 (strip every line, a line starting with ``>`` ends
the current record, empty records vanish), ``_parse_seqs`` :539-554 (plain input: one record per
stripped line) and ``_detect_RNA`` :481-508.

Output: one ``numpy.uint8`` array per file holding the records' characters with ONE ``\\n`` between
records (any non-letter byte separates records for K1), nothing else.
"""
import bz2
import gzip
from pathlib import Path

import numpy as np

SEP = 0x0A
_WS = (0x20, 0x09, 0x0D, 0x0B, 0x0C)   # characters str.strip() would remove besides \n


def read_bytes(filename):
    """kstream._read_file: decompress by extension (fileinput.hook_compressed)."""
    ext = Path(filename).suffix
    if ext == ".gz":
        with gzip.open(filename, "rb") as fh:
            return fh.read()
    if ext == ".bz2":
        with bz2.open(filename, "rb") as fh:
            return fh.read()
    with open(filename, "rb") as fh:
        return fh.read()


def _records_slow(data):
    """Line-by-line restatement (used when lines carry whitespace that strip() would remove)."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if not lines:
        return []
    is_fasta = b">" in lines[0]
    out = []
    if is_fasta:
        seq = []
        for line in lines[1:]:
            line = line.strip()
            if line.startswith(b">"):
                if seq:
                    out.append(b"".join(seq))
                seq = []
            elif line:
                seq.append(line)
        if seq:
            out.append(b"".join(seq))
    else:
        out = [ln.strip() for ln in lines[1:]]
    return [r for r in out if r]


def pack_bytes(data):
    """File content -> uint8 array of records joined by single separators."""
    if not data:
        return np.zeros(0, dtype=np.uint8)
    arr = np.frombuffer(data, dtype=np.uint8)
    first_nl = data.find(b"\n")
    if first_nl < 0:
        return np.zeros(0, dtype=np.uint8)                 # a single line: consumed by the FASTA probe
    is_fasta = b">" in data[:first_nl]
    if any(data.find(bytes([c])) >= 0 for c in _WS) or not is_fasta:
        recs = _records_slow(data)
        if not recs:
            return np.zeros(0, dtype=np.uint8)
        return np.frombuffer(b"\n".join(recs), dtype=np.uint8).copy()
    # fast path: FASTA without stray whitespace.  Drop line 0, drop '\n', turn each header line into one separator.
    body = arr[first_nl + 1:]
    n = body.size
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    is_nl = body == SEP
    gt = np.flatnonzero(body == ord(">"))
    if gt.size:
        # headers = '>' at a line start
        at_start = (gt == 0) | (body[np.maximum(gt, 1) - 1] == SEP)
        hs = gt[at_start]
    else:
        hs = gt
    keep = ~is_nl
    out = body
    if hs.size:
        nl_pos = np.flatnonzero(is_nl)
        idx = np.searchsorted(nl_pos, hs)                  # end of each header line
        he = np.append(nl_pos, n)[idx]
        diff = np.zeros(n + 1, dtype=np.int32)
        np.add.at(diff, hs, 1)
        np.add.at(diff, he, -1)
        in_header = np.cumsum(diff[:-1]) > 0
        keep &= ~in_header
        out = body.copy()
        out[hs] = SEP
        keep[hs] = True                                     # the '>' byte becomes the record separator
    return np.ascontiguousarray(out[keep])


def detect_rna(packed):
    """kstream._detect_RNA on the packed records: the first record holding T/t (DNA) or U/u (RNA) decides."""
    if packed.size == 0:
        return False
    is_u = (packed == ord("U")) | (packed == ord("u"))
    if not is_u.any():
        return False
    is_t = (packed == ord("T")) | (packed == ord("t"))
    pu = int(np.argmax(is_u))
    if not is_t.any():
        return True
    pt = int(np.argmax(is_t))
    if pt < pu:
        return False
    seps = np.flatnonzero(packed == SEP)
    return int(np.searchsorted(seps, pt)) > int(np.searchsorted(seps, pu))   # T only in a later record


def load_file(filename):
    """-> (packed uint8 array, is_rna)."""
    packed = pack_bytes(read_bytes(filename))
    return packed, detect_rna(packed)


def pack_records(records):
    """Iterable of str/bytes sequences (kstream's non-file input) -> packed array."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in records]
    bs = [b for b in bs if b]
    if not bs:
        return np.zeros(0, dtype=np.uint8)
    return np.frombuffer(b"\n".join(bs), dtype=np.uint8).copy()
