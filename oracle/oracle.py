"""ctypes front-end of the C oracle (oracle/krisp_oracle.c) — TEST INFRASTRUCTURE ONLY.

FASTA parsing follows the reference (oracle/model.py: fasta_records, read_lines); the
C side restates stages A-D and the row rendering.  See krisp_oracle.c for the
reference file:line citations.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libkrisp_oracle.so")
_lib = None

KO_EKEY = -3


class OracleKeyError(KeyError):
    """The reference would raise KeyError here (character outside COMP_MAP / iupac_key)."""


def build(force=False):
    src = os.path.join(_HERE, "krisp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p, u64p, i32p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int32))
        L.ko_search.restype = ctypes.c_int
        L.ko_search.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_void_p), u64p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
        L.ko_table.restype = ctypes.c_int
        L.ko_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), u64p]
        L.ko_free.argtypes = [ctypes.c_void_p]
        L.ko_threads.restype = ctypes.c_int
        del u8p, i32p
        _lib = L
    return _lib


def threads():
    return lib().ko_threads()


def set_key_partition(part=0, nparts=1):
    """Keep only the k-mers whose (left,right) key hashes into `part` of `nparts` (memory-lean runs on big panels: the union of the
    rows over all parts is the full answer, every rule being local to one key).  (0, 1) switches it off."""
    lib().ko_set_key_partition(int(part), int(nparts))


def _pack(records_by_file):
    """records_by_file: list (per file) of list of str/bytes records -> (bases, rec_off, rec_file)."""
    chunks, offs, files = [], [0], []
    pos = 0
    for f, recs in enumerate(records_by_file):
        for r in recs:
            b = r.encode() if isinstance(r, str) else bytes(r)
            chunks.append(b)
            pos += len(b)
            offs.append(pos)
            files.append(f)
    bases = np.frombuffer(b"".join(chunks) + b"\0", dtype=np.uint8)
    return bases, np.asarray(offs, dtype=np.uint64), np.asarray(files, dtype=np.int32)


def search_records(records_by_file, labels, ingroup_labels, have_outgroup, L, D, R, omit_soft=False,
                   nthreads=0, want_groups=False):
    """Stages A-D on parsed records.  Returns (sorted rows, per-file k-mer counts[, interchange text])."""
    lb = lib()
    nfiles = len(records_by_file)
    bases, offs, files = _pack(records_by_file)
    is_in = np.asarray([1 if lab in ingroup_labels else 0 for lab in labels], dtype=np.uint8)
    lab_arr = (ctypes.c_char_p * nfiles)(*[lab.encode() for lab in labels])
    rows_p, groups_p, nrows = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64()
    counts = np.zeros(nfiles, dtype=np.uint64)
    rc = lb.ko_search(bases.ctypes.data, offs.ctypes.data, files.ctypes.data, len(files), nfiles,
                      is_in.ctypes.data, lab_arr, int(bool(have_outgroup)), L, D, R, int(bool(omit_soft)), nthreads,
                      ctypes.byref(rows_p), ctypes.byref(nrows), ctypes.byref(groups_p) if want_groups else None,
                      counts.ctypes.data)
    if rc == KO_EKEY:
        raise OracleKeyError("reference raises KeyError on this input (character outside COMP_MAP / iupac_key)")
    if rc != 0:
        raise RuntimeError(f"ko_search failed rc={rc}")
    rows = ctypes.string_at(rows_p.value).decode().splitlines()
    lb.ko_free(rows_p)
    if want_groups:
        gtext = ctypes.string_at(groups_p.value).decode()
        lb.ko_free(groups_p)
        return rows, counts, gtext
    return rows, counts


def search_files(ingroup_files, outgroup_files, L, D, R, omit_soft=False, nthreads=0, want_groups=False):
    """krisp_fasta <ingroup> --outgroup <outgroup> ... -> sorted CSV rows (no header)."""
    files = list(ingroup_files) + list(outgroup_files)
    recs = [model.fasta_records(model.read_lines(f)) for f in files]
    for r in recs:
        if model.detect_rna(r):
            raise NotImplementedError("RNA input: the reference's krisp_fasta output is undefined (render KeyError)")
    labels = ['merged_file'] if len(files) == 1 else [model.simplename(f) for f in files]   # E4
    ingroup = {model.simplename(f) for f in ingroup_files}
    return search_records(recs, labels, ingroup, len(outgroup_files) > 0, L, D, R, omit_soft, nthreads, want_groups)


def table_text(records, L, D, R, omit_soft=False):
    """Stage A+B text table of one file (the content of the reference's ``*.{k}mers`` file)."""
    lb = lib()
    rna = model.detect_rna(records)
    if rna:
        records = [s.replace('U', 'T').replace('u', 't') for s in records]
    bases, offs, _ = _pack([records])
    text_p, n = ctypes.c_void_p(), ctypes.c_uint64()
    rc = lb.ko_table(bases.ctypes.data, offs.ctypes.data, len(records), L, D, R, int(bool(omit_soft)),
                     ctypes.byref(text_p), ctypes.byref(n))
    if rc == KO_EKEY:
        raise OracleKeyError("reference raises KeyError on this input")
    if rc != 0:
        raise RuntimeError(f"ko_table failed rc={rc}")
    text = ctypes.string_at(text_p.value).decode()
    lb.ko_free(text_p)
    if rna:
        text = text.replace('T', 'U').replace('t', 'u')
    return text, int(n.value)
