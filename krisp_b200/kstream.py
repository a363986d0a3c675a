"""``kstream`` on the B200 path: one file's sorted k-mer table.

Mirrors the reference's ``kstream`` class / CLI (kstream/kstream.py:122-248, :835-952) for the
configuration ``krisp_fasta`` drives it with (``extractSortedKmers``, krisp_fasta/krisp_fasta.py:16-43):
``kmers=k, complements=True, disallow="Nn", mapsoft | omitsoft, split=[L,-R], sort=True, sortcols=[0,2]``.
That configuration is one K1 launch + one radix sort on the device (``kb_extract_sorted``; for k > 28 an
LSD sort of the record indices over 32-bit chunks of the multi-word records); the lines come back in the
reference's ``LC_ALL=C sort -t, -k1,1 -k3,3`` order.  Other option combinations
(``--canonicals``, ``--allow``, ``--expand-iupac``, several k, unsorted streaming) are the reference's
generic text pipeline and are not on the hot path: they raise ``UnsupportedError`` here (no CPU fallback).
"""
import argparse
import sys

import numpy as np

from . import ingest
from ._lib import UnsupportedError, KB_EUNSUPPORTED
from .search import Searcher, _decode_bases


def decode_table(records, L, D, R, rna=False):
    """Packed sorted records (uint64 [n] or [n, W], layout [left][right][mid][pad][id]) -> list of ``left,mid,right`` lines.

    Reproduces ``_split([L,-R])`` including its R == 0 quirk (kstream.py:824-830: the remainder lands in the
    third field and the middle stays empty)."""
    n = records.shape[0]
    if n == 0:
        return []
    w = records.reshape(n, -1)                   # one word per record, or W words for k > 28
    left = _decode_bases(w, 0, L)
    right = _decode_bases(w, 2 * L, R)
    mid = _decode_bases(w, 2 * (L + R), D)
    comma = np.full((n, 1), ord(","), dtype=np.uint8)
    if R == 0:
        m = np.concatenate([left, comma, comma, mid], axis=1)
    else:
        m = np.concatenate([left, comma, mid, comma, right], axis=1)
    if rna:
        m = np.where(m == ord("T"), np.uint8(ord("U")), m)
    width = m.shape[1]
    return [x.decode() for x in np.ascontiguousarray(m).view(f"S{width}").ravel().tolist()]


class kstream:
    """Drop-in for ``krisp.kstream.kstream`` restricted to the sorted split-table configuration."""

    def __init__(self, sequences=None, kmers=None, complements=False, canonicals=False, allow=None, disallow=None,
                 omitsoft=False, mapsoft=False, expandiupac=False, split=None, sort=False, sortmem=None,
                 sortcols=None, sortnp=1, parallel=1, device=0):
        if omitsoft is True and mapsoft is True:
            raise ValueError("can't omit and map soft masked nucleotides")            # kstream.py:212
        if complements is True and canonicals is True:
            raise ValueError("canonicals conflicts with complements")                  # kstream.py:218
        self.kmers = [kmers] if isinstance(kmers, int) else (list(kmers) if kmers is not None else None)
        self.split = [split] if isinstance(split, int) else (list(split) if split is not None else None)
        self.sequences = sequences
        self.omitsoft = bool(omitsoft)
        self.device = device
        why = None
        if self.kmers is None or len(self.kmers) != 1:
            why = "exactly one k-mer length"
        elif not complements or canonicals:
            why = "complements=True"
        elif allow is not None or expandiupac:
            why = "no allow / expandiupac"
        elif disallow is None or set(disallow) != set("Nn"):
            why = 'disallow="Nn"'
        elif not (omitsoft or mapsoft):
            why = "mapsoft or omitsoft"
        elif self.split is None or len(self.split) != 2 or self.split[0] < 0 or self.split[1] > 0:
            why = "split=[L,-R]"
        elif not sort or (sortcols is not None and list(sortcols) != [0, 2]):
            why = "sort=True, sortcols=[0,2]"
        if why:
            raise UnsupportedError(KB_EUNSUPPORTED, f"krisp_b200.kstream implements the krisp_fasta configuration only ({why})")
        k = self.kmers[0]
        self.L, self.R = self.split[0], -self.split[1]
        self.D = k - self.L - self.R
        if self.D < 0:
            raise ValueError("split lengths exceed the k-mer length")

    def _table(self, sequences):
        if sequences is None:
            raise ValueError("no sequences given")
        if isinstance(sequences, str):
            packed, rna = ingest.load_file(sequences)
        else:
            recs = list(sequences)
            # an iterable goes through the same FASTA probe as a file (kstream.py:447-452): the first item is consumed
            first = recs[0] if recs else ""
            if ">" in first:
                packed = ingest.pack_bytes(("\n".join(s.rstrip("\n") for s in recs) + "\n").encode())
            else:
                packed = ingest.pack_records([s.strip() for s in recs[1:]])
            rna = ingest.detect_rna(packed)
        if rna:
            packed = packed.copy()
            packed[packed == ord("U")] = ord("T")
            packed[packed == ord("u")] = ord("t")
        s = Searcher(self.device)
        try:
            s.configure(self.L, self.D, self.R, [1], omit_soft=self.omitsoft)
            s.clear_sequences()
            s.add_sequence(0, packed)
            recs = s.extract_sorted(0)
        finally:
            s.close()
        return recs, rna

    def __call__(self, sequences):
        recs, rna = self._table(sequences)
        return iter(decode_table(recs, self.L, self.D, self.R, rna))

    def __iter__(self):
        return iter(self.__call__(self.sequences))

    def write(self, filename, sequences=None):
        """Write the sorted table to `filename`; returns the number of lines (kstream.py:250-325)."""
        recs, rna = self._table(self.sequences if sequences is None else sequences)
        lines = decode_table(recs, self.L, self.D, self.R, rna)
        with open(filename, "w") as fh:
            for ln in lines:
                fh.write(ln)
                fh.write("\n")
        return len(lines)


def parseArgs(sys_args):
    """Same flags as kstream.py:835-922."""
    p = argparse.ArgumentParser(description="Read and parse kmers from fasta (B200 path)", prog="kstream",
                                formatter_class=argparse.RawTextHelpFormatter)
    p.add_argument("file", nargs="?", type=str, default="-")
    p.add_argument("-k", "--kmers", type=int, nargs="+")
    g = p.add_mutually_exclusive_group()
    g.add_argument("--canonicals", action="store_true")
    g.add_argument("--complements", action="store_true")
    p.add_argument("--disallow", type=str)
    p.add_argument("--allow", type=str)
    p.add_argument("--expand-iupac", action="store_true")
    p.add_argument("--omit-softmask", action="store_true")
    p.add_argument("--map-softmask", action="store_true")
    p.add_argument("--split", nargs="+", type=int)
    p.add_argument("-p", "--parallel", type=int, default=1)
    p.add_argument("-s", "--sort", action="store_true")
    p.add_argument("--sort-np", type=int, default=1)
    p.add_argument("--sort-mem", type=str)
    p.add_argument("--sort-cols", nargs="+", type=int)
    p.add_argument("--output")
    p.add_argument("--version", action="version", version="%(prog)s 1.0 (krisp_b200)")
    return p.parse_args(sys_args)


def main(argv=None):
    args = parseArgs(sys.argv[1:] if argv is None else argv)
    streamer = kstream(kmers=args.kmers, complements=args.complements, canonicals=args.canonicals, allow=args.allow,
                       disallow=args.disallow, omitsoft=args.omit_softmask, mapsoft=args.map_softmask,
                       expandiupac=args.expand_iupac, split=args.split, parallel=args.parallel, sort=args.sort,
                       sortnp=args.sort_np, sortmem=args.sort_mem, sortcols=args.sort_cols)
    src = sys.stdin.read().splitlines() if args.file == "-" else args.file
    out = open(args.output, "w") if args.output is not None else sys.stdout
    try:
        for seq in streamer(src):
            print(seq, file=out)
    finally:
        if args.output is not None:
            out.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
