"""``kstream`` on the B200 path: one file's sorted k-mer table.

Mirrors the reference's ``kstream`` class / CLI (kstream/kstream.py:122-248, :835-952) for the sorted tables the device
produces — one K1 launch + one radix sort (``kb_extract_sorted``; for k > 28 an LSD sort of the record indices over 32-bit
chunks of the multi-word records):

* the configuration ``krisp_fasta`` drives it with (``extractSortedKmers``, krisp_fasta/krisp_fasta.py:16-43):
  ``kmers=k, complements=True, disallow="Nn", mapsoft | omitsoft, split=[L,-R], sort=True, sortcols=[0,2]`` — lines in the
  reference's ``LC_ALL=C sort -t, -k1,1 -k3,3`` order;
* the strand choices: ``complements`` (windows + reverse complements, :644-677), neither (kstream's default: the windows only),
  ``canonicals`` (the alphabetically first of the two, :679-694) — the last two for k <= 28;
* no ``split`` (whole k-mers, plain ``LC_ALL=C sort``);
* ``allow`` = A, C, G, T (any case): exactly what the 2-bit records can hold, so no k-mer is lost or kept differently from the
  reference; ``disallow`` = letters outside ACGT (e.g. "Nn"; other IUPAC letters then break k-mers here while the reference
  keeps them: divergence E1).

What stays the reference's generic text pipeline and raises ``UnsupportedError`` here (no CPU fallback): unsorted streaming
(the device emits records in tile order), several k at once, ``expandiupac``, neither ``mapsoft`` nor ``omitsoft`` (lower-case
letters kept as such), ``allow`` sets with other letters, other ``split`` / ``sortcols`` shapes.
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
        up = lambda chars: {c for c in chars if not c.islower()}                       # (k-mers are upper case by the time allow / disallow see them)
        if self.kmers is None or len(self.kmers) != 1:
            why = "exactly one k-mer length"
        elif expandiupac:
            why = "no expandiupac"
        elif allow is not None and up(allow) != set("ACGT"):
            why = 'allow="ACGT"'
        elif allow is None and (disallow is None or up(disallow) & set("ACGT") or "N" not in up(disallow)):
            why = 'disallow="Nn" (letters outside ACGT) or allow="ACGT"'
        elif allow is not None and disallow is not None and up(disallow) & set("ACGT"):
            why = "disallow without A, C, G, T"
        elif not (omitsoft or mapsoft):
            why = "mapsoft or omitsoft"
        elif self.split is not None and (len(self.split) != 2 or self.split[0] < 0 or self.split[1] > 0):
            why = "split=[L,-R] or no split"
        elif not sort:
            why = "sort=True"
        elif self.split is not None and (sortcols is None or list(sortcols) != [0, 2]):
            why = "sortcols=[0,2] with split=[L,-R]"
        elif self.split is None and sortcols is not None and list(sortcols) != [0]:
            why = "no sortcols without split"
        if why:
            raise UnsupportedError(KB_EUNSUPPORTED, f"krisp_b200.kstream builds sorted tables on the device only ({why})")
        k = self.kmers[0]
        self.strands = 0 if complements else (2 if canonicals else 1)
        if self.split is None:
            self.L, self.R = k, 0                                                      # whole k-mers: the sort key is the k-mer
        else:
            self.L, self.R = self.split[0], -self.split[1]
        self.D = k - self.L - self.R
        if self.D < 0:
            raise ValueError("split lengths exceed the k-mer length")
        if self.strands and 2 * k + 8 > 64:
            raise UnsupportedError(KB_EUNSUPPORTED, "krisp_b200.kstream: tables without complements / with canonicals need k <= 28")

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
            s.set_option("strands", self.strands)
            s.clear_sequences()
            s.add_sequence(0, packed)
            recs = s.extract_sorted(0)
        finally:
            s.close()
        return recs, rna

    def _lines(self, recs, rna):
        lines = decode_table(recs, self.L, self.D, self.R, rna)
        return [ln[:self.L] for ln in lines] if self.split is None else lines          # (no split: the bare k-mer)

    def __call__(self, sequences):
        recs, rna = self._table(sequences)
        return iter(self._lines(recs, rna))

    def __iter__(self):
        return iter(self.__call__(self.sequences))

    def write(self, filename, sequences=None):
        """Write the sorted table to `filename`; returns the number of lines (kstream.py:250-325)."""
        recs, rna = self._table(self.sequences if sequences is None else sequences)
        lines = self._lines(recs, rna)
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
