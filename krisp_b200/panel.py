"""Seeded synthetic multi-genome panels.

Host-side numpy only; used by bench.py, the tests and the golden-vector
generator so that the GPU path, the oracle and the live reference all see
identical bytes.  The generator follows the recipe the survey fixed:

* ancestor = iid uniform ACGT from ``default_rng(seed)``;
* group SNP sites every ``snp_every`` bp (pos snp_every//2 + snp_every*j):
  ingroup genomes carry rot(A[p]) (A->C->G->T->A), outgroup genomes keep A[p];
  at 10 % of the sites the odd-indexed ingroup genomes carry rot^2(A[p])
  instead (exercises the IUPAC consensus);
* per-genome private substitutions iid at ``noise`` (``default_rng(seed+1000+i)``);
* per genome ``n_runs`` runs of ``run_len`` N, one ``dup_len`` segment
  duplicated elsewhere, ``soft_frac`` of the bases lower-cased in
  ``soft_block`` bp blocks;
* each genome cut into ``n_records`` FASTA records, 80-column lines.
"""
from dataclasses import dataclass, field
import gzip
import os

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


@dataclass
class Genome:
    name: str
    is_ingroup: bool
    records: list = field(default_factory=list)   # list of np.uint8 arrays (ASCII bases)

    @property
    def n_bases(self):
        return int(sum(len(r) for r in self.records))

    def joined(self, sep=b"\n"):
        """All records as one ASCII buffer, one separator byte between records (K1 ingest layout)."""
        return sep.join(r.tobytes() for r in self.records)

    def fasta_text(self, width=80):
        out = []
        for j, r in enumerate(self.records):
            out.append(f">{self.name} record {j}\n".encode())
            b = r.tobytes()
            for s in range(0, len(b), width):
                out.append(b[s:s + width] + b"\n")
        return b"".join(out)


def make_panel(n_in, n_out, genome_len, seed=1000, snp_every=1000, noise=1e-3,
               n_runs=10, run_len=100, dup_len=2000, soft_frac=0.01, soft_block=500,
               n_records=4, file_offset=0):
    """Return ``n_in`` ingroup + ``n_out`` outgroup :class:`Genome` objects.

    ``file_offset`` lets a rank of a multi-GPU job generate only its own slice
    of a larger panel: genome ``i`` always uses ``default_rng(seed + 1000 + i)``.
    """
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    sites = np.arange(snp_every // 2, genome_len, snp_every)
    two_allele = rng.random(len(sites)) < 0.10
    genomes = []
    for i in range(n_in + n_out):
        is_in = i < n_in
        name = (f"ingroup{i}" if is_in else f"outgroup{i - n_in}")
        genomes.append(_make_genome(anc, sites, two_allele, i + file_offset, is_in,
                                    (i % 2 == 1) and is_in, name, seed, noise, n_runs, run_len,
                                    dup_len, soft_frac, soft_block, n_records))
    return genomes


def make_genome(index, is_ingroup, odd_ingroup, name, genome_len, seed=1000, snp_every=1000, **kw):
    """One genome of the panel ``make_panel`` would build (for per-rank generation)."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    sites = np.arange(snp_every // 2, genome_len, snp_every)
    two_allele = rng.random(len(sites)) < 0.10
    p = dict(noise=1e-3, n_runs=10, run_len=100, dup_len=2000, soft_frac=0.01, soft_block=500, n_records=4)
    p.update(kw)
    return _make_genome(anc, sites, two_allele, index, is_ingroup, odd_ingroup, name, seed,
                        p["noise"], p["n_runs"], p["run_len"], p["dup_len"], p["soft_frac"],
                        p["soft_block"], p["n_records"])


def _make_genome(anc, sites, two_allele, index, is_in, odd_in, name, seed, noise, n_runs, run_len,
                 dup_len, soft_frac, soft_block, n_records):
    n = len(anc)
    g = anc.copy()
    if is_in:
        g[sites] = (anc[sites] + 1) & 3
        if odd_in:
            s2 = sites[two_allele]
            g[s2] = (anc[s2] + 2) & 3
    rng = np.random.default_rng(seed + 1000 + index)
    # private substitutions, iid over the genome
    n_sub = rng.binomial(n, noise)
    pos = rng.integers(0, n, size=n_sub)
    g[pos] = (g[pos] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3
    # one duplicated segment
    if dup_len and n > 4 * dup_len:
        a = int(rng.integers(0, n - dup_len))
        b = int(rng.integers(0, n - dup_len))
        g[b:b + dup_len] = g[a:a + dup_len].copy()
    seq = _ACGT[g]
    # runs of N
    for _ in range(n_runs):
        if n > 2 * run_len:
            s = int(rng.integers(0, n - run_len))
            seq[s:s + run_len] = ord("N")
    # soft-masked blocks
    n_blocks = int(round(soft_frac * n / soft_block)) if soft_block else 0
    for _ in range(n_blocks):
        if n > soft_block:
            s = int(rng.integers(0, n - soft_block))
            seq[s:s + soft_block] |= 0x20
    cuts = np.linspace(0, n, n_records + 1).astype(np.int64)
    recs = [seq[cuts[j]:cuts[j + 1]].copy() for j in range(n_records) if cuts[j + 1] > cuts[j]]
    return Genome(name=name, is_ingroup=is_in, records=recs)


def write_panel(genomes, directory, compress=False):
    """Write one FASTA per genome; returns (ingroup_paths, outgroup_paths)."""
    os.makedirs(directory, exist_ok=True)
    ins, outs = [], []
    for g in genomes:
        path = os.path.join(directory, g.name + (".fasta.gz" if compress else ".fasta"))
        data = g.fasta_text()
        if compress:
            with gzip.open(path, "wb", compresslevel=6) as fh:
                fh.write(data)
        else:
            with open(path, "wb") as fh:
                fh.write(data)
        (ins if g.is_ingroup else outs).append(path)
    return ins, outs
