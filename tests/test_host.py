"""Host-side logic that needs no GPU: ingest, labels, flag deduction, decoding, rendering, shard maths."""
import argparse
import glob
import os
import re

import numpy as np
import pytest

from krisp_b200 import ingest, names, render, sharded
from krisp_b200.search import SearchResult, _decode_bases, _decode_masks, labels_for
from oracle import model
from tests.helpers import GOLDEN_DIR, deduce_ldr, load_golden

_G = load_golden()
_FILES = sorted(set(glob.glob(os.path.join(GOLDEN_DIR, "**", "*.fa*"), recursive=True) +
                    glob.glob(os.path.join(GOLDEN_DIR, "**", "*.fna*"), recursive=True)))


@pytest.mark.parametrize("path", _FILES, ids=[os.path.relpath(p, GOLDEN_DIR) for p in _FILES])
def test_ingest_matches_reference_parser(path):
    """pack_bytes == the reference's _parse_FASTA records (restated in oracle/model.py), joined by separators."""
    arr, rna = ingest.load_file(path)
    recs = [r for r in model.fasta_records(model.read_lines(path)) if r]
    got = re.sub(rb"\n+", b"\n", arr.tobytes().strip(b"\n"))
    assert got == "\n".join(recs).encode()
    assert rna == model.detect_rna(recs)


def test_ingest_edge_cases():
    assert ingest.pack_bytes(b"").size == 0
    assert ingest.pack_bytes(b">only header").size == 0
    # headerless input: the first line is consumed by the FASTA probe (kstream.py:450), the rest are records
    assert ingest.pack_bytes(b"ACGT\nGGCC\nTTAA\n").tobytes() == b"GGCC\nTTAA"
    # CRLF and padding are stripped per line (kstream.py:574)
    assert re.sub(rb"\n+", b"\n", ingest.pack_bytes(b">a\r\nAC GT \r\n  TT\r\n>b\r\n>c\r\nGG\r\n").tobytes().strip(b"\n")) == b"AC GTTT\nGG"
    # '>' inside a sequence line does not start a record
    assert re.sub(rb"\n+", b"\n", ingest.pack_bytes(b">a\nAC>GT\nTT\n").tobytes().strip(b"\n")) == b"AC>GTTT"
    assert ingest.detect_rna(np.frombuffer(b"ACGU\nACGT", dtype=np.uint8))
    assert not ingest.detect_rna(np.frombuffer(b"ACGUT\nACG", dtype=np.uint8))
    assert not ingest.detect_rna(np.frombuffer(b"ACGG", dtype=np.uint8))


@pytest.mark.parametrize("name", ["a/b/GCF_1.1_x.fna.gz", "x.fasta", "x.v2.fa.bz2", "plain", "dir/ingroup0.fasta.gz", "q.frn", "q.ffn.gz"])
def test_names_match_reference(name):
    assert names.basename(name) == model.basename(name)
    assert names.simplename(name) == model.simplename(name)


def test_labels_single_file_and_collisions():
    labs, isin = labels_for(["only.fa"], [])
    assert labs == ["merged_file"] and isin == [0]
    labs, isin = labels_for(["alpha.fasta"], ["alpha.v2.fasta", "gamma.1.fa"])
    assert labs == ["alpha", "alpha", "gamma"] and isin == [1, 1, 0]


@pytest.mark.parametrize("case", _G["cases"], ids=[c["name"] for c in _G["cases"]])
def test_cli_deduction_matches_reference_rule(case):
    from krisp_b200 import krisp_fasta
    argv = ["x.fa"]
    for k, v in case["flags"].items():
        argv += [f"--{k}", str(v)]
    args = krisp_fasta.deduce(krisp_fasta.build_parser().parse_args(argv))
    L, D, R = deduce_ldr(case["flags"])
    assert (args.conserved_left, args.amplicon - args.conserved_left - args.conserved_right, args.conserved_right) == (L, D, R)


def test_cli_exits_when_underspecified(capsys):
    from krisp_b200 import krisp_fasta
    with pytest.raises(SystemExit) as e:
        krisp_fasta.deduce(krisp_fasta.build_parser().parse_args(["x.fa", "--conserved", "5"]))
    assert e.value.code == 1


def _pack(seq):
    v = 0
    for ch in seq:
        v = (v << 2) | "ACGT".index(ch)
    return v


def test_decoders():
    left, right = "ACGTTGCA", "GT"
    bits = 2 * (len(left) + len(right))
    w = np.array([[_pack(left + right) << (64 - bits)]], dtype=np.uint64)
    assert _decode_bases(w, 0, 8).tobytes().decode() == left
    assert _decode_bases(w, 16, 2).tobytes().decode() == right
    # masks: column c = nibble 7 - c%8 of word c/8
    m = np.array([[0x1248F000, 0x30000000]], dtype=np.uint32)
    assert _decode_masks(m, 9).tolist() == [[1, 2, 4, 8, 15, 0, 0, 0, 3]]


def test_rows_and_iupac():
    res = SearchResult(L=2, D=3, R=1, have_outgroup=True)
    res.left = np.frombuffer(b"ACGT", dtype=np.uint8).reshape(2, 2)
    res.right = np.frombuffer(b"GT", dtype=np.uint8).reshape(2, 1)
    res.in_mask = np.array([[1, 3, 15], [8, 12, 2]], dtype=np.uint8)
    res.out_mask = np.array([[2, 4, 1], [1, 1, 1]], dtype=np.uint8)
    assert res.rows() == ["AC,AMN,G", "GT,TKC,T"]
    res.have_outgroup = False
    assert res.rows() == ["AC,MVN,G", "GT,WDM,T"]
    assert render.csv_rows(res) == ["AC,MVN,G", "GT,WDM,T"]


def test_render_alignment_matches_reference_text():
    """Alignment blocks of the golden cases, rebuilt from the oracle's groups, equal the reference's --out_align text."""
    for case in _G["cases"]:
        if "out_align" not in case:
            continue
        ins = [os.path.join(GOLDEN_DIR, p) for p in case["ingroup"]]
        outs = [os.path.join(GOLDEN_DIR, p) for p in case["outgroup"]]
        L, D, R = deduce_ldr(case["flags"])
        groups, ingroup = model.search_groups(ins, outs, L, D, R, case["omit_soft"])
        ing = frozenset(ingroup) if outs else None
        blocks = []
        for (left, right) in sorted(groups):
            amps = {seq[1]: labs for seq, labs in groups[(left, right)].items()}
            blocks.append(render.render_alignment(left, right, amps, ing, case["dot"]) + "\n")
        assert "".join(blocks) == case["out_align"], case["name"]


def test_assign_files_balances():
    assert sharded.assign_files(5, 2) == [0, 1, 0, 1, 0]
    owner = sharded.assign_files(4, 2, sizes=[10, 1, 1, 8])
    loads = [sum(s for s, o in zip([10, 1, 1, 8], owner) if o == r) for r in range(2)]
    assert sorted(loads) == [10, 10]


def test_digit_ranges_partition_the_key_space():
    """Every level-0 digit has exactly one owner shard, and owners are contiguous digit ranges."""
    keys = np.arange(0, 1 << 18, 7, dtype=np.uint64)
    digits = sharded.digit_of_key(keys, 18, 5)
    assert digits.min() == 0 and digits.max() == 31
    for n in (1, 2, 3, 8):
        firsts = [sharded.first_digit(s, n, 32) for s in range(n + 1)]
        assert firsts[0] == 0 and firsts[-1] == 32 and all(a < b for a, b in zip(firsts, firsts[1:]))


def test_piece_tables_tile_every_receive_buffer_exactly():
    """Fused partition + exchange: the pieces (source rank, digit) every rank writes into an owner's buffer must tile that buffer
    without gaps or overlaps, in the order the owner's kb_shard_search expects (source-major, digit-minor)."""
    rng = np.random.default_rng(7)
    for world, nd in ((2, 8), (3, 16), (8, 32), (4, 256)):
        table = rng.integers(0, 50, size=(world, nd))
        table[rng.integers(0, world), rng.integers(0, nd)] = 0
        per_rank = [sharded.piece_tables(table, r) for r in range(world)]
        firsts = [sharded.first_digit(s, world, nd) for s in range(world + 1)]
        for s in range(world):
            need = per_rank[0][0][s]
            cover = np.zeros(need, dtype=np.int64)
            for src in range(world):
                base = per_rank[src][1]
                for d in range(firsts[s], firsts[s + 1]):
                    cover[base[d]:base[d] + table[src, d]] += 1
            assert np.all(cover == 1)
            # arrival order at the owner: pieces listed by source, then digit, with these very counts
            pieces = per_rank[s][2]
            assert pieces == table[:, firsts[s]:firsts[s + 1]].reshape(-1).tolist() and sum(pieces) == need
            pos = 0
            for i, c in enumerate(pieces):
                src, j = divmod(i, firsts[s + 1] - firsts[s])
                assert per_rank[src][1][firsts[s] + j] == pos
                pos += c


def _pack_words(bits_value, nbits, words):
    """MSB-first bit string of `nbits` bits -> `words` uint64 words (zero padded at the end)."""
    v = bits_value << (64 * words - nbits)
    return [(v >> (64 * (words - 1 - j))) & 0xFFFFFFFFFFFFFFFF for j in range(words)]


def test_interchange_lines_round_trip_golden_files():
    """render.interchange_lines on results assembled from the reference's own interchange files (tests/golden/interchange.json)
    reproduces them: record layout [left | right | mid | pad | file id], label grammar name / name(n), group order."""
    import json
    import re
    from krisp_b200.search import labels_for
    with open(os.path.join(GOLDEN_DIR, "interchange.json")) as fh:
        inter = json.load(fh)
    assert len(inter) >= 10
    for name, text in inter.items():
        case = next(c for c in _G["cases"] if c["name"] == name)
        L, D, R = deduce_ldr(case["flags"])
        ins = [os.path.join(GOLDEN_DIR, p) for p in case["ingroup"]]
        outs = [os.path.join(GOLDEN_DIR, p) for p in case["outgroup"]]
        labels, _ = labels_for(ins, outs)
        k = L + D + R
        W = 1 if 2 * k + 8 <= 64 else (2 if 2 * k + 8 <= 128 else (4 if 2 * k + 8 <= 256 else 8))
        FW = max(1, (2 * (L + R) + 63) // 64)
        groups = []
        for ln in text.splitlines():
            l, m, r, labs = ln.split(",")
            if not groups or groups[-1][0] != (l, r):
                groups.append(((l, r), []))
            for item in labs.split(";"):
                mt = re.fullmatch(r"(.+?)(?:\((\d+)\))?", item)
                groups[-1][1].extend([(m, mt.group(1))] * int(mt.group(2) or 1))
        res = SearchResult(L=L, D=D, R=R, have_outgroup=bool(outs))
        flank, recs, off = [], [], [0]
        for (l, r), occ in groups:
            flank.append(_pack_words(_pack(l + r), 2 * (L + R), FW) if L + R else [0] * FW)
            for m, lab in occ:
                words = _pack_words(_pack(l + r + m), 2 * k, W)
                words[-1] |= labels.index(lab)
                recs.append(words)
            off.append(len(recs))
        res.flank_words = np.array(flank, dtype=np.uint64).reshape(len(groups), FW)
        res.records = np.array(recs, dtype=np.uint64).reshape(len(recs), W)
        res.run_offset = np.array(off, dtype=np.uint64)
        got = render.interchange_lines(res, labels)

        def grouped(lines):
            out = []
            for ln in lines:
                l, _m, r, _x = ln.split(",")
                if out and out[-1][0] == (l, r):
                    out[-1][1].append(ln)
                else:
                    out.append(((l, r), [ln]))
            return [(kk, sorted(v)) for kk, v in out]
        assert grouped(got) == grouped(text.splitlines()), name


def test_shard_child_counts_are_the_owned_slice_summed_over_sources():
    """sharded.shard_child_counts: K1's two-level histograms of all ranks -> level-1 child counts of one shard's digits (what the
    owner would count over the records it receives)."""
    from krisp_b200 import sharded
    rng = np.random.default_rng(5)
    for world, bits0, bits1 in [(2, 3, 4), (4, 4, 7), (8, 5, 7), (3, 3, 2)]:
        nd, per = 1 << bits0, 1 << bits1
        # records as (source, digit, child) triples
        src = rng.integers(0, world, 5000)
        key = rng.integers(0, nd * per, 5000)
        children = np.zeros((world, nd * per), dtype=np.int64)
        np.add.at(children, (src, key), 1)
        covered = 0
        for rank in range(world):
            got = sharded.shard_child_counts(children, nd, rank)
            lo, hi = sharded.first_digit(rank, world, nd), sharded.first_digit(rank + 1, world, nd)
            mine = key[(key >> bits1 >= lo) & (key >> bits1 < hi)]            # every source's records of the digits this rank owns
            want = np.bincount(mine - lo * per, minlength=(hi - lo) * per)
            assert got.dtype == np.uint64 and got.tolist() == want.tolist()
            covered += int(got.sum())
        assert covered == 5000


def test_decode_table_handles_one_word_and_multi_word_records():
    """kstream.decode_table on hand-packed records ([left | right | mid | pad | id], 1 / 2 / 4 / 8 words) == the C oracle's sorted
    table lines of the same sequence, including the R == 0 quirk (middle lands in the third field)."""
    from krisp_b200.kstream import decode_table
    from oracle import oracle
    rng = np.random.default_rng(11)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, 600))
    for L, D, R in [(25, 1, 2), (9, 1, 0), (20, 10, 17), (32, 60, 32), (100, 20, 100), (0, 30, 10)]:
        text, n = oracle.table_text([seq], L, D, R)
        want = text.splitlines()
        k = L + D + R
        W = 1 if 2 * k + 8 <= 64 else (2 if 2 * k + 8 <= 128 else (4 if 2 * k + 8 <= 256 else 8))
        recs = []
        for ln in want:
            f = ln.split(",")
            left, mid, right = (f[0], f[2], "") if R == 0 else (f[0], f[1], f[2])      # quirk S9: `left,,mid`
            recs.append(_pack_words(_pack(left + right + mid), 2 * k, W))
        arr = np.array(recs, dtype=np.uint64).reshape(len(recs), W)
        got = decode_table(arr if W > 1 else arr[:, 0], L, D, R)
        assert got == want and len(got) == n, (L, D, R)


def test_primer3_post_filter_wiring(tmp_path, monkeypatch):
    """--primer3 (render.render_output find_primers=True): every region's template left + consensus + right goes to
    primer3.bindings.design_primers with the diagnostic region as target and the reference's settings (Amplicon.py:103-151), regions
    without a primer pair are dropped, the CSV gains the 20 primer columns (outputAlignments.py:10-31) and the alignment the primer
    annotation and statistics (Amplicon.py:640-660).  primer3-py is not in this image: a stand-in module records the calls."""
    import sys
    import types
    calls = []

    def design_primers(seq_args, global_args):
        calls.append((seq_args, global_args))
        if seq_args["SEQUENCE_TEMPLATE"].startswith("GT"):
            return {"PRIMER_PAIR_NUM_RETURNED": 0}
        out = {"PRIMER_PAIR_NUM_RETURNED": 1, "PRIMER_LEFT_0": (0, 4), "PRIMER_RIGHT_0": (11, 4)}
        for i, n in enumerate(render.P3_COLS):
            out[n] = "ACGA" if n == "PRIMER_LEFT_0_SEQUENCE" else ("TTCA" if n == "PRIMER_RIGHT_0_SEQUENCE" else 1.5 + i)
        return out

    fake = types.ModuleType("primer3")
    fake.bindings = types.SimpleNamespace(design_primers=design_primers)
    monkeypatch.setitem(sys.modules, "primer3", fake)

    res = SearchResult(L=4, D=3, R=5, have_outgroup=True)
    res.left = np.frombuffer(b"ACGAGTTT", dtype=np.uint8).reshape(2, 4)
    res.right = np.frombuffer(b"GTGAATCCCC", dtype=np.uint8).reshape(2, 5)
    res.in_mask = np.array([[1, 3, 8], [8, 4, 2]], dtype=np.uint8)
    res.out_mask = np.array([[2, 4, 1], [1, 1, 1]], dtype=np.uint8)
    res.rows_blob = None                                   # rows from the host decoder
    csv = tmp_path / "out.csv"
    n = render.render_output(res, ["a", "b"], ingroup=["a"], out_csv=str(csv), find_primers=True,
                             p3_args={"tm": [53, 68], "gc": [40, 70], "amp_size": [70, 150], "primer_size": [25, 35], "max_sec_tm": 40, "gc_clamp": 1, "max_end_gc": 4})
    assert n == 1
    lines = csv.read_text().splitlines()
    assert lines[0] == "left_seq,diag_seq,right_seq," + ",".join(render.P3_KEYS)
    assert lines[0].split(",")[3:6] == ["pair_product_size", "pair_penalty", "left_sequence"] and len(lines[0].split(",")) == 23
    assert len(lines) == 2 and lines[1].startswith("ACGA,AMT,GTGAA,1.5,2.5,ACGA,TTCA,")
    assert [c[0]["SEQUENCE_TEMPLATE"] for c in calls] == ["ACGAAMTGTGAA", "GTTTTGCTCCCC"]
    assert all(c[0]["SEQUENCE_TARGET"] == [4, 3] for c in calls)
    g = calls[0][1]
    assert g["PRIMER_PRODUCT_SIZE_RANGE"] == [[70, 150]] and g["PRIMER_OPT_SIZE"] == 30 and g["PRIMER_OPT_TM"] == 60.5
    assert g["PRIMER_MIN_GC"] == 40 and g["PRIMER_MAX_HAIRPIN_TH"] == 40 and g["PRIMER_TASK"] == "generic" and len(g) == 23
    # annotation: merged into the bracket line, statistics underneath
    block = render.render_alignment("ACGA", "GTGAA", {"AAT": ["a"], "CCA": ["b"]}, frozenset(["a"]), False)
    p3 = design_primers({"SEQUENCE_TEMPLATE": "ACGAAMTGTGAA"}, {})
    text = render.annotate_alignment(block, p3, False).split("\n")
    assert text[0].startswith("ACGAAATGTGAA : a") and text[1].startswith("ACGACCAGTGAA : b")
    assert text[2] == "└Fo{###}┘    └Reverse┘"      # annotation from the primer positions, the bracket's own characters win
    assert text[3] == "" and text[4] == "Primer statistics:" and "Direction" in text[5] and text[6].lstrip().startswith("Forward")
    assert "Pair statistics:" in text
    dot = render.annotate_alignment(render.render_alignment("ACGA", "GTGAA", {"AAT": ["a"], "CCA": ["b"]}, frozenset(["a"]), True), p3, True).split("\n")
    assert dot[1].startswith("....CCA..... : b") and dot[2] == "└Forward┘    └Reverse┘"


def test_render_output_file_reproduces_reference_rows_and_alignments(tmp_path):
    """The renderer stage on a FILE (render.render_output_file = render_output(kmerfile, ...), outputAlignments.py:101-162): fed the
    reference's own filtered.txt / merged_file.txt (tests/golden/interchange.json) it writes the reference's CSV rows and, where
    the golden case has it, the reference's --out_align text."""
    import json
    from krisp_b200.names import simplename
    with open(os.path.join(GOLDEN_DIR, "interchange.json")) as fh:
        inter = json.load(fh)
    checked_align = 0
    for name, text in inter.items():
        case = next(c for c in _G["cases"] if c["name"] == name)
        if "rows" not in case:
            continue
        src = tmp_path / f"{name}.txt"
        src.write_text(text)
        ingroup = [simplename(p) for p in case["ingroup"]] if case["outgroup"] else None
        csv, align = tmp_path / f"{name}.csv", tmp_path / f"{name}.align"
        n = render.render_output_file(str(src), out_align=str(align) if "out_align" in case else None, out_csv=str(csv), ingroup=ingroup, dot=case["dot"])
        lines = csv.read_text().splitlines()
        assert lines[0] == render.CSV_HEADER and sorted(lines[1:]) == case["rows"] and n == len(case["rows"]), name
        if "out_align" in case:
            assert align.read_text() == case["out_align"], name
            checked_align += 1
    assert checked_align >= 1


def test_kstream_constructor_accepts_device_tables_and_rejects_the_text_pipeline():
    """krisp_b200.kstream maps the reference's constructor arguments (kstream.py:122-248) onto the sorted tables the device builds
    and raises UnsupportedError — never a silent CPU fallback — for what stays the reference's generic text pipeline."""
    from krisp_b200._lib import UnsupportedError
    from krisp_b200.kstream import kstream
    ok = [(dict(kmers=28, complements=True, disallow="Nn", mapsoft=True, split=[25, -2], sort=True, sortcols=[0, 2]), (25, 1, 2, 0)),
          (dict(kmers=28, disallow="Nn", mapsoft=True, sort=True), (28, 0, 0, 1)),
          (dict(kmers=28, canonicals=True, allow="ACGT", mapsoft=True, split=[25, -2], sort=True, sortcols=[0, 2]), (25, 1, 2, 2)),
          (dict(kmers=15, canonicals=True, allow="ACGTacgt", omitsoft=True, sort=True), (15, 0, 0, 2)),
          (dict(kmers=124, complements=True, disallow="Nn", omitsoft=True, split=[32, -32], sort=True, sortcols=[0, 2]), (32, 60, 32, 0))]
    for kw, want in ok:
        k = kstream(**kw)
        assert (k.L, k.D, k.R, k.strands) == want
    bad = [dict(kmers=28, disallow="Nn", mapsoft=True),                                   # unsorted streaming
           dict(kmers=[5, 6], disallow="Nn", mapsoft=True, sort=True),                     # several k
           dict(kmers=28, allow="ACGTN", mapsoft=True, sort=True),                         # letters the 2-bit records cannot hold
           dict(kmers=28, disallow="NnA", mapsoft=True, sort=True),
           dict(kmers=28, disallow="Nn", sort=True),                                       # lower case kept as such
           dict(kmers=28, disallow="Nn", mapsoft=True, expandiupac=True, sort=True),
           dict(kmers=40, canonicals=True, disallow="Nn", mapsoft=True, sort=True),        # strand modes beyond one-word records
           dict(kmers=28, disallow="Nn", mapsoft=True, split=[25, -2], sort=True),         # split without its sort columns
           dict(kmers=28, disallow="Nn", mapsoft=True, split=[5, 5, 5], sort=True, sortcols=[0, 2])]
    for kw in bad:
        with pytest.raises(UnsupportedError):
            kstream(**kw)
    with pytest.raises(ValueError):
        kstream(kmers=28, complements=True, canonicals=True)
    with pytest.raises(ValueError):
        kstream(kmers=28, omitsoft=True, mapsoft=True)


def test_bit_parallel_fasta_walk_equals_byte_walk(tmp_path):
    """The de-lining kernels' per-thread logic (csrc/kb_ingest.cuh: SWAR byte classes, prefix propagation of the header state, keep /
    separator masks) against its byte-at-a-time form on 4 M random 16-byte chunks — host code compiled from the same header
    (tools/fa_walk_check.cu), no GPU needed."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(GOLDEN_DIR.rstrip("/")).rsplit("/tests", 1)[0]
    exe = str(tmp_path / "fa_walk_check")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(root, "krisp_b200", "csrc"),
                    "-o", exe, os.path.join(root, "tools", "fa_walk_check.cu")], check=True, capture_output=True, env=env)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "0 mismatches" in p.stdout, p.stdout[-2000:]
