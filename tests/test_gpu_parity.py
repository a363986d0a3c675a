"""GPU parity: the CUDA path (through the C ABI) against the golden vectors of the unmodified
reference and against the CPU oracle on seeded panels."""
import hashlib

import numpy as np
import pytest

from tests.helpers import deduce_ldr, golden_paths, load_golden

pytestmark = pytest.mark.gpu
_G = load_golden()


@pytest.fixture(scope="module")
def searcher():
    from krisp_b200.search import Searcher
    s = Searcher()
    yield s
    s.close()


@pytest.mark.parametrize("case", _G["cases"], ids=[c["name"] for c in _G["cases"]])
def test_rows_match_reference_golden(case, searcher):
    from krisp_b200.search import search_files
    ins, outs = golden_paths(case)
    L, D, R = deduce_ldr(case["flags"])
    res = search_files(ins, outs, L, D, R, omit_soft=case["omit_soft"], searcher=searcher)
    rows = res.rows()
    assert len(rows) == case["n_rows"]
    assert hashlib.sha256("\n".join(rows).encode()).hexdigest() == case["rows_sha256"]
    if "rows" in case:
        assert rows == case["rows"]


@pytest.mark.parametrize("sort_bits", [8, 16, 24, 64])
@pytest.mark.parametrize("name", ["c1_spacer_25_1_2", "p_spacer_3x3", "p_primer_3x3", "p_5_2_3", "p_30_40_30"])
def test_prefix_sort_is_exact(name, sort_bits, searcher):
    """Sorting on a short prefix forces runs that mix several flank keys: the group pass must still be exact."""
    from krisp_b200.search import search_files
    case = next(c for c in _G["cases"] if c["name"] == name)
    ins, outs = golden_paths(case)
    L, D, R = deduce_ldr(case["flags"])
    res = search_files(ins, outs, L, D, R, omit_soft=case["omit_soft"], searcher=searcher,
                       options={"sort_bits": sort_bits})
    searcher.set_option("sort_bits", 32)
    assert res.rows() == case["rows"]
    if sort_bits == 8:
        assert res.stats["mixed_runs"] > 0


@pytest.mark.parametrize("name", ["c1_spacer_25_1_2", "p_spacer_4x5", "p_6_1_2", "p_0_2_6", "p_8_0_8", "c1_no_outgroup", "c1_single_file"])
def test_generic_group_kernel_matches_too(name, searcher):
    """The warp-per-run kernel (used for multi-word records, > 64 files, D > 8) on shapes the fast path normally takes."""
    import hashlib as _h
    from krisp_b200.search import search_files
    case = next(c for c in _G["cases"] if c["name"] == name)
    ins, outs = golden_paths(case)
    L, D, R = deduce_ldr(case["flags"])
    try:
        res = search_files(ins, outs, L, D, R, omit_soft=case["omit_soft"], searcher=searcher, options={"fast_group": 0})
    finally:
        searcher.set_option("fast_group", 1)
    rows = res.rows()
    assert len(rows) == case["n_rows"]
    assert _h.sha256("\n".join(rows).encode()).hexdigest() == case["rows_sha256"]


_TABLES = _G["tables"]


@pytest.mark.parametrize("tab", _TABLES, ids=[f'{t["file"]}:{t["L"]}_{t["D"]}_{t["R"]}{"_omit" if t["omit_soft"] else ""}' for t in _TABLES])
def test_kstream_sorted_table_matches_reference(tab):
    """kstream(...).write == the reference's `*.{k}mers` file (count = write()'s return value, sha256 of the text)."""
    import os
    import tempfile
    from krisp_b200._lib import UnsupportedError
    from krisp_b200.kstream import kstream
    from tests.helpers import GOLDEN_DIR
    L, D, R = tab["L"], tab["D"], tab["R"]
    ks = kstream(kmers=L + D + R, complements=True, disallow="Nn", omitsoft=tab["omit_soft"], mapsoft=not tab["omit_soft"],
                 split=[L, -R], sort=True, sortcols=[0, 2])
    path = os.path.join(GOLDEN_DIR, tab["file"])
    if 2 * (L + D + R) + 8 > 64:
        with pytest.raises(UnsupportedError):
            list(ks(path))
        return
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "t.kmers")
        n = ks.write(out, path)
        text = open(out).read()
    lines = text.splitlines()
    assert n == tab["count"] == len(lines)
    assert lines[:5] == tab["head"] and lines[-5:] == tab["tail"]
    assert hashlib.sha256(text.encode()).hexdigest() == tab["sha256"]


_ALIGN = [c for c in _G["cases"] if "out_align" in c]


@pytest.mark.parametrize("case", _ALIGN, ids=[c["name"] for c in _ALIGN])
def test_cli_csv_and_out_align_match_reference(case, capsys, tmp_path):
    """The krisp_fasta command line end to end: stdout CSV (header + rows) and the --out_align file."""
    from krisp_b200 import krisp_fasta
    ins, outs = golden_paths(case)
    argv = list(ins)
    if outs:
        argv += ["--outgroup"] + list(outs)
    for k, v in case["flags"].items():
        argv += [f"--{k}", str(v)]
    if case["omit_soft"]:
        argv.append("--omit-soft")
    ap = tmp_path / "align.txt"
    argv += ["--out_align", str(ap)]
    if case["dot"]:
        argv.append("--dot-alignment")
    assert krisp_fasta.main(argv) == 0
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "left_seq,diag_seq,right_seq"
    assert sorted(out[1:]) == case["rows"]
    assert ap.read_text() == case["out_align"]
