"""kstream option combinations beyond krisp_fasta's own on the device: tables without complements, canonical k-mers, whole
k-mers without a split, ``--allow ACGT`` — against the unmodified reference's output (tests/golden/kstream_modes.json, written
by tests/golden/make_kstream_modes.py)."""
import hashlib
import json
import os

import pytest

from tests.helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN_DIR, "kstream_modes.json")) as _fh:
    _CASES = json.load(_fh)["cases"]


@pytest.mark.parametrize("case", _CASES, ids=[c["name"] for c in _CASES])
def test_kstream_modes_equal_reference(case, tmp_path):
    from krisp_b200 import kstream as ks
    out = tmp_path / "table.txt"
    assert ks.main([os.path.join(GOLDEN_DIR, case["file"])] + list(case["flags"]) + ["--output", str(out)]) == 0
    text = out.read_text()
    lines = text.splitlines()
    assert len(lines) == case["count"]
    assert lines[:3] == case["head"] and lines[-3:] == case["tail"]
    assert hashlib.sha256(text.encode()).hexdigest() == case["sha256"]


def test_kstream_write_counts_and_class_interface(tmp_path):
    """The class interface with the same options: write() returns the line count (kstream.py:250-325), iteration yields the lines."""
    from krisp_b200.kstream import kstream
    case = next(c for c in _CASES if c["name"] == "canonical_split_allow_28")
    path = os.path.join(GOLDEN_DIR, case["file"])
    k = kstream(path, kmers=28, canonicals=True, mapsoft=True, allow="ACGT", split=[25, -2], sort=True, sortcols=[0, 2])
    out = tmp_path / "t.txt"
    assert k.write(str(out)) == case["count"]
    assert hashlib.sha256(out.read_bytes()).hexdigest() == case["sha256"]
    assert list(k)[:3] == case["head"]


def test_search_ignores_the_strand_option_of_the_table_path():
    """"strands" belongs to kb_extract_sorted: a search on the same context still sees both strands."""
    from krisp_b200.search import Searcher, search_files
    from tests.helpers import golden_paths, load_golden
    g = load_golden()
    case = g["cases"][0]
    ins, outs = golden_paths(case)
    s = Searcher()
    try:
        s.set_option("strands", 2)
        from tests.helpers import deduce_ldr
        L, D, R = deduce_ldr(case["flags"])
        res = search_files(ins, outs, L, D, R, omit_soft=case["omit_soft"], searcher=s)
        assert res.rows() == case["rows"]
    finally:
        s.close()
