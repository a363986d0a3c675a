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
