"""Letters other than ACGTN (edge behaviour E1): what the unmodified reference does (tests/golden/iupac.json, made by
tests/golden/make_iupac.py), that the oracle restates it, and the documented contract of the CUDA path:

    krisp_b200(input) == reference(input with every letter outside ACGTNacgtn replaced by N)

i.e. such a letter breaks the k-mers that contain it.  Against the reference on the ORIGINAL input the row sets differ in exactly one
situation: an IUPAC letter in the diagnostic MIDDLE of an OUTGROUP occurrence — the reference keeps that k-mer (the letter counts as a
base of its own, disjoint from every ingroup base) and still prints the row, here the outgroup file lacks the flank key and the row
is lost.  A letter in a flank loses the row in both; a letter in an ingroup middle makes the reference's render worker raise KeyError
(exit 0, truncated output) while the rows of the untouched sites are still found here.
"""
import json
import os
import re

import pytest

from tests.helpers import GOLDEN_DIR

with open(os.path.join(GOLDEN_DIR, "iupac.json")) as _fh:
    _I = json.load(_fh)
_CASES = {c["name"]: c for c in _I["cases"]}


def _paths(c):
    return [os.path.join(GOLDEN_DIR, p) for p in c["ingroup"]], [os.path.join(GOLDEN_DIR, p) for p in c["outgroup"]]


def _with_n(paths, tmp_path):
    """Copies of the FASTA files with every sequence letter outside ACGTNacgtn replaced by N."""
    out = []
    for p in paths:
        lines = open(p).read().splitlines()
        fixed = [ln if ln.startswith(">") else re.sub(r"[^ACGTNacgtn]", "N", ln) for ln in lines]
        q = tmp_path / os.path.basename(p)
        q.write_text("\n".join(fixed) + "\n")
        out.append(str(q))
    return out


@pytest.mark.parametrize("name", sorted(_CASES))
def test_oracle_restates_the_reference_on_iupac_letters(name):
    from oracle import oracle
    c = _CASES[name]
    ins, outs = _paths(c)
    if c["reference_keyerror"]:
        with pytest.raises(oracle.OracleKeyError):
            oracle.search_files(ins, outs, c["L"], c["D"], c["R"])
    else:
        rows, _ = oracle.search_files(ins, outs, c["L"], c["D"], c["R"])
        assert rows == c["reference_rows"]


def test_reference_keeps_the_row_when_an_outgroup_middle_is_ambiguous():
    """The golden facts themselves: 40 rows without the letter; an outgroup-middle letter changes nothing, a flank letter costs the
    two rows (both strands) of that site, an ingroup-middle letter kills the reference's output."""
    assert len(_I["base"]["rows"]) == 40
    assert _CASES["out_mid"]["reference_rows"] == _I["base"]["rows"] == _CASES["out_mid_lower"]["reference_rows"]
    assert len(_CASES["flank"]["rows_lost_by_reference"]) == 2 and not _CASES["flank"]["rows_gained_by_reference"]
    assert _CASES["in_mid"]["reference_keyerror"] and _CASES["in_mid"]["reference_rows"] == []


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_CASES))
def test_cuda_path_treats_other_letters_like_n(name, tmp_path):
    from krisp_b200.search import search_files
    from oracle import oracle
    c = _CASES[name]
    ins, outs = _paths(c)
    got = search_files(ins, outs, c["L"], c["D"], c["R"]).rows()
    want, _ = oracle.search_files(_with_n(ins, tmp_path), _with_n(outs, tmp_path), c["L"], c["D"], c["R"])
    assert got == want                                              # the contract
    base = _I["base"]["rows"]
    assert set(got) <= set(base)
    if name.startswith("out_mid"):
        lost = sorted(set(c["reference_rows"]) - set(got))          # the documented divergence: the site's two rows (both strands)
        assert len(lost) == 2 and len(got) == 38
    elif name == "flank":
        assert got == c["reference_rows"]                           # no divergence: the reference loses the same rows
    else:
        assert len(got) == 38 and c["reference_rows"] == []         # the reference crashed; the untouched sites are still reported here
