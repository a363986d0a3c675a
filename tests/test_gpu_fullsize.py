"""The BASELINE panels at FULL size (config 2 and 3: 20+20 genomes x 5 Mbp, 4e8 records) against the C oracle's rows, pinned as
count + sha256 in tests/golden/fullsize.json (tests/golden/make_fullsize.py ran oracle/krisp_oracle.c on the same seeded panel)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN_DIR
from tests.test_gpu_parity import _search_panel

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN_DIR, "fullsize.json")) as _fh:
    _F = json.load(_fh)


@pytest.fixture(scope="module")
def panel():
    from krisp_b200.panel import make_panel
    return make_panel(20, 20, 5_000_000)


@pytest.fixture(scope="module")
def searcher():
    from krisp_b200.search import Searcher
    s = Searcher()
    yield s
    s.close()


@pytest.mark.parametrize("opts", [{}, {"sym": 1}, {"slab": 0}, {"bucket_bits": 18, "hash_slots_log2": 8, "hash_shared": 0}],
                         ids=["default", "window_items", "exact_path", "warp_tables"])
@pytest.mark.parametrize("which", sorted(_F))
def test_full_size_rows_equal_the_oracle(which, opts, panel, searcher):
    f = _F[which]
    if which != "c2" and opts:
        pytest.skip("plan variants are exercised on the one-word configuration")
    try:
        res = _search_panel(searcher, panel, f["L"], f["D"], f["R"], options=opts)
    finally:
        for k, v in (("slab", 1), ("sym", -1), ("bucket_bits", -1), ("hash_slots_log2", 0), ("hash_shared", -1)):
            searcher.set_option(k, v)
    rows = res.rows()
    assert res.n_records == f["records"]
    assert len(rows) == f["rows"]
    assert hashlib.sha256("\n".join(rows).encode()).hexdigest() == f["rows_sha256"]
    # the rows rendered and ordered on the device are the same set
    assert sorted(res.csv_rows_text().splitlines()) == rows


def test_full_size_from_host_buffers_in_batches(panel, searcher):
    """The e2e arm of bench.py: pinned host buffers, K1 + partition levels 0 and 1 per batch of arrived files."""
    import torch
    f = _F["c2"]
    pinned = []
    for g in panel:
        b = g.joined()
        t = torch.empty(len(b), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = np.frombuffer(b, dtype=np.uint8)
        pinned.append(t)
    searcher.configure(f["L"], f["D"], f["R"], [1 if g.is_ingroup else 0 for g in panel])
    searcher.clear_sequences()
    searcher.reserve(sum(t.numel() + 1 for t in pinned))
    for i, t in enumerate(pinned):
        searcher.add_sequence(i, t.numpy())
    rows = searcher.search(have_outgroup=True).rows()
    assert len(rows) == f["rows"] and hashlib.sha256("\n".join(rows).encode()).hexdigest() == f["rows_sha256"]
