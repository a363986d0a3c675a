"""GPU parity of the slab search path (K1 fused with partition level 0, kb_extract_part.cuh; warp-private bucket hash,
kb_hash_warp.cuh) against the C oracle, the golden vectors and the exact (histogram-based) path."""
import hashlib

import numpy as np
import pytest

from tests.helpers import deduce_ldr, golden_paths, load_golden
from tests.test_gpu_parity import _oracle_panel, _search_panel

pytestmark = pytest.mark.gpu
_G = load_golden()

_RESET = {"slab": 1, "sym": -1, "hash_warp": 1, "hash_shared": -1, "bucket_bits": -1, "hash_slots_log2": 0, "slab_cap": 0, "want_records": 0, "profile": 0,
          "hash_warps": 0, "rank_rows": 1, "group_sizes": 0}


@pytest.fixture(scope="module")
def searcher():
    from krisp_b200.search import Searcher
    s = Searcher()
    yield s
    s.close()


def _reset(searcher):
    for k, v in _RESET.items():
        searcher.set_option(k, v)


def _stages(res):
    return [nm for nm, _ in res.profile]


def _slab(res):
    """The search ran K1 fused with partition level 0 (the slab path)."""
    return any(nm.startswith("K1 extract + partition 0") for nm in _stages(res))


@pytest.mark.parametrize("mode", [(1, 1, 0, 1), (1, 1, 1, 1), (1, 1, 1, 0), (0, 1, 0, 1), (0, 1, 1, 1), (0, 0, 0, 1)],
                         ids=["slab+warp", "slab+cta", "slab_records+cta", "exact+warp", "exact+cta", "exact+stream"])
@pytest.mark.parametrize("shape", [(6, 6, 300_000, 25, 1, 2, False), (6, 6, 300_000, 25, 1, 2, True), (3, 3, 400_000, 12, 3, 12, False),
                                   (20, 20, 100_000, 25, 1, 2, False), (40, 40, 40_000, 10, 4, 10, False), (50, 50, 40_000, 25, 1, 2, False),
                                   (70, 70, 20_000, 10, 4, 10, False), (2, 1, 40_000, 20, 0, 7, False), (3, 2, 300_000, 9, 8, 9, False)],
                         ids=["spacer", "spacer_omit", "12_3_12", "bench_shape_small", "80_files", "100_files_spacer", "140_files", "D0", "D8"])
def test_slab_and_exact_paths_match_oracle(shape, mode, searcher):
    """Same rows, group sizes and gathered records whichever way the records were grouped."""
    from krisp_b200.panel import make_panel
    n_in, n_out, glen, L, D, R, omit = shape
    kw = dict(n_runs=1, run_len=30, noise=2e-4 if n_in + n_out <= 80 else 1e-4, dup_len=300, soft_block=100) if n_in + n_out > 64 else {}
    gs = make_panel(n_in, n_out, glen, **kw)
    try:
        res = _search_panel(searcher, gs, L, D, R, omit, options={"slab": mode[0], "hash_warp": mode[1], "hash_shared": mode[2], "sym": mode[3], "profile": 1}, want_records=True)
    finally:
        _reset(searcher)
    want = _oracle_panel(gs, L, D, R, omit)
    assert len(want) > 0
    assert res.rows() == want
    assert _slab(res) == bool(mode[0])
    # every group's gathered records carry its flank key, and there are group_size of them
    FB = 2 * (L + R)
    for g in range(res.n_groups):
        a, b = int(res.run_offset[g]), int(res.run_offset[g + 1])
        recs = res.records[a:b, 0]
        fw = res.flank_words[g, 0]
        mine = recs[(recs >> np.uint64(64 - FB)) == (fw >> np.uint64(64 - FB))] if FB else recs
        assert int(res.group_size[g]) == mine.size >= n_in + n_out


@pytest.mark.parametrize("shared", [0, 1, 2], ids=["warp", "cta", "cta_records"])
@pytest.mark.parametrize("bucket_bits,slots", [(1, 0), (3, 5), (9, 0), (10, 6), (13, 4), (17, 0), (18, 7), (20, 0), (2, 11), (4, 12)])
@pytest.mark.parametrize("name", ["c1_spacer_25_1_2", "p_spacer_3x3", "p_5_2_3", "p_6_1_2", "c1_single_file", "p_spacer_4x5"])
def test_slab_path_on_golden_cases_for_any_depth(name, bucket_bits, slots, shared, searcher):
    """One to three slab levels and tiny hash tables (deferred buckets, splits) on the reference's golden cases."""
    from krisp_b200.search import search_files
    case = next(c for c in _G["cases"] if c["name"] == name)
    ins, outs = golden_paths(case)
    L, D, R = deduce_ldr(case["flags"])
    try:
        res = search_files(ins, outs, L, D, R, omit_soft=case["omit_soft"], searcher=searcher,
                           options={"bucket_bits": bucket_bits, "hash_slots_log2": slots, "hash_shared": min(shared, 1), "sym": 0 if shared == 2 else 1,
                                    "profile": 1})
    finally:
        _reset(searcher)
    rows = res.rows()
    assert _slab(res)
    assert len(rows) == case["n_rows"]
    assert hashlib.sha256("\n".join(rows).encode()).hexdigest() == case["rows_sha256"]


def test_slab_overflow_falls_back_to_the_exact_path(searcher):
    """Slabs of 2 records overflow at once: the search must notice, repeat itself on the exact path, stay there for these
    sequences, and try slabs again once the sequences change."""
    from krisp_b200.panel import make_panel
    gs = make_panel(4, 4, 200_000)
    want = _oracle_panel(gs, 25, 1, 2)
    try:
        res = _search_panel(searcher, gs, 25, 1, 2, options={"slab_cap": 2, "profile": 1})
        assert res.rows() == want
        assert "K1 extract" in _stages(res) and not _slab(res)
        res2 = searcher.search(have_outgroup=True)
        assert res2.rows() == want and not _slab(res2)
        searcher.set_option("slab_cap", 0)
        res3 = _search_panel(searcher, gs, 25, 1, 2, options={"profile": 1})
        assert res3.rows() == want and _slab(res3)
    finally:
        _reset(searcher)


@pytest.mark.parametrize("warps", [8, 10])
@pytest.mark.parametrize("rank_rows", [0, 1], ids=["radix_rows", "rank_rows"])
@pytest.mark.parametrize("shape", [(6, 6, 300_000, 25, 1, 2), (3, 3, 400_000, 12, 3, 12), (50, 50, 40_000, 25, 1, 2), (3, 2, 200_000, 40, 6, 50)],
                         ids=["spacer", "12_3_12", "100_files", "multiword"])
def test_result_tail_variants_agree(shape, rank_rows, warps, searcher):
    """The shared-table bucket hash with 8 or 10 warps per CTA, rows ordered by counting ranks or by the chunked radix sort, group
    sizes asked for or not: same rows in the same (ascending) order as the host decoder, sizes only where asked for."""
    from krisp_b200.panel import make_panel
    n_in, n_out, glen, L, D, R = shape
    kw = dict(n_runs=1, run_len=30, noise=1e-4, dup_len=300, soft_block=100) if n_in + n_out > 64 else {}
    gs = make_panel(n_in, n_out, glen, **kw)
    want = _oracle_panel(gs, L, D, R)
    try:
        res = _search_panel(searcher, gs, L, D, R, options={"hash_warps": warps, "hash_shared": 1, "rank_rows": rank_rows, "group_sizes": rank_rows})
    finally:
        _reset(searcher)
    assert len(want) > 0 and res.rows() == want
    text = res.csv_rows_text().splitlines()
    assert sorted(text) == want
    from krisp_b200.render import csv_rows
    assert text == csv_rows(res)                                  # device order == ascending (left, right)
    if rank_rows or 2 * (L + D + R) + 8 > 64:                     # (multi-word records: the generic hash kernel counts as it goes)
        assert res.group_size is not None and int(res.group_size.min()) >= n_in + n_out
    else:
        assert res.group_size is None


def test_repeats_overflow_one_slab_and_still_give_the_oracle_rows(searcher):
    """A genome family full of one repeated unit: thousands of records share a few keys, far beyond a slab's slack."""
    from krisp_b200.panel import Genome
    rng = np.random.default_rng(7)
    unit = rng.integers(0, 4, size=40, dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    gs = []
    for i in range(6):
        body = rng.integers(0, 4, size=150_000, dtype=np.uint8)
        body[20_000:20_000 + 40 * 1500] = np.tile(unit, 1500)
        if i < 3:
            body[100::1000] = (body[100::1000] + 1) & 3
        gs.append(Genome(name=f"g{i}", is_ingroup=i < 3, records=[acgt[body]]))
    try:
        res = _search_panel(searcher, gs, 25, 1, 2, options={"profile": 1})
    finally:
        _reset(searcher)
    assert res.rows() == _oracle_panel(gs, 25, 1, 2)


def test_host_buffers_in_batches_take_the_slab_path(searcher):
    """Pinned host buffers: K1 + level 0 run per batch of arrived files and append to the same slabs."""
    import torch
    from krisp_b200.panel import make_panel
    gs = make_panel(10, 10, 1_000_000)
    want = _oracle_panel(gs, 25, 1, 2)
    pinned = []
    for g in gs:
        t = torch.empty(len(g.joined()), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = np.frombuffer(g.joined(), dtype=np.uint8)
        pinned.append(t)
    try:
        searcher.configure(25, 1, 2, [1 if g.is_ingroup else 0 for g in gs])
        searcher.set_option("profile", 1)
        for _ in range(2):
            searcher.clear_sequences()
            searcher.reserve(sum(t.numel() + 1 for t in pinned))
            for i, t in enumerate(pinned):
                searcher.add_sequence(i, t.numpy())
            res = searcher.search(have_outgroup=True)
            assert "K1 extract + partition 0 + partition 1 per batch" in _stages(res)
            assert res.rows() == want
    finally:
        _reset(searcher)


def test_file_id_outside_the_configuration_is_an_error(searcher):
    """An id >= n_files would set a presence bit outside the "all files" mask and silently empty the result."""
    from krisp_b200._lib import KrispB200Error
    searcher.configure(5, 1, 2, [1, 0])
    searcher.clear_sequences()
    searcher.add_sequence(0, np.frombuffer(b"ACGTACGTACGTAA", dtype=np.uint8))
    searcher.add_sequence(3, np.frombuffer(b"ACGTACGTACGTAA", dtype=np.uint8))
    with pytest.raises(KrispB200Error):
        searcher.search()
    searcher.clear_sequences()


def test_raw_fasta_files_in_batches(searcher):
    """kb_add_fasta is asynchronous: raw file bytes travel on the copy stream, the de-lining kernels and K1 + the partition levels run
    per batch of arrived files.  Rows == oracle, twice in a row (buffers reused), flags clean."""
    import torch
    from krisp_b200.panel import make_panel
    gs = make_panel(10, 10, 1_000_000)
    want = _oracle_panel(gs, 25, 1, 2)
    pinned = []
    for g in gs:
        raw = np.frombuffer(g.fasta_text(), dtype=np.uint8)
        t = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = raw
        pinned.append(t)
    try:
        searcher.configure(25, 1, 2, [1 if g.is_ingroup else 0 for g in gs])
        searcher.set_option("profile", 1)
        for _ in range(2):
            searcher.clear_sequences()
            searcher.reserve(sum(t.numel() + 1 for t in pinned))
            for i, t in enumerate(pinned):
                searcher.add_fasta(i, t.numpy())
            res = searcher.search(have_outgroup=True)
            assert res.rows() == want
            assert searcher.fasta_flags() == 0
        # pageable bytes objects work too (the CLI path)
        searcher.clear_sequences()
        for i, g in enumerate(gs):
            searcher.add_fasta(i, g.fasta_text())
        assert searcher.search(have_outgroup=True).rows() == want
    finally:
        _reset(searcher)
