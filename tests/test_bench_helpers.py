"""bench.py's host-side helpers (no GPU): stage folding, the ncu traffic record with its source-hash guard, NUMA binding that must
never be fatal, the digit-group choice of the sharded search."""
import importlib.util
import json
import os

from tests.helpers import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fold_stages_sums_groups_and_means_steps():
    b = _bench()
    prof = {"K1 extract + partition 0": [1.0, 3.0], "K3 bucket hash (group 0)": [0.5, 0.5], "K3 bucket hash (group 1)": [0.25, 0.75]}
    out = b.fold_stages(prof)
    assert out["K1 extract + partition 0"] == 2.0
    assert out["K3 bucket hash (all groups)"] == 1.0


def test_ncu_traffic_is_reported_only_for_the_profiled_sources_and_workload():
    b = _bench()
    with open(os.path.join(ROOT, "profiles", "r05_kernel_traffic.json")) as fh:
        doc = json.load(fh)
    n = doc["records_per_launch"]
    for name, rec in doc["kernels"].items():
        import hashlib
        h = hashlib.sha256()
        for f in rec["source_files"]:
            with open(os.path.join(ROOT, "krisp_b200", "csrc", f), "rb") as fh:
                h.update(fh.read())
        traffic, note = b.ncu_traffic(name + " (whatever follows)", n)
        if h.hexdigest() == rec["source_sha256"]:
            assert traffic == rec["dram_bytes"] and 0.9 < traffic / (rec["dram_bytes_read"] + rec["dram_bytes_write"]) < 1.1
        else:
            assert traffic is None and "changed" in note
        assert b.ncu_traffic(name, n + 1)[0] is None           # another workload: no number


def test_numa_binding_is_never_fatal():
    b = _bench()
    rec = b.bind_to_gpu_numa_node(0)                           # (no GPU here: it must come back with an empty record)
    assert isinstance(rec, dict) and "node" in rec


def test_slab_groups_choice():
    from krisp_b200 import sharded
    os.environ.pop("KRISP_SLAB_GROUPS", None)
    assert sharded._slab_groups(32) == 4 and sharded._slab_groups(2) == 2 and sharded._slab_groups(1) == 1
    os.environ["KRISP_SLAB_GROUPS"] = "8"
    try:
        assert sharded._slab_groups(32) == 8 and sharded._slab_groups(4) == 4
    finally:
        os.environ.pop("KRISP_SLAB_GROUPS", None)
