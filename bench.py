#!/usr/bin/env python3
"""bench.py — diagnostic-region search throughput (input Gbp/s) on synthetic multi-genome panels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genome-len BP]

A "step" is one complete search (K1 extract -> K2 radix sort -> K3 group/filter -> survivor table on
the host) over one synthetic panel.  N=1 workload = BASELINE.json configs[1]: 20 ingroup + 20 outgroup
5 Mbp genomes, --conserved-left 25 --diagnostic 1 --conserved-right 2.  For N>1 (torchrun, one rank per
GPU) the genomes are N times longer (weak scaling: 0.2 Gbp per GPU), every rank ingests 40/N... files
round-robin, and records are exchanged once by flank-hash shard (NCCL all-to-all).

`value`  : panel bases / device time of K steps, sequences already resident in HBM.
`e2e`    : same through the C ABI with pinned HOST buffers: H2D of every base and D2H of the survivor
           table inside the timed region, rows decoded on the host.
`--impl reference` : the CPU restatement of the reference's algorithm (oracle/krisp_oracle.c, all host
           threads) on a bounded sample of the same workload (the reference itself is pure Python at
           ~4e-5 Gbp/s).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_, D_, R_ = 25, 1, 2
N_IN, N_OUT = 20, 20
PANEL_KW = {}            # --noise: private substitution rate of the synthetic genomes (default: the panel generator's 1e-3)
METRIC = "diagnostic-region search throughput (input bases / s)"
UNIT = "Gbp/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        """Launch nvidia-smi (20 ms period) and wait until it delivers samples; mark() brackets the timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
            t = time.time()
            while not self.lines and time.time() - t < 5.0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = (self.t0 or 0) - 0.02
        t1 = (self.t1 or 1e18) + 0.02
        for ts, ln in self.lines:
            if not (t0 <= ts <= t1):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def build_panel(genome_len, rank=0, world=1):
    """This rank's genomes of the 20+20 panel: (global file id, is_ingroup, packed uint8 array)."""
    from krisp_b200.panel import make_genome
    import numpy as np
    out = []
    for i in range(N_IN + N_OUT):
        if i % world != rank:
            continue
        is_in = i < N_IN
        g = make_genome(i, is_in, is_in and (i % 2 == 1), f"ingroup{i}" if is_in else f"outgroup{i - N_IN}", genome_len, **PANEL_KW)
        out.append((i, is_in, np.frombuffer(g.joined(), dtype=np.uint8)))
    return out


def cpu_baseline(sample_len, steps=1, warmup=0):
    """The oracle port on a 20+20 x sample_len panel, all host threads."""
    from krisp_b200.panel import make_panel
    from oracle import oracle
    gs = make_panel(N_IN, N_OUT, sample_len)
    recs = [[r.tobytes() for r in g.records] for g in gs]
    labels = [g.name for g in gs]
    ing = {g.name for g in gs if g.is_ingroup}
    bases = sum(g.n_bases for g in gs)
    for _ in range(warmup):
        oracle.search_records(recs, labels, ing, True, L_, D_, R_)
    t0 = time.perf_counter()
    for _ in range(steps):
        rows, _ = oracle.search_records(recs, labels, ing, True, L_, D_, R_)
    dt = (time.perf_counter() - t0) / steps
    return {"value": bases / dt / 1e9, "unit": UNIT, "cores": oracle.threads(), "kind": "port",
            "sample": f"{N_IN}+{N_OUT} genomes x {sample_len} bp ({bases / 1e6:.1f} Mbp), {L_}/{D_}/{R_}, oracle/krisp_oracle.c (OpenMP)",
            "seconds_per_step": dt, "rows": len(rows)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.cpu_sample_len, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": min(args.warmup, 1), "ms_per_step": cb["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"synthetic bacterial panel {N_IN}+{N_OUT} genomes, {L_}/{D_}/{R_}; bounded sample: {cb['sample']}"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    from krisp_b200.search import Searcher
    from krisp_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — krisp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    genome_len = args.genome_len * world           # weak scaling: 0.2 Gbp of input per GPU
    mine = build_panel(genome_len, rank, world)
    total_bases = (N_IN + N_OUT) * genome_len       # bases of the whole panel (separators excluded: 4 records/genome)
    is_in = [1] * N_IN + [0] * N_OUT

    stream = torch.cuda.current_stream()
    s = Searcher(device=local_rank, stream=stream.cuda_stream)
    s.configure(L_, D_, R_, is_in)
    s.set_option("profile", 1)
    for k, v in args.option or []:
        s.set_option(k, int(v))

    # pinned host copies (e2e arm) and device-resident copies (value arm)
    pinned = []
    for gid, _, arr in mine:
        t = torch.empty(arr.size, dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = arr
        pinned.append((gid, t))
    resident = [(gid, t.to(dev)) for gid, t in pinned]
    h2d_bytes = sum(t.numel() for _, t in pinned)
    s.reserve(h2d_bytes + len(pinned))

    def load_resident():
        s.clear_sequences()
        for gid, t in resident:
            s.add_sequence(gid, (t.data_ptr(), t.numel()))

    def load_host():
        s.clear_sequences()
        for gid, t in pinned:
            s.add_sequence(gid, t.numpy())

    def search():
        if world == 1:
            return s.search(have_outgroup=True)
        return sharded.sharded_search(s, dev, have_outgroup=True, total_bases=total_bases)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    results = {}

    host_ms = {}

    def timed(load, steps, collect_rows, sampler=None):
        launches, prof_acc, d2h = 0, {}, 0
        host_ms.clear()
        barrier()
        if sampler:
            sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            t_a = time.perf_counter()
            if load is not None:
                load()
            t_b = time.perf_counter()
            res = search()
            t_c = time.perf_counter()
            if collect_rows:
                results["rows"] = res.csv_rows_text()          # the CSV body krisp_fasta prints (rendered + ordered on the device)
            t_d = time.perf_counter()
            for nm, dt in (("load", t_b - t_a), ("search", t_c - t_b), ("rows", t_d - t_c)):
                host_ms[nm] = host_ms.get(nm, 0.0) + dt * 1e3 / steps
            launches += s.last_counters()["kernel_launches"]
            for name, ms in res.profile:
                prof_acc.setdefault(name, []).append(ms)
            d2h = res.n_groups * (8 * res.flank_words.shape[1] + 2 * 4 * res.in_words.shape[1] + 4 + 16 + res.row_bytes) + 64
            results["last"] = res
        e1.record(stream)
        barrier()
        if sampler:
            sampler.mark_end()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof_acc, d2h

    # ---- value arm: sequences resident in HBM ----------------------------------------------------
    load_resident()
    s.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    timed(None, args.warmup, False)
    ms_dev, launches, prof, _ = timed(None, args.steps, False, clocks if rank == 0 else None)
    clk = clocks.stop() if rank == 0 else None
    counters = s.last_counters()
    last = results["last"]

    # ---- e2e arm: pinned host buffers through the C ABI, rows decoded on the host -----------------
    timed(load_host, 1, True)
    ms_e2e, _, prof_e2e, d2h_bytes = timed(load_host, args.steps, True)
    e2e_host_ms = dict(host_ms)
    e2e_stage_ms = {k: sum(v) / len(v) for k, v in prof_e2e.items()}
    rows = results["rows"]
    n_rows_total = rows.count("\n")
    assert sorted(rows.splitlines()) == results["last"].rows(), "device-rendered rows differ from the host decoder's"   # (outside the timed region)
    if world > 1:
        t = torch.tensor([n_rows_total], device=dev)
        dist.all_reduce(t)
        n_rows_total = int(t.item())

    # ---- roofline of the dominant kernel: the kernel family with the largest share of the step --------------
    peak, peak_src = measured_peak()
    n_rec = last.n_records
    local_bases = h2d_bytes
    direct = 2 * (L_ + D_ + R_) + 8 <= 64        # (multi-word records: the exact K3 pass only sees what the hash filter keeps)
    fam = {"K1": ("kb_extract_kernel (K1: bases -> packed records)", lambda k: k.startswith("K1 extract"), local_bases + 8.0 * n_rec),
           "K2": ("kb_part_kernel (K2: one radix-partition level, read + write of every record)",
                  lambda k: k.startswith("K2 partition") or k.startswith("K2 pass"), 16.0 * n_rec),
           "K3": ("kb_hash_stream_kernel (K3: bucket hash aggregation, one read of every record)",
                  lambda k: k in ("K3 bucket hash", "K3 group") and direct, 8.0 * n_rec),
           "K3a": ("kb_prefilter_kernel (K3a: flank-hash presence filter, one read of every element)",
                   lambda k: k == "K3a hash prefilter", 8.0 * n_rec)}
    roof, best = None, -1.0
    for key, (kname, match, bytes_per_launch) in fam.items():
        per = [sum(v) / len(v) for k, v in prof.items() if match(k)]
        if not per:
            continue
        total = sum(per)
        if total > best:
            best = total
            avg = total / len(per)
            achieved = bytes_per_launch / (avg * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "avg_launch_ms": avg, "launches_per_step": len(per), "family_ms_per_step": total}
    if roof:
        ncu_traffic = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(ncu_traffic):
            with open(ncu_traffic) as fh:
                tj = json.load(fh)
            if tj.get("records") == n_rec and tj.get("kernel", "") in roof["kernel"]:
                roof["traffic"] = tj.get("dram_bytes_per_launch")
    stage_ms = {k: sum(v) / len(v) for k, v in prof.items()}
    whole = {"algorithmic_bytes_per_step": counters["algorithmic_bytes"],
             "achieved_gbs": counters["algorithmic_bytes"] / (ms_dev * 1e-3) / 1e9,
             "frac_of_peak": counters["algorithmic_bytes"] / (ms_dev * 1e-3) / 1e9 / peak}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cb = cpu_baseline(args.cpu_sample_len) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": METRIC, "value": total_bases / (ms_dev * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (2-bit packed bases)", "data": "synthetic",
        "config": {"workload": f"synthetic bacterial panel: {N_IN} ingroup + {N_OUT} outgroup genomes x {genome_len} bp "
                               f"with planted group SNPs, --conserved-left {L_} --diagnostic {D_} --conserved-right {R_} ({L_ + D_ + R_}-mer, "
                               f"{'one 64-bit record' if 2 * (L_ + D_ + R_) + 8 <= 64 else 'multi-word records'})",
                   "total_bases": total_bases, "records": int(n_rec) if world == 1 else None,
                   "radix_passes": counters["radix_passes"], "rows": n_rows_total, "k3_stats": dict(last.stats),
                   "l2": "inputs larger than L2 (>= 0.2 GB of bases, 3.2 GB of records per GPU; 126 MB L2)",
                   "parallelism": f"flank-hash sharded x{world}" if world > 1 else "single GPU"},
        "e2e": {"value": total_bases / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes,
                "host_ms": e2e_host_ms, "stage_ms": e2e_stage_ms},
        "gpu_launches": launches, "clocks": clk, "roofline": roof, "whole_step": whole, "stage_ms": stage_ms,
    }
    if cb:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-len", type=int, default=5_000_000, help="bases per genome per GPU (default: BASELINE config 2)")
    ap.add_argument("--cpu-sample-len", type=int, default=1_000_000, help="genome length of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--genomes", nargs=2, type=int, metavar=("N_IN", "N_OUT"), help="ingroup / outgroup genome counts (default 20 20; "
                    "50 50 = the shape of BASELINE config 4, 5..100 each = the genome-count sweep of config 5)")
    ap.add_argument("--ldr", nargs=3, type=int, metavar=("L", "D", "R"), help="--conserved-left / --diagnostic / --conserved-right "
                    "(default 25 1 2 = BASELINE config 2; 32 60 32 = config 3, primer mode)")
    ap.add_argument("--noise", type=float, help="private substitution rate per genome (default 1e-3); higher = more divergent genomes, "
                    "fewer shared k-mers (robustness probe, not the BASELINE workload)")
    ap.add_argument("--option", nargs=2, action="append", metavar=("NAME", "VALUE"), help="kb_set_option passthrough")
    args = ap.parse_args()
    global L_, D_, R_, N_IN, N_OUT
    if args.ldr:
        L_, D_, R_ = args.ldr
    if args.genomes:
        N_IN, N_OUT = args.genomes
    if args.noise is not None:
        PANEL_KW["noise"] = args.noise
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
