#!/usr/bin/env python3
"""bench.py — diagnostic-region search throughput (input Gbp/s) on synthetic multi-genome panels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genome-len BP]

A "step" is one complete search: K1 fused with partition level 0 (bases -> packed records in level-0 slabs) -> partition level 1
-> bucket hash aggregation (intersection over all files + diagnostic filter) -> survivor table and CSV rows on the host.
N=1 workload = BASELINE.json configs[1]: 20 ingroup + 20 outgroup 5 Mbp genomes, --conserved-left 25 --diagnostic 1
--conserved-right 2.  For N>1 (torchrun, one rank per GPU) the genomes are N times longer (weak scaling: 0.2 Gbp per GPU), every
rank ingests its files (round-robin), and the level-0 slabs travel once to the GPU that owns their flank-key range as bulk peer
copies over NVLink, digit group by digit group, while the owner already works on the previous group (krisp_b200/sharded.py).

`value`   : panel bases / device time of K steps, sequences already resident in HBM.
`e2e`     : same through the C ABI from RAW FASTA FILE BYTES in pinned host memory (kb_add_fasta): H2D of every byte, de-lining on the
            device, one D2H of the result image (survivor table + the rows, which arrive as text, rendered and ordered on the device)
            inside the timed region; `e2e.parsed_sequences` = the same from already parsed sequences (kb_add_sequence).
`roofline`: the kernel family with the largest share of the step; `kernel_families` lists all of them; `whole_step` = as-built
            algorithmic bytes of the whole search / device time.
`parity`  : (N>1) before timing, a sharded search of a panel small enough for the CPU oracle (25/1/2 and 32/60/32) is compared with
            the oracle's rows; the bench aborts if they differ.
`nvlink`  : (N>1) bytes every GPU sends / time from the first send to the last group landed, against 900 GB/s.
`also`    : (N=1) short runs of the other BASELINE shapes that fit one GPU: C3 primer mode (32/60/32) and the top of the C5 sweep.
`--impl reference` : the CPU restatement of the reference's algorithm (oracle/krisp_oracle.c, OpenMP, every host core — set
            explicitly, torchrun exports OMP_NUM_THREADS=1) on a bounded sample of the same workload.  The unmodified Python
            reference cannot travel to the GPU box; its timing in the build container is quoted as `python_reference`
            (tools/time_reference.py -> profiles/r04_python_reference.json).
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


def cpu_baseline(sample_len, steps=1, warmup=0):
    """The oracle port on a 20+20 x sample_len panel, all host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1)."""
    from krisp_b200.panel import make_panel
    from oracle import oracle
    gs = make_panel(N_IN, N_OUT, sample_len)
    recs = [[r.tobytes() for r in g.records] for g in gs]
    labels = [g.name for g in gs]
    ing = {g.name for g in gs if g.is_ingroup}
    bases = sum(g.n_bases for g in gs)
    cores = os.cpu_count() or 1
    for _ in range(warmup):
        oracle.search_records(recs, labels, ing, True, L_, D_, R_, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        rows, _ = oracle.search_records(recs, labels, ing, True, L_, D_, R_, nthreads=cores)
    dt = (time.perf_counter() - t0) / steps
    return {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{N_IN}+{N_OUT} genomes x {sample_len} bp ({bases / 1e6:.1f} Mbp), {L_}/{D_}/{R_}, oracle/krisp_oracle.c (OpenMP, {cores} threads)",
            "seconds_per_step": dt, "rows": len(rows)}


def python_reference_record():
    """The UNMODIFIED Python reference cannot travel to the GPU box; its timing on the scaled panel of BASELINE.md
    section 3 (20+20 x 100 kbp), taken in the build container by tools/time_reference.py, is committed under profiles/."""
    p = os.path.join(ROOT, "profiles", "r04_python_reference.json")
    try:
        with open(p) as fh:
            d = json.loads(fh.read().strip().splitlines()[-1])
        return {k: d[k] for k in ("value", "unit", "cores", "kind", "sample", "seconds", "rows", "rows_equal_c_oracle") if k in d} | \
               {"where": "build container (8 cores), not this box; see tools/time_reference.py"}
    except Exception:
        return None


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
            "python_reference": python_reference_record(),
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


_FAMILIES = {
    "K1": ("kb_extract_part_kernel (K1 fused with partition level 0: bases -> records in level-0 slabs)", lambda k: k.startswith("K1 extract"),
           lambda bases, n: bases + 8.0 * n),
    "K2": ("kb_part_kernel (K2: one radix-partition level, read + write of every record)", lambda k: k.startswith("K2 partition") or k.startswith("K2 pass"),
           lambda bases, n: 16.0 * n),
    "K3": ("kb_hash_warp_kernel (K3: bucket hash aggregation, one read of every record)", lambda k: k.startswith("K3 bucket hash") or k == "K3 group",
           lambda bases, n: 8.0 * n),
    "K3a": ("kb_prefilter_kernel (K3a: flank-hash presence filter, one read of every element)", lambda k: k == "K3a hash prefilter", lambda bases, n: 8.0 * n),
}


def ncu_traffic(kname, n_rec):
    """DRAM bytes (read + written) of one launch of the kernel from the committed ncu capture (profiles/r05_kernel_traffic.json) — only
    while the kernel's source files are still the ones that were profiled (sha256), and only for the headline workload it was taken on."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "r05_kernel_traffic.json")) as fh:
            doc = json.load(fh)
        if int(n_rec) != int(doc["records_per_launch"]):
            return None, "the ncu capture in profiles/r05_kernel_traffic.json is of the headline workload (panel C2 on one GPU), not of this one"
        rec = doc["kernels"]
        k = next(v for name, v in rec.items() if name in kname)
        h = hashlib.sha256()
        for f in k["source_files"]:
            with open(os.path.join(ROOT, "krisp_b200", "csrc", f), "rb") as fh:
                h.update(fh.read())
        if h.hexdigest() != k["source_sha256"]:
            return None, "the kernel's sources changed since the ncu capture in profiles/r05_kernel_traffic.json: traffic not reported"
        return k["dram_bytes"], f"ncu --set full, one launch on panel C2 (profiles/r05_kernel_traffic.json: {k['kernel']}, sources unchanged since)"
    except Exception as exc:
        return None, f"no ncu record ({exc!r})"[:160]


def roofline_of(prof, local_bases, n_rec, direct, peak, peak_src):
    """The kernel family with the largest share of the step: algorithmic bytes of the family per step / its device time per step."""
    roof, best, fams = None, -1.0, {}
    for key, (kname, match, nbytes) in _FAMILIES.items():
        if key == "K3" and not direct:
            continue                                      # (multi-word records: the exact pass only sees what the hash filter keeps)
        per = {k: sum(v) / len(v) for k, v in prof.items() if match(k)}
        if not per:
            continue
        total = sum(per.values())
        b = nbytes(local_bases, n_rec) * (len({k.split(" (group")[0] for k in per}) if key == "K2" else 1)   # (K2: every level moves all records)
        achieved = b / (total * 1e-3) / 1e9
        fams[key] = {"ms_per_step": total, "achieved_gbs": achieved, "frac": achieved / peak, "launches_per_step": len(per)}
        if total > best:
            best = total
            roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_step": b, "family_ms_per_step": total, "launches_per_step": len(per)}
    if roof:
        roof["traffic"], roof["traffic_note"] = ncu_traffic(roof["kernel"], n_rec)
    return roof, fams


def fold_stages(prof):
    """Mean ms per stage; the per-group stages of the pipelined multi-GPU search are also summed per kernel."""
    out, fold = {}, {}
    for k, v in prof.items():
        m = sum(v) / len(v)
        out[k] = m
        if " (group " in k:
            fold[k.split(" (group ")[0] + " (all groups)"] = fold.get(k.split(" (group ")[0] + " (all groups)", 0.0) + m
    out.update(fold)
    return out


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and thereby the pinned host buffers it is about to allocate: first touch) to the NUMA node its GPU hangs
    off, so that the host -> device copies of concurrent ranks do not cross the socket interconnect.  Returns a record for the JSON
    line; never fatal."""
    rec = {"node": None, "cpus": None}
    try:
        import torch
        pci = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        bus = None
        if pci is not None:
            dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
            dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
            bus = f"{dom:04x}:{pci:02x}:{dev:02x}.0"
        if bus is None:
            return rec
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        rec["pci"] = bus
        rec["node"] = node
        if node < 0:
            return rec
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            rec["cpus"] = len(allowed)
    except Exception as exc:                                           # (containers without /sys access: leave the default placement)
        rec["error"] = repr(exc)[:120]
    return rec


def run_ours(args):
    import hashlib
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
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and os.environ.get("KRISP_NUMA_BIND", "1") == "1" else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = measured_peak()
    stream = torch.cuda.current_stream()
    s = Searcher(device=local_rank, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Workload:
        """One panel on this rank: pinned host copies (e2e arm) and device-resident copies (value arm)."""

        def __init__(self, n_in, n_out, genome_len, ldr, panel_kw=None, fasta=True):
            from krisp_b200.panel import make_genome
            self.n_in, self.n_out, self.genome_len, self.ldr = n_in, n_out, genome_len, ldr
            self.total_bases = (n_in + n_out) * genome_len
            self.is_in = [1] * n_in + [0] * n_out
            self.pinned, self.fasta = [], []
            for i in range(n_in + n_out):
                if i % world != rank:
                    continue
                g = make_genome(i, i < n_in, i < n_in and (i % 2 == 1), f"ingroup{i}" if i < n_in else f"outgroup{i - n_in}", genome_len, **(panel_kw or {}))
                arr = np.frombuffer(g.joined(), dtype=np.uint8)
                t = torch.empty(arr.size, dtype=torch.uint8, pin_memory=True)
                t.numpy()[:] = arr
                self.pinned.append((i, t))
                if fasta:                                   # the file as it is on disk: '>' header lines, 80-column lines
                    raw = np.frombuffer(g.fasta_text(), dtype=np.uint8)
                    tf = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True)
                    tf.numpy()[:] = raw
                    self.fasta.append((i, tf))
            self.resident = [(gid, t.to(dev)) for gid, t in self.pinned]
            self.h2d_bytes = sum(t.numel() for _, t in self.pinned)
            self.fasta_bytes = sum(t.numel() for _, t in self.fasta)

        def configure(self, ldr=None):
            if ldr:
                self.ldr = ldr
            s.configure(*self.ldr, self.is_in)
            s.set_option("profile", 1)
            for k, v in args.option or []:
                s.set_option(k, int(v))
            s.reserve(max(self.h2d_bytes, self.fasta_bytes) + len(self.pinned))

        def load_resident(self):
            s.clear_sequences()
            for gid, t in self.resident:
                s.add_sequence(gid, (t.data_ptr(), t.numel()))

        def load_host(self):
            s.clear_sequences()
            for gid, t in self.pinned:
                s.add_sequence(gid, t.numpy())

        def load_fasta(self):
            s.clear_sequences()
            for gid, t in self.fasta:
                s.add_fasta(gid, t.numpy())

        def search(self):
            if world == 1:
                return s.search(have_outgroup=True)
            return sharded.sharded_search(s, dev, have_outgroup=True, total_bases=self.total_bases)

        def timed(self, load, steps, collect_rows, sampler=None):
            out = {"launches": 0, "prof": {}, "host_ms": {}, "d2h": 0}
            barrier()
            if sampler:
                sampler.mark_begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            res = None
            for i in range(steps):
                res = None                                   # (a caller that does not keep a result does not pay for host copies of it)
                t_a = time.perf_counter()
                if load is not None:
                    load()
                t_b = time.perf_counter()
                res = self.search()
                t_c = time.perf_counter()
                if collect_rows:
                    out["rows"] = res.csv_rows_bytes()       # the CSV body krisp_fasta prints (rendered + ordered on the device), as bytes
                t_d = time.perf_counter()
                for nm, dt in (("load", t_b - t_a), ("search", t_c - t_b), ("rows", t_d - t_c)):
                    out["host_ms"][nm] = out["host_ms"].get(nm, 0.0) + dt * 1e3 / steps
                out["launches"] += s.last_counters()["kernel_launches"]
                per_step = {}
                for name, ms in res.profile:                          # (a stage name may occur twice in one search, e.g. the bucket hash
                    per_step[name] = per_step.get(name, 0.0) + ms     #  and its deferred-bucket fallback: they add up, they do not average)
                for name, ms in per_step.items():
                    out["prof"].setdefault(name, []).append(ms)
                # one copy of the result image: flank words | ingroup sets | outgroup sets | rows text (+ 72 B of counters)
                out["d2h"] = res.n_groups * (8 * res.flank_word_count + 2 * 4 * res.mask_word_count + res.row_bytes) + 72
            out["last"] = res
            if collect_rows:
                out["rows"] = out["rows"].decode("ascii")    # (outside the timed region: the parity checks below compare text)
            e1.record(stream)
            barrier()
            if sampler:
                sampler.mark_end()
            ms = e0.elapsed_time(e1) / steps
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            out["ms"] = ms
            return out

        def measure(self, steps, warmup, sampler=None):
            """value arm (resident) + e2e arm (host buffers) -> dict for the JSON line."""
            self.load_resident()
            s.synchronize()
            self.timed(None, warmup, False)
            dv = self.timed(None, steps, False, sampler)
            counters = s.last_counters()
            # e2e: raw FASTA bytes (what the krisp_fasta command line reads) from pinned host memory -> rows on the host; de-lining on
            # the device.  The same from already parsed sequences (the reference parser's output) is kept as e2e.parsed_sequences
            loader = self.load_fasta if self.fasta else self.load_host
            self.timed(loader, 1, True)
            de = self.timed(loader, steps, True)
            dp = None
            if self.fasta:
                self.timed(self.load_host, 1, True)
                dp = self.timed(self.load_host, steps, True)
                assert dp["rows"] == de["rows"], "rows from raw FASTA differ from rows from parsed sequences"
            rows = de["rows"]
            assert sorted(rows.splitlines()) == de["last"].rows(), "device-rendered rows differ from the host decoder's"   # (outside the timed region)
            n_rows = rows.count("\n")
            if world > 1:
                t = torch.tensor([n_rows], device=dev)
                dist.all_reduce(t)
                n_rows = int(t.item())
            last = dv["last"]
            direct = 2 * sum(self.ldr) + 8 <= 64
            roof, fams = roofline_of(dv["prof"], self.h2d_bytes, last.n_records, direct, peak, peak_src)
            rows_sha = hashlib.sha256("\n".join(sorted(rows.splitlines())).encode()).hexdigest() if world == 1 else None
            return {"ms": dv["ms"], "value": self.total_bases / (dv["ms"] * 1e-3) / 1e9, "launches": dv["launches"], "counters": counters,
                    "stage_ms": fold_stages(dv["prof"]), "roofline": roof, "families": fams, "last": last, "rows": n_rows, "rows_sha256": rows_sha,
                    "e2e": {"value": self.total_bases / (de["ms"] * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": de["ms"],
                            "h2d_bytes_per_step": (self.fasta_bytes if self.fasta else self.h2d_bytes) * world, "d2h_bytes_per_step": de["d2h"],
                            "input": "raw FASTA file bytes (headers + 80-column lines), de-lined on the device (kb_add_fasta)" if self.fasta else "parsed sequences",
                            "host_ms": de["host_ms"], "stage_ms": fold_stages(de["prof"]),
                            "parsed_sequences": None if dp is None else {"value": self.total_bases / (dp["ms"] * 1e-3) / 1e9, "ms_per_step": dp["ms"],
                                                                         "h2d_bytes_per_step": self.h2d_bytes * world}},
                    "exchange": getattr(last, "exchange", None)}

    # ---- N > 1: parity of the sharded search against the CPU oracle, before anything is timed ------------------------------------
    parity = None
    if world > 1:
        from krisp_b200.panel import make_panel
        parity = {"n": world, "cases": [], "ok": True}
        for ldr in ((25, 1, 2), (32, 60, 32)):
            gs = make_panel(5, 4, 200_000)
            s.configure(*ldr, [1 if g.is_ingroup else 0 for g in gs])
            s.clear_sequences()
            for i, g in enumerate(gs):
                if i % world == rank:
                    s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
            res = sharded.sharded_search(s, dev, have_outgroup=True)
            rows = sharded.gather_rows(res.rows())
            sha = hashlib.sha256("\n".join(rows).encode()).hexdigest()
            case = {"panel": "5+4 x 200 kbp", "ldr": list(ldr), "rows": len(rows), "rows_sha256": sha, "exchange": "slab" if getattr(res, "exchange", {}).get("slab") else "exact"}
            if rank == 0:
                from oracle import oracle
                want, _ = oracle.search_records([[r.tobytes() for r in g.records] for g in gs], [g.name for g in gs],
                                                {g.name for g in gs if g.is_ingroup}, True, *ldr, nthreads=os.cpu_count() or 1)
                case["oracle_sha256"] = hashlib.sha256("\n".join(want).encode()).hexdigest()
                case["ok"] = case["oracle_sha256"] == sha and len(want) > 0
                parity["ok"] = parity["ok"] and case["ok"]
            parity["cases"].append(case)
        ok = torch.tensor([1 if parity["ok"] else 0], device=dev)
        dist.broadcast(ok, 0)
        if not int(ok.item()):
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "sharded search differs from the CPU oracle", "parity": parity}))
            dist.destroy_process_group()
            raise SystemExit(1)

    # ---- headline workload ------------------------------------------------------------------------------------------------------
    genome_len = args.genome_len * world           # weak scaling: 0.2 Gbp of input per GPU
    w = Workload(N_IN, N_OUT, genome_len, (L_, D_, R_), PANEL_KW)
    w.configure()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    m = w.measure(args.steps, args.warmup, clocks if rank == 0 else None)
    clk = clocks.stop() if rank == 0 else None
    last, counters = m["last"], m["counters"]
    n_rec = last.n_records
    whole = {"algorithmic_bytes_per_step": counters["algorithmic_bytes"],
             "achieved_gbs": counters["algorithmic_bytes"] / (m["ms"] * 1e-3) / 1e9,
             "frac_of_peak": counters["algorithmic_bytes"] / (m["ms"] * 1e-3) / 1e9 / peak,
             "note": "as-built bytes: K1 fused with partition level 0 (bases read + records written once), one more partition level (read + write), "
                     "one read by the bucket hash = 1 + 4*8*(1 + P) B/bp with P = 1 standalone pass (SURVEY 8d); round 1 moved 97 B/bp (P = 2)",
             "frac_of_peak_at_round1_bytes": (w.h2d_bytes + 48.0 * n_rec) / (m["ms"] * 1e-3) / 1e9 / peak}

    if args.diag and world > 1:
        variants = [("default", {}, {}), ("copy_streams=1", {"KRISP_COPY_STREAMS": "1"}, {}), ("copy_streams=2", {"KRISP_COPY_STREAMS": "2"}, {}),
                    ("copy_streams=3", {"KRISP_COPY_STREAMS": "3"}, {}), ("groups=8", {"KRISP_SLAB_GROUPS": "8"}, {}),
                    ("groups=8, copy_streams=2", {"KRISP_SLAB_GROUPS": "8", "KRISP_COPY_STREAMS": "2"}, {}), ("own_first=0", {"KRISP_OWN_FIRST": "0"}, {})]
        os.environ["KRISP_TIMELINE"] = "1"
        for name, env, opts in variants:
            for k, v in env.items():
                os.environ[k] = v
            for k, v in opts.items():
                s.set_option(k, v)
            s.__dict__.pop("_slab_plan", None)
            try:
                w.load_resident()
                s.synchronize()
                w.timed(None, 2, False)
                dv = w.timed(None, 5, False)
                ex = getattr(dv["last"], "exchange", None) or {}
                rec = {"diag": name, "n_gpus": world, "ms": dv["ms"], "stage_ms": {k: round(v, 3) for k, v in fold_stages(dv["prof"]).items()},
                       "exchange_ms": ex.get("exchange_ms"), "groups": ex.get("groups"), "window_items": ex.get("window_items"),
                       "timeline": {k: ([round(x, 3) for x in v] if isinstance(v, list) else round(v, 3)) for k, v in (ex.get("timeline") or {}).items()}}
            except Exception as exc:                                   # (a variant that does not apply must not end the run)
                rec = {"diag": name, "error": repr(exc)}
            if rank == 0:
                print(json.dumps(rec), file=sys.stderr, flush=True)
            for k in env:
                os.environ.pop(k, None)
            for k in opts:
                s.set_option(k, -1 if k == "sym" else 0)
            s.__dict__.pop("_slab_plan", None)
        os.environ.pop("KRISP_TIMELINE", None)

    # ---- also: the other BASELINE configurations that fit one GPU (N = 1 only; short runs) ------------------------------------------
    also = []
    if world == 1 and not args.no_also and not args.ldr and not args.genomes:
        def sub(name, wl, steps=3, warmup=1):
            mm = wl.measure(steps, warmup)
            r = mm["roofline"] or {}
            return {"config": name, "value": mm["value"], "unit": UNIT, "ms_per_step": mm["ms"], "steps": steps, "warmup": warmup,
                    "e2e": {"value": mm["e2e"]["value"], "ms_per_step": mm["e2e"]["ms_per_step"]}, "rows": mm["rows"], "rows_sha256": mm["rows_sha256"],
                    "records": int(mm["last"].n_records), "dominant_kernel": r.get("kernel"), "frac": r.get("frac"),
                    "stage_ms": {k: round(v, 4) for k, v in mm["stage_ms"].items()}}
        w.configure((32, 60, 32))
        also.append(sub("C3 primer mode: the same 20+20 x 5 Mbp panel, --conserved-left 32 --diagnostic 60 --conserved-right 32 (124-mers, multi-word records)", w))
        w.configure((L_, D_, R_))
        del w
        torch.cuda.empty_cache()
        w5 = Workload(100, 100, args.genome_len, (L_, D_, R_), PANEL_KW)
        w5.configure()
        also.append(sub("C5 genome-count sweep, top end: 100+100 x 5 Mbp (1 Gbp), 25/1/2", w5))
        del w5

    if rank != 0:
        if world > 1:
            sharded.shutdown(s)
            dist.destroy_process_group()
        return
    cb = cpu_baseline(args.cpu_sample_len) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (2-bit packed bases)", "data": "synthetic",
        "config": {"workload": f"synthetic bacterial panel: {N_IN} ingroup + {N_OUT} outgroup genomes x {genome_len} bp "
                               f"with planted group SNPs, --conserved-left {L_} --diagnostic {D_} --conserved-right {R_} ({L_ + D_ + R_}-mer, "
                               f"{'one 64-bit record' if 2 * (L_ + D_ + R_) + 8 <= 64 else 'multi-word records'})",
                   "total_bases": N_IN * genome_len + N_OUT * genome_len, "records": int(n_rec) if world == 1 else None,
                   "radix_passes": counters["radix_passes"], "rows": m["rows"], "rows_sha256": m["rows_sha256"], "k3_stats": dict(last.stats),
                   "l2": "inputs larger than L2 (>= 0.2 GB of bases, 3.2 GB of records per GPU; 126 MB L2)",
                   "parallelism": f"flank-hash sharded x{world}" if world > 1 else "single GPU"},
        "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": clk, "roofline": m["roofline"], "kernel_families": m["families"],
        "whole_step": whole, "stage_ms": m["stage_ms"],
    }
    if world > 1:
        line["parity"] = parity
        # the e2e arm at N > 1 is bound by the host: N ranks read their pinned buffers at once (copy phase = the K1 stage of the e2e arm)
        k1 = next((v for k, v in m["e2e"]["stage_ms"].items() if k.startswith("K1 extract")), None)
        line["e2e"]["h2d"] = {"bytes_per_gpu": m["e2e"]["h2d_bytes_per_step"] // world, "copy_phase_ms": k1,
                              "gbs_per_gpu": (m["e2e"]["h2d_bytes_per_step"] / world / (k1 * 1e-3) / 1e9) if k1 else None,
                              "numa_binding_rank0": numa}
        ex = m["exchange"] or {}
        if ex.get("slab") and ex.get("exchange_ms"):
            sent = float(ex["sent_bytes"])
            line["nvlink"] = {"bytes_sent_per_gpu": sent, "bytes_copied_per_gpu": float(ex.get("copied_bytes", sent)), "ms": ex["exchange_ms"],
                              "achieved_gbs": sent / (ex["exchange_ms"] * 1e-3) / 1e9, "peak": 900.0, "frac_of_900": sent / (ex["exchange_ms"] * 1e-3) / 1e9 / 900.0,
                              "groups": ex.get("groups"), "how": "bulk peer copies (copy engines) of whole level-0 slabs, digit group by digit group, "
                              "from the first send to the last group landed (rank 0); level 1 + bucket hash of group g run under the copies of group g + 1"}
    if also:
        line["also"] = also
    if cb:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["python_reference"] = python_reference_record()
    print(json.dumps(line))
    if world > 1:
        sharded.shutdown(s)
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
    ap.add_argument("--no-also", action="store_true", help="skip the short C3 / C5 sub-records of the N = 1 line")
    ap.add_argument("--genomes", nargs=2, type=int, metavar=("N_IN", "N_OUT"), help="ingroup / outgroup genome counts (default 20 20; "
                    "50 50 = the shape of BASELINE config 4, 5..100 each = the genome-count sweep of config 5)")
    ap.add_argument("--ldr", nargs=3, type=int, metavar=("L", "D", "R"), help="--conserved-left / --diagnostic / --conserved-right "
                    "(default 25 1 2 = BASELINE config 2; 32 60 32 = config 3, primer mode)")
    ap.add_argument("--noise", type=float, help="private substitution rate per genome (default 1e-3); higher = more divergent genomes, "
                    "fewer shared k-mers (robustness probe, not the BASELINE workload)")
    ap.add_argument("--diag", action="store_true", help="N > 1: after the headline measurement, time the resident search under exchange variants "
                    "(digit groups, copy streams, records instead of window items, NCCL all-to-all) and print one JSON line each on stderr")
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
