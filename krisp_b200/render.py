"""Host-side output: CSV rows and human-readable alignments from a SearchResult.

Restates the reference's renderers on the survivors only (kB-MB of data, off the timed path):
``render_output`` outputAlignments.py:101-162 (header :26-31), ``ConservedEndAmplicons.render_csv``
Amplicon.py:663-671, ``render_alignment`` :598-661, ``makeBracket`` :523-540, ``diagnosticColumns``
:483-493, ``ingroupUniqueColumns`` :495-521, ``Amplicon.__str__`` / ``_labelsToString`` :170-210.
Groups are written in ascending (left, right) order, which is what the reference produces with
``--cores 1`` (with more cores its order depends on process scheduling).
"""
import os
import sys

import numpy as np

from .search import _IUPAC, _decode_bases

CSV_HEADER = "left_seq,diag_seq,right_seq"

# ---- Primer3 post-filter (host, off the timed path; north_star: "Primer3 filtering ... stay on the host") ---------------------------
# The reference runs primer3-py on every surviving region and keeps the ones with at least one primer pair
# (ConservedEndAmplicons.find_primers, Amplicon.py:560-564; run_primer3 :103-151; CSV columns outputAlignments.py:10-31;
# alignment annotations Amplicon.py:640-660, _render_primer3_stats :566-595).  Same call, same settings, same columns here.
P3_COLS = ("PRIMER_PAIR_0_PRODUCT_SIZE", "PRIMER_PAIR_0_PENALTY", "PRIMER_LEFT_0_SEQUENCE", "PRIMER_RIGHT_0_SEQUENCE",
           "PRIMER_LEFT_0_PENALTY", "PRIMER_RIGHT_0_PENALTY", "PRIMER_LEFT_0_TM", "PRIMER_RIGHT_0_TM",
           "PRIMER_LEFT_0_GC_PERCENT", "PRIMER_RIGHT_0_GC_PERCENT", "PRIMER_LEFT_0_SELF_ANY_TH", "PRIMER_RIGHT_0_SELF_ANY_TH",
           "PRIMER_LEFT_0_SELF_END_TH", "PRIMER_RIGHT_0_SELF_END_TH", "PRIMER_LEFT_0_HAIRPIN_TH", "PRIMER_RIGHT_0_HAIRPIN_TH",
           "PRIMER_LEFT_0_END_STABILITY", "PRIMER_RIGHT_0_END_STABILITY", "PRIMER_PAIR_0_COMPL_ANY_TH", "PRIMER_PAIR_0_COMPL_END_TH")
P3_KEYS = tuple(n.replace("PRIMER_", "").replace("_0", "").lower() for n in P3_COLS)


class Primer3Unavailable(ImportError):
    """--primer3 was asked for but primer3-py is not importable (the reference cannot even start without it: Amplicon.py:3)."""


def run_primer3(template, target_start, target_len, tm=(53, 68), gc=(40, 70), amp_size=(80, 300), primer_size=(25, 35),
                max_sec_tm=40, gc_clamp=1, max_end_gc=4):
    """One primer3 design call on `template` = left + consensus + right with the diagnostic region as target — the settings of
    run_primer3 (Amplicon.py:103-151).  Returns primer3's result dict."""
    try:
        import primer3
    except ImportError as exc:
        raise Primer3Unavailable("--primer3 needs the primer3-py package (pip install primer3-py)") from exc
    from statistics import mean                 # (statistics.mean keeps [25, 35] -> 30 an int, as primer3 wants PRIMER_OPT_SIZE)
    settings = {
        "PRIMER_TASK": "generic", "PRIMER_PICK_LEFT_PRIMER": 1, "PRIMER_PICK_RIGHT_PRIMER": 1, "PRIMER_LIBERAL_BASE": 1,
        "PRIMER_OPT_SIZE": mean(primer_size), "PRIMER_MIN_SIZE": primer_size[0], "PRIMER_MAX_SIZE": primer_size[1],
        "PRIMER_OPT_TM": mean(tm), "PRIMER_MIN_TM": tm[0], "PRIMER_MAX_TM": tm[1], "PRIMER_MIN_GC": gc[0], "PRIMER_MAX_GC": gc[1],
        "PRIMER_MAX_POLY_X": 4, "PRIMER_MAX_NS_ACCEPTED": 0, "PRIMER_THERMODYNAMIC_OLIGO_ALIGNMENT": 1,
        "PRIMER_MAX_SELF_ANY_TH": max_sec_tm, "PRIMER_MAX_SELF_END_TH": max_sec_tm, "PRIMER_PAIR_MAX_COMPL_ANY_TH": max_sec_tm,
        "PRIMER_PAIR_MAX_COMPL_END_TH": max_sec_tm, "PRIMER_MAX_HAIRPIN_TH": max_sec_tm, "PRIMER_PRODUCT_SIZE_RANGE": [amp_size],
        "PRIMER_GC_CLAMP": gc_clamp, "PRIMER_MAX_END_GC": max_end_gc,
    }
    return primer3.bindings.design_primers({"SEQUENCE_TEMPLATE": template, "SEQUENCE_TARGET": [target_start, target_len]}, settings)


def _plain_table(header, rows):
    """Borderless left-aligned table: prettytable's ``get_string(border=False)`` with ``align = 'l'`` (one blank of padding on either
    side of every cell, no rules).  prettytable itself is used when it is importable."""
    try:
        from prettytable import PrettyTable
        t = PrettyTable(list(header))
        for r in rows:
            t.add_row(list(r))
        t.align = "l"
        return t.get_string(border=False)
    except ImportError:
        cells = [[str(x) for x in header]] + [[str(x) for x in r] for r in rows]
        width = [max(len(r[c]) for r in cells) for c in range(len(header))]
        return "\n".join("".join(" " + r[c].ljust(width[c]) + " " for c in range(len(header))) for r in cells)


def primer3_stats_text(p3):
    """The two statistics tables under an alignment (_render_primer3_stats, Amplicon.py:566-595)."""
    left = {k[14:]: v for k, v in p3.items() if "PRIMER_LEFT_0_" in k}
    right = {k[15:]: v for k, v in p3.items() if "PRIMER_RIGHT_0_" in k}
    pair = {k[14:]: v for k, v in p3.items() if "PRIMER_PAIR_0_" in k}
    names = lambda ks: [x.title().replace("_", " ") for x in ks]
    values = lambda vs: [str(round(x, 5)) if isinstance(x, float) else x for x in vs]
    primers = _plain_table(["Direction"] + names(left.keys()), [["Forward"] + values(left.values()), ["Reverse"] + values(right.values())])
    pairs = _plain_table(names(pair.keys()), [values(pair.values())])
    return "\nPrimer statistics:\n" + primers + "\n\nPair statistics:\n" + pairs


def annotate_alignment(block, p3, dot=False):
    """Add the primer positions and the statistics to one rendered alignment block (Amplicon.py:640-660): in bracket mode the
    annotation is merged into the bracket line, in dot mode it is a line of its own."""
    result = block[:-1].split("\n") if block.endswith("\n") else block.split("\n")
    fwd, rev = p3["PRIMER_LEFT_0_SEQUENCE"], p3["PRIMER_RIGHT_0_SEQUENCE"]
    f0 = p3["PRIMER_LEFT_0"][0]
    r0 = p3["PRIMER_RIGHT_0"][0] - p3["PRIMER_RIGHT_0"][1]
    text = (" " * f0 + "\u2514" + "Forward".center(len(fwd) - 2, "\u2500") + "\u2518" + " " * (r0 - f0 - len(fwd) + 1)
            + "\u2514" + "Reverse".center(len(rev) - 2, "\u2500") + "\u2518")
    if dot:
        result.append(text)
    else:
        last = result[-1].ljust(len(text))
        result[-1] = "".join(a if b == " " else b for b, a in zip(last, text))
    result.append(primer3_stats_text(p3))
    result[-1] += "\n"
    return "\n".join(result)


def _group_order(res):
    keys = [(res.left[i].tobytes(), res.right[i].tobytes()) for i in range(res.n_groups)]
    return sorted(range(res.n_groups), key=lambda i: keys[i])


def csv_rows(res, order=None):
    """Rows in output order (not canonically sorted; see SearchResult.rows for the sorted set)."""
    n = res.n_groups
    if n == 0:
        return []
    order = _group_order(res) if order is None else order
    cons = res.in_mask if res.have_outgroup else (res.in_mask | res.out_mask)
    out = []
    for i in order:
        out.append(f"{res.left[i].tobytes().decode()},{_IUPAC[cons[i]].tobytes().decode()},{res.right[i].tobytes().decode()}")
    return out


def group_amplicons(res, g, labels):
    """Distinct sequences of surviving group g with their label multisets: {mid: [label, ...]}.

    The run returned by the library may hold records of other flank keys (prefix sort), so records are
    matched on their flank bits first."""
    L, D, R = res.L, res.D, res.R
    a, b = int(res.run_offset[g]), int(res.run_offset[g + 1])
    recs = res.records[a:b]
    if recs.shape[0] == 0:
        raise ValueError("search was run without want_records")
    FB = 2 * (L + R)
    W = recs.shape[1]
    # flank bits of each record vs the group's flank
    fw = res.flank_words[g]
    keep = np.ones(recs.shape[0], dtype=bool)
    for j in range(len(fw)):
        nb = min(64, FB - 64 * j)
        if nb <= 0:
            break
        mask = np.uint64(0xFFFFFFFFFFFFFFFF) if nb == 64 else np.uint64(((1 << nb) - 1) << (64 - nb))
        keep &= (recs[:, j] & mask) == (fw[j] & mask)
    recs = recs[keep]
    mids = _decode_bases(recs, FB, D)
    ids = (recs[:, W - 1] & np.uint64(0xFF)).astype(np.int64)
    amps = {}
    for m, f in zip(mids, ids):
        amps.setdefault(m.tobytes().decode(), []).append(labels[int(f)])
    return amps


def _labels_to_string(labs):
    counts = {}
    for lab in labs:
        counts[lab] = counts.get(lab, 0) + 1
    return ";".join(name if c == 1 else f"{name}({c})" for name, c in sorted(counts.items()))


def render_alignment(left, right, amps, ingroup=None, dot=False):
    """One alignment block (render_alignment, Amplicon.py:598-661, without Primer3 annotations)."""
    L, D = len(left), len(next(iter(amps)))
    items = sorted(((sorted(labs), mid) for mid, labs in amps.items()), key=lambda x: (x[0], x[1]))
    lines_in, lines_out = [], []
    for labs, mid in items:
        text = f"{left}{mid}{right} : {_labels_to_string(labs)}"
        if ingroup is not None and not (set(labs) & ingroup):
            lines_out.append(text)
        else:
            lines_in.append(text)
    result = lines_in + lines_out
    if dot:
        top = result[0]
        new = [top]
        k = L + D + len(right)
        for seq in result[1:]:
            s = list(seq)
            for i in range(k):
                if top[i] == s[i]:
                    s[i] = "."
            new.append("".join(s))
        result = new
    else:
        bracket = list(" " * (L - 1) + "{" + "-" * D + "}")
        mids = list(amps)
        for c in range(D):
            if len({m[c] for m in mids}) > 1:
                bracket[L + c] = "*"
        if ingroup is not None:
            in_d, out_d = [], []
            for mid, labs in amps.items():
                for lab in labs:
                    (in_d if lab in ingroup else out_d).append(mid)
            for c in range(D):
                if {m[c] for m in in_d}.isdisjoint({m[c] for m in out_d}):
                    bracket[L + c] = "#"
        result.append("".join(bracket))
    result[-1] += "\n"
    return "\n".join(result)


def interchange_lines(res, labels, order=None):
    """The survivors in the reference's interchange text (Amplicon.write, Amplicon.py:330-348; label grammar :170-206): one line
    ``left,mid,right,label(n);label...`` per distinct sequence, groups in ascending (left, right) order — the content of the
    reference's ``filtered.txt`` (``merged_file.txt`` when D == 0; krisp_fasta.py:256-283), which its unchanged renderer
    (``render_output``, outputAlignments.py:101; ``--out_align``, ``--dot-alignment``, ``--primer3``) reads back.
    Inside a group the reference's line order follows its merge tree; here lines are sorted by middle (readers do not care:
    alignmentStream shared.py:442-475 collects the whole group).  Needs a search run with want_records."""
    order = _group_order(res) if order is None else order
    out = []
    for g in order:
        left, right = res.left[g].tobytes().decode(), res.right[g].tobytes().decode()
        amps = group_amplicons(res, g, labels)
        for mid in sorted(amps):
            out.append(f"{left},{mid},{right},{_labels_to_string(amps[mid])}")
    return out


def write_interchange(res, labels, filename):
    """Write interchange_lines to `filename`; returns the number of lines (what the reference's stage functions return)."""
    lines = interchange_lines(res, labels)
    with open(filename, "w") as fh:
        for ln in lines:
            fh.write(ln + "\n")
    return len(lines)


def render_output(res, labels, ingroup=None, out_csv=None, out_align=None, dot=False, find_primers=False, p3_args=None):
    """Write the CSV (stdout when out_csv is None) and, if asked, the alignment file.  Returns the number of regions written
    (render_output, outputAlignments.py:101-162).  find_primers: every region goes through Primer3 first, regions without a primer
    pair are dropped, the CSV gains the primer columns and the alignments the primer annotation (render_output_part :67-98)."""
    text = res.csv_rows_text()                   # rendered and ordered on the device (ascending (left, right))
    lines = text.splitlines()
    p3s = None
    if find_primers:
        p3s = []
        for ln in lines:
            left, cons, right = ln.split(",")
            p3 = run_primer3(left + cons + right, target_start=len(left), target_len=len(cons), **(p3_args or {}))
            p3s.append(p3 if p3["PRIMER_PAIR_NUM_RETURNED"] != 0 else None)
        lines = [ln + "," + ",".join(str(p3[n]) for n in P3_COLS) for ln, p3 in zip(lines, p3s) if p3 is not None]
        text = "".join(ln + "\n" for ln in lines)
    n_rows = len(lines)
    stream = sys.stdout if out_csv is None else open(out_csv, "w")
    try:
        stream.write(CSV_HEADER + ("," + ",".join(P3_KEYS) if find_primers else "") + "\n" + text)
    finally:
        if out_csv is not None:
            stream.close()
    if out_align is not None:
        order = _group_order(res)               # the same ascending (left, right) order as the rows
        if os.path.isfile(out_align):
            os.remove(out_align)
        ing = frozenset(ingroup) if ingroup is not None else None
        with open(out_align, "a") as fh:
            for i, g in enumerate(order):
                if p3s is not None and (i >= len(p3s) or p3s[i] is None):
                    continue
                amps = group_amplicons(res, g, labels)
                block = render_alignment(res.left[g].tobytes().decode(), res.right[g].tobytes().decode(), amps, ing, dot)
                if p3s is not None:
                    block = annotate_alignment(block, p3s[i], dot)
                print(block, file=fh)
    return n_rows


# ---- the renderer stage on an interchange FILE (host; survivors only) ---------------------------------------------------------------
_IUPAC_OF = {frozenset("A"): "A", frozenset("C"): "C", frozenset("G"): "G", frozenset("T"): "T", frozenset("AC"): "M", frozenset("AG"): "R",
             frozenset("AT"): "W", frozenset("CG"): "S", frozenset("CT"): "Y", frozenset("GT"): "K", frozenset("ACG"): "V", frozenset("ACT"): "H",
             frozenset("AGT"): "D", frozenset("CGT"): "B", frozenset("ACGT"): "N"}


def _collapse(seqs):
    """collapse_to_iupac (Amplicon.py:42-66) for equally long strings over ACGT (+ N / * / ? -> N)."""
    out = []
    for col in zip(*seqs):
        c = frozenset(col)
        out.append("N" if c & frozenset("N*?") else _IUPAC_OF[c])
    return "".join(out)


def read_interchange(kmerfile):
    """Groups of an interchange file (``left,mid,right,label(n);...`` lines, equal (left, right) adjacent: alignmentStream,
    shared.py:442-475): [((left, right), [(mid, [label, ...]), ...]), ...] in file order."""
    import re
    groups = []
    with open(kmerfile) as fh:
        for ln in fh:
            ln = ln.rstrip("\n")
            if not ln:
                continue
            left, mid, right, labs = ln.split(",")
            labels = []
            for item in labs.split(";"):
                m = re.fullmatch(r"(.+?)(?:\((\d+)\))?", item)
                labels.extend([m.group(1)] * int(m.group(2) or 1))
            if not groups or groups[-1][0] != (left, right):
                groups.append(((left, right), []))
            groups[-1][1].append((mid, labels))
    return groups


def render_output_file(kmerfile, out_align=None, out_csv=None, cores=1, print_block=10, ingroup=None, find_primers=False, dot=False, p3_args=None):
    """``render_output(kmerfile, ...)`` of the reference (outputAlignments.py:101-162) on an interchange file — e.g. the one
    ``write_interchange`` wrote, or the reference's own ``filtered.txt``: CSV rows (header first; stdout when out_csv is None), the
    alignment file when asked for, Primer3 post-filter when asked for.  Returns the number of regions written.  `cores` and
    `print_block` are accepted for compatibility (one process; the reference's row order with --cores 1).  `dot` stands for the
    reference's class-wide ``ConservedEndAmplicons.ENABLE_DOT`` switch (krisp_fasta.py:215)."""
    ing = frozenset(ingroup) if ingroup is not None else None
    rows, blocks = [], []
    for (left, right), amps in read_interchange(kmerfile):
        pick = amps
        if len(amps) > 1 and ing is not None:                          # render_csv, Amplicon.py:663-671: one amplicon -> its own sequence
            pick = [(m, labs) for m, labs in amps if set(labs) <= ing]
        cons = _collapse([m for m, _ in pick]) if pick and len(amps[0][0]) else ""
        row = f"{left},{cons},{right}"
        p3 = None
        if find_primers:
            p3 = run_primer3(left + cons + right, target_start=len(left), target_len=len(cons), **(p3_args or {}))
            if p3["PRIMER_PAIR_NUM_RETURNED"] == 0:
                continue
            row += "," + ",".join(str(p3[n]) for n in P3_COLS)
        rows.append(row)
        if out_align is not None:
            merged = {}
            for m, labs in amps:
                merged.setdefault(m, []).extend(labs)
            block = render_alignment(left, right, merged, ing, dot)
            blocks.append(annotate_alignment(block, p3, dot) if p3 is not None else block)
    stream = sys.stdout if out_csv is None else open(out_csv, "w")
    try:
        stream.write(CSV_HEADER + ("," + ",".join(P3_KEYS) if find_primers else "") + "\n" + "".join(r + "\n" for r in rows))
    finally:
        if out_csv is not None:
            stream.close()
    if out_align is not None:
        if os.path.isfile(out_align):
            os.remove(out_align)
        with open(out_align, "a") as fh:
            for b in blocks:
                print(b, file=fh)
    return len(rows)
