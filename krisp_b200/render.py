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


def render_output(res, labels, ingroup=None, out_csv=None, out_align=None, dot=False):
    """Write the CSV (stdout when out_csv is None) and, if asked, the alignment file.  Returns the number of regions."""
    text = res.csv_rows_text()                   # rendered and ordered on the device (ascending (left, right))
    n_rows = text.count("\n")
    stream = sys.stdout if out_csv is None else open(out_csv, "w")
    try:
        stream.write(CSV_HEADER + "\n" + text)
    finally:
        if out_csv is not None:
            stream.close()
    if out_align is not None:
        order = _group_order(res)
        if os.path.isfile(out_align):
            os.remove(out_align)
        ing = frozenset(ingroup) if ingroup is not None else None
        with open(out_align, "a") as fh:
            for g in order:
                amps = group_amplicons(res, g, labels)
                print(render_alignment(res.left[g].tobytes().decode(), res.right[g].tobytes().decode(), amps, ing, dot), file=fh)
    return n_rows
