"""Golden-vector helpers shared by the tests."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def load_golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


def deduce_ldr(flags):
    """(L, D, R) exactly as krisp_fasta.py:178-213 deduces them from the CLI flags (k = amplicon)."""
    c, cl, cr = flags.get("conserved"), flags.get("conserved-left"), flags.get("conserved-right")
    d, a = flags.get("diagnostic"), flags.get("amplicon")
    if a is not None:
        if d is not None:
            cl = cr = (a - d) // 2
        elif c is not None:
            cl = cr = c
        elif cl is None or cr is None:
            raise ValueError("cannot deduce")
    elif d is not None:
        if c is not None:
            cl = cr = c
            a = d + 2 * c
        elif cl is not None and cr is not None:
            a = d + cl + cr
        else:
            raise ValueError("cannot deduce")
    else:
        raise ValueError("cannot deduce")
    return cl, a - cl - cr, cr


def golden_paths(case):
    return ([os.path.join(GOLDEN_DIR, p) for p in case["ingroup"]],
            [os.path.join(GOLDEN_DIR, p) for p in case["outgroup"]])
