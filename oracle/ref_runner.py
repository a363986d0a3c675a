"""Run the UNMODIFIED reference (grunwaldlab/krisp at /root/reference) as a live oracle.

TEST INFRASTRUCTURE ONLY, and only usable in the build container where
/root/reference is mounted (it does not exist on the GPU box).  It is used by
tests/golden/make_golden.py to produce the committed golden vectors and by the
optional ``-m "not gpu"`` differential tests that skip when the reference is absent.

The reference imports four third-party modules that are not installed here;
``oracle/stubs`` supplies import stubs.  None of their
arithmetic is on the hot path except Biopython's IUPAC table, reproduced in
``stubs/Bio/Data/IUPACData.py``.
"""
import os
import subprocess
import sys
import tempfile

REFERENCE_SRC = os.environ.get("KRISP_REFERENCE_SRC", "/root/reference/src")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "krisp", "krisp_fasta"))


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = _STUBS + os.pathsep + REFERENCE_SRC
    env["PYTHONWARNINGS"] = "ignore"
    return env


def krisp_fasta(argv, cwd=None, timeout=3600):
    """``krisp_fasta <argv>`` -> (stdout, stderr).  A --workdir is added if absent."""
    argv = list(map(str, argv))
    with tempfile.TemporaryDirectory() as wd:
        if "--workdir" not in argv and "-w" not in argv:
            argv += ["--workdir", wd]
        p = subprocess.run([sys.executable, "-c",
                            "from krisp.krisp_fasta.krisp_fasta import main; main()"] + argv,
                           cwd=cwd, env=_env(), capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference krisp_fasta failed rc={p.returncode}\n{p.stderr}")
    return p.stdout, p.stderr


def kstream(argv, cwd=None, stdin=None, timeout=3600):
    """``kstream <argv>`` -> stdout."""
    p = subprocess.run([sys.executable, "-m", "krisp.kstream.kstream"] + list(map(str, argv)),
                       cwd=cwd, env=_env(), capture_output=True, text=True, input=stdin, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference kstream failed rc={p.returncode}\n{p.stderr}")
    return p.stdout


def rows_of(stdout):
    """CSV rows without the header, canonically sorted (output is a *set* of rows)."""
    lines = [ln for ln in stdout.splitlines() if ln]
    assert lines and lines[0].startswith("left_seq,diag_seq,right_seq"), lines[:1]
    return sorted(lines[1:])


_SPY = r"""
import shutil, sys
import krisp.krisp_fasta.krisp_fasta as m
_dest = sys.argv.pop(1)
_real = m.render_output
def _spy(kmerfile, *a, **k):
    shutil.copyfile(kmerfile, _dest)          # observe the interchange file; the reference code itself is untouched
    return _real(kmerfile, *a, **k)
m.render_output = _spy
m.main()
"""


def krisp_fasta_interchange(argv, cwd=None, timeout=3600):
    """``krisp_fasta <argv>`` -> (stdout, text of the k-mer file handed to render_output).

    That file (``filtered.txt``, or ``merged_file.txt`` when D == 0; krisp_fasta.py:256-283) is the reference's
    interchange format ``left,mid,right,label(n);label...`` restricted to the surviving groups."""
    argv = list(map(str, argv))
    with tempfile.TemporaryDirectory() as wd:
        dest = os.path.join(wd, "interchange.txt")
        if "--workdir" not in argv and "-w" not in argv:
            argv += ["--workdir", wd]
        p = subprocess.run([sys.executable, "-c", _SPY, dest] + argv, cwd=cwd, env=_env(), capture_output=True, text=True,
                           timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError(f"reference krisp_fasta failed rc={p.returncode}\n{p.stderr}")
        with open(dest) as fh:
            text = fh.read()
    return p.stdout, text
