"""File name -> label, as the reference derives it (krisp_fasta/shared.py:34-73)."""
from pathlib import Path

_OMIT_EXT = ("gz", "bz2", "fna", "fasta", "fa", "ffn", "frn")


def basename(filename):
    """shared.py:34-55: strip directories and trailing sequence / compression extensions."""
    parts = Path(filename).name.split(".")
    while parts[-1] in _OMIT_EXT:
        parts.pop()
    return ".".join(parts)


def simplename(filename):
    """shared.py:58-73: the text before the first '.' of basename()."""
    return basename(filename).split(".")[0]
