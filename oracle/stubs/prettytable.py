"""Import stub: PrettyTable is only reached with --primer3 (Amplicon.py:566-595)."""


class PrettyTable:
    def __init__(self, *a, **k):
        raise RuntimeError("prettytable is not installed (stub); --primer3 is unavailable")
