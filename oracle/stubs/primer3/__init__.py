"""Import stub: primer3-py is only reached with --primer3 (Amplicon.py:143)."""


class bindings:
    @staticmethod
    def design_primers(*a, **k):
        raise RuntimeError("primer3-py is not installed (stub); --primer3 is unavailable")
