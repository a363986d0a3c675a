"""Import stub so the unmodified reference imports without colorama (test infrastructure only).
The reference uses only Fore.GREEN / Style.RESET_ALL in stderr messages (krisp_fasta.py:63)."""


class _Blank:
    def __getattr__(self, name):
        return ""


Fore = Back = Style = _Blank()
