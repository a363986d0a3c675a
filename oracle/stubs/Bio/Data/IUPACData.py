"""Import stub carrying the one constant table of Biopython the reference reads
(Amplicon.py:10-11 inverts it; insertion order matters: N after X)."""
ambiguous_dna_values = {
    "A": "A", "C": "C", "G": "G", "T": "T",
    "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT", "K": "GT",
    "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT",
    "X": "GATC", "N": "GATC",
}
