"""Stub: the one Biopython name the reference uses from Bio.Seq (utils.py:104).

`ambiguous_dna_complement` is the IUPAC complement table published in
Bio/Data/IUPACData.py (upper-case keys only, so lower-case input raises KeyError in the
reference exactly as it does with real Biopython).
"""
ambiguous_dna_complement = {
    "A": "T", "C": "G", "G": "C", "T": "A",
    "M": "K", "R": "Y", "W": "W", "S": "S", "Y": "R", "K": "M",
    "V": "B", "H": "D", "D": "H", "B": "V",
    "X": "X", "N": "N",
}
