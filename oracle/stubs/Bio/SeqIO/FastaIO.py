"""Stub of Bio.SeqIO.FastaIO.SimpleFastaParser (used at utils.py:60,190,574).

Semantics restated from Biopython's documentation: yields (title, sequence) where title is
the header line without '>' and without the trailing newline, and sequence is the
concatenation of the following lines with whitespace at line ends removed and *no*
separator inserted between lines.
"""


def SimpleFastaParser(handle):
    title = None
    lines = []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield title, "".join(lines).replace(" ", "").replace("\r", "")
            title = line[1:].rstrip()
            lines = []
        elif title is not None:
            lines.append(line.rstrip())
    if title is not None:
        yield title, "".join(lines).replace(" ", "").replace("\r", "")
