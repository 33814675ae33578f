"""End-to-end check of the scan_genome command line on a GPU box: exact motif strings against a
plain Python string search, both strands.  (r01: the B200 run wrote 175 sites = the case-insensitive
string search below; the first version of this check compared case-sensitively and expected 171.)"""
import os
import re
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from discover_tfbs_b200 import scan_genome, synth  # noqa: E402

d = tempfile.mkdtemp()
chroms = {"chrA": synth.genome(1, 0, 200_000).tobytes().decode(), "chrB": synth.genome(2, 0, 50_000).tobytes().decode()}
with open(f"{d}/g.fa", "w") as fh:
    for k, v in chroms.items():
        fh.write(f">{k}\n" + "\n".join(v[i:i + 60] for i in range(0, len(v), 60)) + "\n")
open(f"{d}/m.txt", "w").write("CACGTG\nTGACTCA\nGGGCGG\n")
assert scan_genome.main([f"{d}/g.fa", f"{d}/m.txt", f"{d}/out", "--exact"]) == 0
got = sorted(tuple(ln.split("\t")[i] for i in (0, 1, 2, 3, 5)) for ln in open(f"{d}/out.bed"))
comp = str.maketrans("ACGT", "TGCA")
want = []
for c, s in chroms.items():
    s = s.upper()  # the synthetic genome is soft-masked; matching is case-insensitive, like BLASTn's
    for w in ("CACGTG", "TGACTCA", "GGGCGG"):
        for strand, pat in (("+", w), ("-", w.translate(comp)[::-1])):
            want += [(c, str(m.start()), str(m.start() + len(w)), w, strand) for m in re.finditer(f"(?={pat})", s)]
assert got == sorted(want), (len(got), len(want))
print("cli ok:", len(got), "sites")
