"""compute-sanitizer target: edge cases x levels through the C ABI (round trip)."""
import sys
sys.path.insert(0, ".")
import slimfastq_b200 as S
from slimfastq_b200 import synth
c = S.Codec(0)
cases = dict(synth.edge_cases())
cases["illumina"] = synth.illumina(3000)
cases["ont"] = synth.ont(20)
for name, data in cases.items():
    for level in (1, 3):
        blob = c.compress(data, level, 1 << 19)
        back = c.decompress(blob)
        print(name, level, len(data), len(blob), back == data, flush=True)
