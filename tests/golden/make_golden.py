"""Regenerates tests/golden/golden.json: stream md5s and info keys that the UNMODIFIED reference
binary (oracle/_ref/slimfastq, built by oracle/Makefile from /root/reference) produces for inputs
that slimfastq_b200.synth regenerates deterministically.  Run in the dev container:

    python tests/golden/make_golden.py

The JSON travels with the repo, so the GPU box can check parity without /root/reference.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from slimfastq_b200 import synth  # noqa: E402


def inputs() -> dict[str, bytes]:
    d = dict(synth.edge_cases())
    d["illumina_6000"] = synth.illumina(6000)
    d["illumina8_3000"] = synth.illumina(3000, bins8=True)
    d["ont_60"] = synth.ont(60)
    return d


def main():
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    gold = {}
    for name, data in inputs().items():
        entry = {"input_md5": hashlib.md5(data).hexdigest(), "input_bytes": len(data), "levels": {}}
        for level in (1, 2, 3, 4):
            enc = O.ref_encode(data, level)
            entry["levels"][str(level)] = {
                "info": {"llen": enc.llen, "solid": enc.solid, "two_id": enc.two_id, "n_byte": enc.n_byte,
                         "num_records": enc.num_records, "rec_first": enc.rec_first.decode("latin1")},
                "streams": {k: [len(v), hashlib.md5(v).hexdigest()] for k, v in sorted(enc.streams.items())},
                "decoded_md5": hashlib.md5(O.ref_roundtrip(data, level)).hexdigest(),
            }
        gold[name] = entry
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", len(gold), "inputs")


if __name__ == "__main__":
    main()
