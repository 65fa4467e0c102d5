"""CPU: pins the oracle (oracle/sfq_oracle.c) against the reference binary and the golden vectors."""
import hashlib
import json
import os

import pytest

from conftest import sample_files

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
SAMPLES = sample_files()


def golden_inputs():
    from golden.make_golden import inputs

    return inputs()


@pytest.fixture(scope="module")
def gold_inputs():
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    return golden_inputs()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_golden_reference_vectors(oracle, gold_inputs, name):
    data = gold_inputs[name]
    g = GOLD[name]
    assert hashlib.md5(data).hexdigest() == g["input_md5"], "synthetic generator drifted; regenerate golden.json"
    for level, ge in g["levels"].items():
        enc = oracle.encode(data, int(level))
        info = ge["info"]
        assert (enc.llen, enc.solid, enc.two_id, enc.n_byte, enc.num_records, enc.rec_first.decode("latin1")) == \
               (info["llen"], info["solid"], info["two_id"], info["n_byte"], info["num_records"], info["rec_first"])
        assert {k: [len(v), hashlib.md5(v).hexdigest()] for k, v in enc.streams.items()} == ge["streams"]
        assert hashlib.md5(oracle.decode(enc)).hexdigest() == ge["decoded_md5"]


@pytest.mark.skipif(not SAMPLES, reason="oracle/_ref/samples not present (built from /root/reference)")
@pytest.mark.parametrize("path", SAMPLES, ids=[os.path.basename(p) for p in SAMPLES])
def test_oracle_equals_reference_binary_on_its_samples(oracle, path):
    if not oracle.have_ref():
        pytest.skip("reference binary not built")
    data = open(path, "rb").read()
    for level in (1, 2, 3, 4):
        mine, ref = oracle.encode(data, level), oracle.ref_encode(data, level)
        assert mine.info_tuple() == ref.info_tuple()
        assert mine.streams == ref.streams
        if level in (1, 3):
            assert oracle.decode(mine) == oracle.ref_roundtrip(data, level)


@pytest.mark.parametrize("name", sorted(__import__("slimfastq_b200.synth", fromlist=["x"]).oversized_cases()))
def test_oracle_equals_reference_binary_on_oversized_records(oracle, name):
    """usr.lrec / usr.lgen / usr.lqlt (usrs.cpp:269-301, 473-485): no reference sample has such records, so the pin
    is the reference binary run here on synthetic ones (70 kb reads, 8.5 KiB ids, the exact limits, leading and
    quality-only oversize, SOLiD, a file of nothing else)."""
    if not oracle.have_ref():
        pytest.skip("reference binary not built")
    from slimfastq_b200 import synth

    data = synth.oversized_cases()[name]
    for level in (1, 3):
        mine, ref = oracle.encode(data, level), oracle.ref_encode(data, level)
        assert "usr.lrec" in ref.streams and "usr.lgen" in ref.streams and "usr.lqlt" in ref.streams
        assert mine.info_tuple() == ref.info_tuple()
        assert mine.streams == ref.streams
        assert oracle.decode(mine) == oracle.ref_roundtrip(data, level) == data


def test_oracle_rejects_what_the_reference_croaks_on(oracle):
    for bad in (b"", b"hello\n", b"@r1\nACGT\n-\nIIII\n", b"@r1\nACXT\n+\nIIII\n", b"@r1\nACGT\n+\nIIII"):
        with pytest.raises(oracle.OracleError):
            oracle.encode(bad, 3)


def test_extractor_reads_reference_container(oracle):
    if not oracle.have_ref():
        pytest.skip("reference binary not built")
    data = golden_inputs()["twoid_varlen"]
    enc = oracle.ref_encode(data, 3)
    assert enc.two_id == 1 and enc.num_records == 200 and "usr.x" in enc.streams
    assert all(s[:4] == b"\0\0\0\0" for s in enc.streams.values())       # coder.hpp:34-39: 32-bit range under 64-bit low
