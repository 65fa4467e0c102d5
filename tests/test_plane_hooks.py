"""-m gpu: the per-plane C-ABI hooks (include/sfq_b200.h): one plane's kernels at a time, against
the oracle's streams for that plane and against the input's own lines."""
import pytest

from slimfastq_b200 import container as K
from slimfastq_b200 import synth

pytestmark = pytest.mark.gpu

PLANE_STREAMS = {"gen": {"gen", "gen.Ns", "gen.Nn"}, "qlt": {"qlt"},
                 "rec": {"rec", "rec.x", "usr.x", "usr.x.q", "usr.pfg", "usr.pfq", "usr.lrec", "usr.lgen", "usr.lqlt"}}


def cases():
    return {"illumina": synth.illumina(5000), "ont": synth.ont(30), "varlen": synth.edge_cases()["twoid_varlen"],
            "badqlt": synth.edge_cases()["badqlt"], "oversize": synth.oversized_cases()["oversize_mid"]}


@pytest.mark.parametrize("name", sorted(cases()))
@pytest.mark.parametrize("plane", ["gen", "qlt", "rec"])
def test_encode_hooks_give_the_oracles_streams_of_that_plane(codec, oracle, name, plane):
    data = cases()[name]
    ct = K.parse(codec.encode_plane(plane, data, 3, 1 << 19))
    pos = 0
    for ch in ct.chunks:
        o = oracle.encode(data[pos:pos + ch.text_len], 3)
        pos += ch.text_len
        assert set(ch.streams) <= PLANE_STREAMS[plane]
        assert ch.streams == {k: v for k, v in o.streams.items() if k in PLANE_STREAMS[plane]}
    assert pos == len(data)


@pytest.mark.parametrize("name", ["illumina", "ont", "varlen", "badqlt"])
def test_decode_hooks_give_the_planes_lines(codec, name):
    data = cases()[name]
    blob = codec.compress(data, 3, 1 << 19)
    lines = data.split(b"\n")
    ids, seqs, quals = lines[0::4], lines[1::4], lines[3::4]
    assert codec.decode_plane("qlt", blob) == b"".join(q + b"\n" for q in quals if True)
    assert codec.decode_plane("rec", blob) == b"".join(h[1:] + b"\n" for h in ids if h)
    want = bytearray()
    for s, q in zip(seqs, quals):
        # an N under a '!' quality is coded as base 0 and restored by the rule the gen plane alone cannot apply (gens.cpp:200-213)
        want += bytes(ord("A") if (b == ord("N") and k < len(q) and q[k] == ord("!")) else b for k, b in enumerate(s)) + b"\n"
    assert codec.decode_plane("gen", blob) == bytes(want)
    assert codec.decompress(blob) == data           # the context is back to the whole path afterwards
