"""Interchange with the reference's own 8 KiB-page .sfq file.

CPU part: the two host-side format conversions, checked with the unmodified reference binary on its
own sample files (no coding happens in the conversions, so no GPU is needed).  GPU part: a file
compressed on the GPU is decoded by the reference binary, and a file compressed by the reference
binary is decoded on the GPU.
"""
import os
import subprocess
import tempfile

import pytest

import slimfastq_b200 as S
from conftest import sample_files
from oracle import oracle as O
from oracle import sfq_extract
from slimfastq_b200 import container as K
from slimfastq_b200 import synth

SAMPLES = sample_files()
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/slimfastq is not built")


def ref_compress(data: bytes, level: int, d: str) -> bytes:
    src, dst = os.path.join(d, "in.fq"), os.path.join(d, "out.sfq")
    open(src, "wb").write(data)
    subprocess.run([O.REF_BIN, "-u", src, "-f", dst, "-O", "-q", "-l", str(level)], check=True)
    return open(dst, "rb").read()


def ref_decompress(sfq: bytes, d: str) -> bytes:
    src, dst = os.path.join(d, "x.sfq"), os.path.join(d, "x.fq")
    open(src, "wb").write(sfq)
    subprocess.run([O.REF_BIN, "-d", "-f", src, "-u", dst, "-O"], check=True)
    return open(dst, "rb").read()


@needs_ref
@pytest.mark.parametrize("path", SAMPLES, ids=[os.path.basename(p) for p in SAMPLES])
def test_format_conversions_with_the_reference_binary(path):
    data = open(path, "rb").read()
    with tempfile.TemporaryDirectory() as d:
        ref = ref_compress(data, 3, d)
        assert S.is_reference_file(ref)
        info, streams = sfq_extract.extract(ref)
        blob = S.import_reference(ref)                            # reference file -> single-chunk container
        assert not S.is_reference_file(blob)
        ct = K.parse(blob)
        assert len(ct.chunks) == 1
        ch = ct.chunks[0]
        assert ch.streams == {k: v for k, v in streams.items() if v}
        assert ch.num_records == int(info["num_records"]) and ch.level == 3
        assert ch.rec_first == info.get("rec.first", "").encode("latin1")
        back = S.export_reference(blob, "in.fq")                  # ... and back to a reference file
        info2, streams2 = sfq_extract.extract(back)
        assert streams2 == streams
        for k in ("config.level", "llen", "usr.2id", "usr.solid", "rec.first", "gen.N_byte", "num_records", "orig.size"):
            assert info2.get(k) == info.get(k), k
        assert int(info2["comp.size"]) == len(back)
        assert ref_decompress(back, d) == ref_decompress(ref, d)  # the reference decodes our pages like its own


@needs_ref
@pytest.mark.parametrize("name", sorted(synth.oversized_cases()))
def test_format_conversions_keep_oversized_record_streams(name):
    data = synth.oversized_cases()[name]
    with tempfile.TemporaryDirectory() as d:
        ref = ref_compress(data, 3, d)
        info, streams = sfq_extract.extract(ref)
        assert all(k in streams for k in ("usr.lrec", "usr.lgen", "usr.lqlt"))
        blob = S.import_reference(ref)
        assert K.parse(blob).chunks[0].streams == {k: v for k, v in streams.items() if v}
        back = S.export_reference(blob, "in.fq")
        assert sfq_extract.extract(back)[1] == streams
        assert ref_decompress(back, d) == data


@needs_ref
def test_multi_page_streams_and_node_pages():
    # > 2047 data pages in one stream (a 17 MiB `gen` stream: 70 M random bases) exercises a second node page
    import numpy as np

    rng = np.random.default_rng(7)
    L, n = 10000, 7000
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, L))]
    qual = b"I" * L
    data = b"".join(b"@r%d\n" % i + bases[i].tobytes() + b"\n+\n" + qual + b"\n" for i in range(n))
    with tempfile.TemporaryDirectory() as d:
        ref = ref_compress(data, 1, d)
        blob = S.import_reference(ref)
        assert max(len(v) for v in K.parse(blob).chunks[0].streams.values()) > 2047 * 8192
        back = S.export_reference(blob)
        assert sfq_extract.extract(back)[1] == sfq_extract.extract(ref)[1]
        assert ref_decompress(back, d) == data


def test_conversions_reject_what_they_cannot_represent():
    with pytest.raises(S.SfqError):
        S.import_reference(b"@r1\nACGT\n+\nIIII\n" * 2000)
    with pytest.raises(S.SfqError):
        S.export_reference(b"whoami=slimfastq" + b"\0" * 20000)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("level", [1, 3])
def test_gpu_output_is_decoded_by_the_reference_binary(codec, level):
    cases = [open(p, "rb").read() for p in SAMPLES[:6]] + [synth.illumina(5000), synth.edge_cases()["badqlt"]] + list(synth.oversized_cases().values())
    with tempfile.TemporaryDirectory() as d:
        for data in cases:
            blob = codec.compress(data, level, 1 << 40)           # one chunk = one reference file
            ref_file = S.export_reference(blob, "in.fq")
            assert ref_decompress(ref_file, d) == O.ref_roundtrip(data, level)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("level", [2, 4])
def test_reference_output_is_decoded_on_the_gpu(codec, level):
    cases = [open(p, "rb").read() for p in SAMPLES] + [synth.illumina(5000), synth.ont(40), synth.edge_cases()["badqlt"]] + list(synth.oversized_cases().values())
    with tempfile.TemporaryDirectory() as d:
        for data in cases:
            ref = ref_compress(data, level, d)
            assert codec.decompress(S.import_reference(ref)) == ref_decompress(ref, d)
