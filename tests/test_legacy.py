"""Files of format versions below 5: their header stream has an older layout, which the reference
still reads (RecLoad::load_pre5, recs.cpp:397-398, 463-510; `config.level` missing means level 2, config.cpp:363).

The reference has no encoder for that layout any more, so the test input comes from oracle/sfq_oracle.c's
`sfq_oracle_encode_pre5` (test infrastructure) and is pinned the other way round: the UNMODIFIED reference binary must
decode it to the original text.  Then the product's decoder - CPU emulation of the kernel routine here, the kernel itself
under -m gpu - must print the same bytes."""
import os
import subprocess
import tempfile

import pytest

import emul
import slimfastq_b200 as S
from conftest import sample_files
from helpers import container_from_oracle
from oracle import oracle as O
from slimfastq_b200 import container as K
from slimfastq_b200 import synth

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/slimfastq is not built")


def cases():
    out = {"illumina": synth.illumina(3000), "headers": synth.edge_cases()["headers"], "twoid_varlen": synth.edge_cases()["twoid_varlen"],
           "solid": synth.edge_cases()["solid"], "ont": synth.ont(25)}
    for p in sample_files():
        if os.path.basename(p) in ("tst3.fq", "tstb.fq", "badsprintf.fq", "fast5.to.fq"):
            out[os.path.basename(p)] = open(p, "rb").read()
    return out


def legacy_reference_file(data: bytes, level: int, version_line: bytes) -> bytes:
    """A reference-format file whose info stream says `version_line` (same length as b"version=6") and whose `rec` stream
    is in the pre-v5 layout."""
    enc = O.encode(data, level, pre5=True)
    ref_file = bytearray(S.export_reference(container_from_oracle(enc, len(data)), "in.fq"))
    at = ref_file.index(b"version=6\n", 0, 8192)
    ref_file[at:at + 9] = version_line
    return bytes(ref_file)


def ref_decompress(sfq: bytes) -> bytes:
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "x.sfq"), os.path.join(d, "x.fq")
        open(src, "wb").write(sfq)
        subprocess.run([O.REF_BIN, "-d", "-f", src, "-u", dst, "-O"], check=True)
        return open(dst, "rb").read()


@needs_ref
@pytest.mark.parametrize("name", sorted(cases()))
def test_reference_binary_reads_the_legacy_stream_and_so_do_we(name):
    data = cases()[name]
    for level, vline in ((3, b"version=4"), (1, b"versiom=6")):           # an absent `version` key reads 0 (config.cpp:373)
        legacy = legacy_reference_file(data, level, vline)
        want = ref_decompress(legacy)                                     # the reference takes load_pre5 for this file
        assert want == O.decode(O.encode(data, level))                    # ... and prints what it prints for a current file
        blob = S.import_reference(legacy)
        ch = K.parse(blob).chunks[0]
        assert K.BLOB_HDR.unpack_from(blob, K.FILE_HDR.size)[12] & 2, "import must flag the pre-v5 header stream"
        assert ch.streams["rec"] == O.encode(data, level, pre5=True).streams["rec"]
        assert O.decode(O.encode(data, level, pre5=True)) == want          # the oracle's own legacy decoder
        assert emul.decompress(blob) == want                              # the kernel routine, on the CPU


@needs_ref
def test_missing_config_level_means_level_2():
    data = synth.illumina(1500)
    enc = O.encode(data, 2, pre5=True)
    ref_file = bytearray(S.export_reference(container_from_oracle(enc, len(data)), "in.fq"))
    at = ref_file.index(b"config.level=2\n", 0, 8192)
    ref_file[at:at + 14] = b"config.lewel=2"
    at = ref_file.index(b"version=6\n", 0, 8192)
    ref_file[at:at + 9] = b"version=3"
    want = ref_decompress(bytes(ref_file))
    assert want == data
    blob = S.import_reference(bytes(ref_file))
    assert K.parse(blob).chunks[0].level == 2
    assert emul.decompress(blob) == data


@pytest.mark.gpu
@needs_ref
def test_gpu_decodes_legacy_files(codec):
    for name, data in cases().items():
        legacy = legacy_reference_file(data, 3, b"version=4")
        assert codec.decompress(S.import_reference(legacy)) == ref_decompress(legacy), name
