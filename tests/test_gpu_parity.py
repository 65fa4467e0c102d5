"""-m gpu: the CUDA path, called through the C ABI, versus the oracle / the reference binary."""
import os

import pytest

from conftest import sample_files
from helpers import check_container_against_oracle
from slimfastq_b200 import synth

pytestmark = pytest.mark.gpu

SAMPLES = sample_files()


@pytest.mark.parametrize("path", SAMPLES, ids=[os.path.basename(p) for p in SAMPLES])
@pytest.mark.parametrize("level", [1, 2, 3, 4])
def test_reference_samples_bit_exact_and_roundtrip(codec, oracle, path, level):
    data = open(path, "rb").read()
    for chunk_bytes in (1 << 40, 1 << 17):          # whole file as one chunk, and 128 KiB chunks
        blob = codec.compress(data, level, chunk_bytes)
        check_container_against_oracle(oracle, data, blob, level)
        back = codec.decompress(blob)
        assert back == oracle.decode(oracle.encode(data, level)) or chunk_bytes != 1 << 40
        if oracle.have_ref() and chunk_bytes == 1 << 40:
            assert back == oracle.ref_roundtrip(data, level)       # what the reference itself prints
        else:
            assert back == data


@pytest.mark.parametrize("name", sorted(synth.edge_cases()))
@pytest.mark.parametrize("level", [1, 3, 4])
def test_edge_cases_bit_exact(codec, oracle, name, level):
    data = synth.edge_cases()[name]
    blob = codec.compress(data, level, 1 << 40)
    check_container_against_oracle(oracle, data, blob, level)
    assert codec.decompress(blob) == oracle.decode(oracle.encode(data, level))


@pytest.mark.parametrize("name", sorted(synth.oversized_cases()))
def test_oversized_records_bit_exact(codec, oracle, name):
    """Records the reference stores verbatim in usr.lrec / usr.lgen / usr.lqlt (id of 8 191+ characters, line of 65 535+)."""
    data = synth.oversized_cases()[name]
    for level, chunk in ((3, 1 << 40), (1, 1 << 16), (4, 1 << 18)):
        blob = codec.compress(data, level, chunk)
        check_container_against_oracle(oracle, data, blob, level)
        assert codec.decompress(blob) == data
    if oracle.have_ref():
        from slimfastq_b200 import container as K

        ref = oracle.ref_encode(data, 3)                          # the binary itself, not its restatement
        ch = K.parse(codec.compress(data, 3, 1 << 40)).chunks[0]
        assert ch.streams == ref.streams and ch.info_tuple() == ref.info_tuple()


def test_long_read_workload_with_reads_beyond_64k(codec, oracle):
    data = synth.ont(40, max_len=150000, mu=10.6)
    assert max(len(l) for l in data.split(b"\n")[1::4]) > 65535
    blob = codec.compress(data, 3, 1 << 20)
    check_container_against_oracle(oracle, data, blob, 3)
    assert codec.decompress(blob) == data


@pytest.mark.parametrize("level", [1, 2, 3, 4])
def test_illumina_chunks_bit_exact(codec, oracle, level):
    data = synth.illumina(12000)                     # ~4.3 MB -> five 1 MiB chunks
    blob = codec.compress(data, level, 1 << 20)
    ct = check_container_against_oracle(oracle, data, blob, level)
    assert len(ct.chunks) >= 4
    assert codec.decompress(blob) == data


def test_illumina_8bin_and_ont_bit_exact(codec, oracle):
    for data in (synth.illumina(6000, bins8=True), synth.ont(120)):
        for level in (3, 4):
            blob = codec.compress(data, level, 1 << 20)
            check_container_against_oracle(oracle, data, blob, level)
            assert codec.decompress(blob) == data


def test_waves_do_not_change_the_bytes(oracle):
    import slimfastq_b200 as S

    data = synth.illumina(9000)
    a = S.Codec(max_resident=2)
    b = S.Codec()
    try:
        blob_a = a.compress(data, 3, 1 << 19)
        assert a.stats()["waves"] > 1
        assert blob_a == b.compress(data, 3, 1 << 19)
        assert a.decompress(blob_a) == data
    finally:
        a.close(); b.close()


def test_errors_are_croaks_not_crashes(codec):
    import slimfastq_b200 as S

    for bad in (b"hello\nworld\n", b"@r1\nACGT\n+\nIIII", b"@r1\nACGT\n-\nIIII\n", b"@r1\nACXT\n+\nIIII\n"):
        with pytest.raises(S.SfqError):
            codec.compress(bad, 3)
    with pytest.raises(S.SfqError):
        codec.decompress(b"whoami=slimfastq\nversion=6\n" + b"\0" * 200)
    good = codec.compress(b"@r1\nACGT\n+\nIIII\n", 3)
    assert codec.decompress(good) == b"@r1\nACGT\n+\nIIII\n"


def test_decoder_survives_wrong_table_hints(codec):
    """The blob header's context counts only size the decoder's hash tables: with hints that are far too
    small the tables fill up, the library reruns the wave with larger ones (SFQ_E_TABLE), same output."""
    import struct

    from slimfastq_b200 import container as K

    data = synth.illumina(7000)
    blob = bytearray(codec.compress(data, 3, 1 << 20))
    _, _, _, _, _, nchunks, _, index_off, _ = K.FILE_HDR.unpack_from(blob, 0)
    hint_off = K.BLOB_HDR.size - 4 * len(K.STREAM_NAMES) - 16 - 8      # q_used, g_used sit before the four oversized-record counts and ssize[]
    for off in struct.unpack_from(f"<{nchunks}Q", blob, index_off):
        q_used, g_used = struct.unpack_from("<II", blob, off + hint_off)
        assert q_used > 1000 and g_used > 100000
        struct.pack_into("<II", blob, off + hint_off, 40, 3000)
    assert codec.decompress(bytes(blob)) == data
    assert codec.stats()["retries"] >= 1
    for off in struct.unpack_from(f"<{nchunks}Q", blob, index_off):
        struct.pack_into("<II", blob, off + hint_off, 0, 0)     # hints absent (e.g. an older writer)
    assert codec.decompress(bytes(blob)) == data


def test_alternating_sizes_reuse_the_scratch_arena(codec, oracle):
    """compress / decompress carve one arena differently; sizes going up and down must not leak state."""
    big, small = synth.illumina(9000), synth.illumina(700, seed=5)
    for data in (big, small, big, small):
        blob = codec.compress(data, 3, 1 << 19)
        check_container_against_oracle(oracle, data, blob, 3)
        assert codec.decompress(blob) == data
