"""CPU: the per-chunk coder routines of the CUDA kernels (slimfastq_b200/csrc/*.cuh), compiled for
the host by tests/emul, versus the oracle.  Catches logic errors before any GPU time is spent."""
import os

import pytest

import emul
from conftest import sample_files
from helpers import check_container_against_oracle
from slimfastq_b200 import synth

SAMPLES = [p for p in sample_files() if os.path.getsize(p) < 1_000_000]


@pytest.mark.parametrize("name", sorted(synth.edge_cases()))
def test_edge_cases(oracle, name):
    data = synth.edge_cases()[name]
    for level in (1, 2, 3, 4):
        blob = emul.compress(data, level, 1 << 40)
        check_container_against_oracle(oracle, data, blob, level)
        assert emul.decompress(blob) == oracle.decode(oracle.encode(data, level))


@pytest.mark.parametrize("name", sorted(synth.oversized_cases()))
def test_oversized_records(oracle, name):
    """usr.lrec / usr.lgen / usr.lqlt through the planner, the header coder and the usr decoder (whole file and in chunks)."""
    data = synth.oversized_cases()[name]
    for level, chunk, two in ((3, 1 << 40, False), (1, 1 << 40, True), (3, 1 << 16, True)):
        blob = emul.compress(data, level, chunk, two_phase=two)
        check_container_against_oracle(oracle, data, blob, level)
        assert emul.decompress(blob) == data


def test_long_reads_beyond_64k_in_chunks(oracle):
    data = synth.ont(40, max_len=150000, mu=10.6)                # 15 of the 40 reads are oversized records; N runs in both kinds
    blob = emul.compress(data, 3, 1 << 20, two_phase=True)
    ct = check_container_against_oracle(oracle, data, blob, 3)
    assert len(ct.chunks) > 2 and sum("usr.lgen" in c.streams for c in ct.chunks) > 2
    assert emul.decompress(blob) == data


def test_synthetic_shapes_chunked(oracle):
    for data, level in ((synth.illumina(7000), 3), (synth.illumina(3000, bins8=True), 4), (synth.ont(40), 3), (synth.illumina(3000), 1)):
        blob = emul.compress(data, level, 1 << 19)
        ct = check_container_against_oracle(oracle, data, blob, level)
        assert len(ct.chunks) > 1
        assert emul.decompress(blob) == data


@pytest.mark.skipif(not SAMPLES, reason="oracle/_ref/samples not present")
@pytest.mark.parametrize("path", SAMPLES, ids=[os.path.basename(p) for p in SAMPLES])
def test_reference_samples(oracle, path):
    data = open(path, "rb").read()
    for level, chunk in ((3, 1 << 40), (2, 1 << 17)):
        blob = emul.compress(data, level, chunk)
        check_container_against_oracle(oracle, data, blob, level)
        assert emul.decompress(blob) == oracle.decode(oracle.encode(data, level)) or chunk != 1 << 40


def test_two_phase_encoder_algorithm(oracle):
    """Coding steps first (per-context replay models), coder chain second: same bytes as the oracle."""
    cases = [(synth.illumina(4000), 3), (synth.illumina(2500, bins8=True), 4), (synth.ont(30), 3), (synth.illumina(2500), 1),
             (synth.illumina(2500), 2)]
    cases += [(d, 3) for d in synth.edge_cases().values()]
    for p in SAMPLES:
        cases.append((open(p, "rb").read(), 3))
    for data, level in cases:
        blob = emul.compress(data, level, 1 << 19, two_phase=True)
        check_container_against_oracle(oracle, data, blob, level)


def test_parts_on_the_chunk_grid_give_the_one_call_container():
    """sfq_slot_target / sfq_set_chunk_phase: parts that begin at the first record after a grid line, coded
    with their phase and merged, are the whole-file container byte for byte (the rule the N-GPU sharding
    and the streaming CLI rely on).  Run through the CPU emulation of the planner and coders."""
    import emul
    from slimfastq_b200 import api, synth

    data = synth.illumina(5200) + synth.ont(25) + synth.illumina(1500, seed=9)
    for chunk in (1 << 17, 300_000):
        whole = emul.compress(data, 3, chunk)
        for nparts in (2, 3, 7):
            ranges = api.split_on_grid(data, nparts, chunk)
            assert ranges[0][0] == 0 and ranges[-1][1] == len(data)
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            parts = [emul.compress(data[s:e], 3, chunk, phase=ph) for s, e, ph in ranges]
            assert api.merge_containers(parts) == whole, (chunk, nparts)
        lens = [ln for s, e, ph in api.split_on_grid(data, 3, chunk) for ln in api.chunk_lengths(data[s:e], chunk, ph)]
        assert lens == api.chunk_lengths(data, chunk)


def test_stream_cuts_keep_the_chunk_grid():
    """The streaming CLI's loop (read a segment, sfq_stream_cut, code the part with its phase, carry the tail)
    replayed here with the CPU emulation as the coder: whatever the segment size, the appended blobs are the
    one-shot container."""
    import ctypes as C

    import emul
    from slimfastq_b200 import api, synth

    L = api.load_library()
    data = synth.illumina(4000) + synth.ont(20) + synth.illumina(2500, seed=3)
    chunk = 1 << 17
    whole = emul.compress(data, 3, chunk)
    for seg in (400_000, 1_000_003, len(data) + 5):
        parts, pos, carry, goff, phase = [], 0, b"", 0, 0
        while pos < len(data) or carry:
            take = max(0, seg - len(carry))
            buf = carry + data[pos:pos + take]
            pos += take
            eof = pos >= len(data)
            cut, nph = len(buf), C.c_uint64(0)
            if not eof:
                cut = L.sfq_stream_cut(C.cast(C.c_char_p(buf), C.c_void_p), len(buf), goff, chunk, C.byref(nph))
                if cut == 0:                      # no grid line with a record after it yet: read on
                    seg_more = data[pos:pos + seg]
                    carry, pos = buf + seg_more, pos + len(seg_more)
                    continue
            parts.append(emul.compress(buf[:cut], 3, chunk, phase=phase))
            carry, goff, phase = buf[cut:], goff + cut, nph.value
        assert api.merge_containers(parts) == whole, seg
        assert len(parts) > 1 or seg > len(data)
