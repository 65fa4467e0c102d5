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
