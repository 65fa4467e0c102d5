"""-m gpu: the kernel shapes the 10 GB benchmark actually runs (they are chosen from the wave size: from 4 096 chunks per
wave the quality decoder uses 4 lanes per chunk and the thread-per-chunk coders 16 chunk-streams per warp), forced on
a small input through the same environment knobs and checked against the oracle and the reference binary itself."""
import json
import os
import subprocess
import sys

import pytest

from helpers import check_container_against_oracle
from slimfastq_b200 import container as K
from slimfastq_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def codec_with(env: dict):
    import slimfastq_b200 as S

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return S.Codec()                 # the knobs are read when the context is created
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("env", [
    {"SFQ_QLPC": "4", "SFQ_LANES": "16", "SFQ_DEC_WARPS": "4", "SFQ_SPREAD": "3"},      # the >= 4 096-chunk shapes of bench.py
    {"SFQ_QLPC": "8", "SFQ_LANES": "8"},
    {"SFQ_QLPC": "4", "SFQ_LANES": "4", "SFQ_RC_LANES": "32", "SFQ_ENC_REC_LANES": "16"},
    {"SFQ_GM_TABLE": "1", "SFQ_QSCATTER": "1", "SFQ_ENC_PRIO": "0"},                    # round-1 encoder kernels (kept for A/B)
], ids=["bench-10GB", "mid-wave", "wide-chains", "round1-encoder"])
def test_benchmark_kernel_shapes_bit_exact(oracle, env):
    data = synth.illumina(24000)                     # ~8.6 MB -> nine 1 MiB chunks
    c = codec_with(env)
    try:
        for level in (3, 4, 1):
            blob = c.compress(data, level, 1 << 20)
            ct = check_container_against_oracle(oracle, data, blob, level)
            assert len(ct.chunks) >= 8
            assert c.decompress(blob) == data
        ont = synth.ont(60)
        blob = c.compress(ont, 3, 1 << 20)
        check_container_against_oracle(oracle, ont, blob, 3)
        assert c.decompress(blob) == ont
    finally:
        c.close()


def test_pipelined_parts_give_the_one_call_container(codec):
    """sfq_compress of a large host buffer codes it in parts cut on the chunk grid (copies beside the coding); forced here
    on a small one: the container must be the plain call's, byte for byte."""
    data = synth.illumina(9000)
    c = codec_with({"SFQ_PARTS": "3"})
    try:
        for level in (3, 1):
            blob = c.compress(data, level, 1 << 17)
            assert c.stats()["waves"] >= 3
            assert blob == codec.compress(data, level, 1 << 17)
            assert c.decompress(blob) == data
        ont = synth.ont(80)
        assert c.compress(ont, 3, 1 << 18) == codec.compress(ont, 3, 1 << 18)
    finally:
        c.close()
    # a buffer that is itself a shard of a file (sfq_set_chunk_phase) cut into parts: the shards' containers must still merge
    # into the one-call container (the parts' grid lines lie at k*B - phase)
    from slimfastq_b200 import api
    c = codec_with({"SFQ_PARTS": "2"})
    try:
        B = 1 << 17
        shards = api.split_on_grid(data, 2, B)
        assert len(shards) == 2 and shards[1][2] != 0
        blobs = [c.compress(data[a:b], 3, B, phase=ph) for a, b, ph in shards]
        assert api.merge_containers(blobs) == codec.compress(data, 3, B)
    finally:
        c.close()
    # the automatic split: a short head part (its copy-in is the exposed one) + one part per coder wave
    c = codec_with({"SFQ_PARTS": "-1", "SFQ_HEAD_FRAC": "0.3"})
    try:
        blob = c.compress(data, 3, 1 << 17)
        assert c.stats()["waves"] == 2
        assert blob == codec.compress(data, 3, 1 << 17) and c.decompress(blob) == data
    finally:
        c.close()


def test_chunks_equal_the_reference_binary_itself(codec, oracle):
    """Not the restatement: the unmodified reference run on each chunk as a standalone file."""
    if not oracle.have_ref():
        pytest.skip("reference binary not built")
    for data, level in ((synth.illumina(9000), 3), (synth.illumina(5000, bins8=True), 4), (synth.ont(40), 3), (synth.edge_cases()["solid"], 2)):
        ct = K.parse(codec.compress(data, level, 1 << 20))
        pos = 0
        for ch in ct.chunks:
            ref = oracle.ref_encode(data[pos:pos + ch.text_len], level)
            pos += ch.text_len
            assert ch.streams == ref.streams and ch.info_tuple() == ref.info_tuple()


def test_large_chunks_bit_exact(codec, oracle):
    """8 MiB chunks (the chunk size at which the ratio loss against the whole-file reference falls below 2 %)."""
    data = synth.illumina(50000)                     # ~18 MB -> three chunks
    blob = codec.compress(data, 3, 8 << 20)
    ct = check_container_against_oracle(oracle, data, blob, 3)
    assert len(ct.chunks) == 3
    assert codec.decompress(blob) == data


def test_single_file_over_two_gpus():
    """bench.py --single-file: one file cut on the chunk grid over two ranks, sizes all_gather'ed, blobs written at their
    offsets of one file; the result must be the 1-GPU container byte for byte (checked inside the run)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--single-file", "--gb", "0.5", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["check"]["merged_equals_single_gpu_container"] and line["check"]["merged_round_trip"]
