"""CPU: host-side bookkeeping - record-boundary search, chunk lengths, container merge."""
import emul
from slimfastq_b200 import api, container as K, synth


def test_record_start_search_handles_at_sign_qualities():
    fq = b"@r1\nACGT\n+\n@III\n@r2\nACGT\n+\n@@@@\n@r3\nAC\n+\nII\n"
    starts = [0, 16, 32]
    for pos in range(len(fq) + 1):
        want = next((s for s in starts if s >= pos), len(fq))
        assert api.record_start_at_or_after(fq, pos) == want, pos


def test_chunk_lengths_match_device_planner():
    data = synth.illumina(5000)
    for cb in (1 << 18, 1 << 20):
        ct = K.parse(emul.compress(data, 1, cb))
        assert api.chunk_lengths(data, cb) == [c.text_len for c in ct.chunks]


def test_split_and_merge_roundtrip():
    data = synth.illumina(4000)
    parts = api.split_records(data, 3)
    assert parts[0][0] == 0 and parts[-1][1] == len(data) and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    blobs = [emul.compress(data[a:b], 3, 1 << 18) for a, b in parts]
    merged = api.merge_containers(blobs)
    ct = K.parse(merged)
    assert ct.orig_size == len(data) and sum(c.text_len for c in ct.chunks) == len(data)
    assert emul.decompress(merged) == data
