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


def test_record_boundaries_of_the_c_abi_match_the_python_rule():
    """sfq_record_start_at_or_after / sfq_last_record_start (the streaming CLI's segment cuts) against
    api.record_start_at_or_after, including quality lines that begin with '@' and '+'."""
    import ctypes as C
    import random

    from slimfastq_b200 import api, synth

    L = api.load_library()
    recs = [b"@r%d extra\nACGTNACGT\n+\n@+@II+@II\n" % i if i % 3 == 0 else b"@r%d\nAC\n+r%d\n+@\n" % (i, i) for i in range(400)]
    for data in (b"".join(recs), synth.illumina(300), synth.ont(6)):
        starts = [0]
        while True:                                   # ground truth by walking four lines at a time
            p = starts[-1]
            for _ in range(4):
                p = data.index(b"\n", p) + 1
            if p >= len(data):
                break
            starts.append(p)
        buf = C.cast(C.c_char_p(data), C.c_void_p)
        rnd = random.Random(5)
        for pos in [0, 1, len(data) - 1, len(data)] + [rnd.randrange(len(data)) for _ in range(300)]:
            want = next((s for s in starts if s >= pos), len(data))
            assert L.sfq_record_start_at_or_after(buf, len(data), pos) == want == api.record_start_at_or_after(data, pos)
        for cut in [len(data)] + [rnd.randrange(40, len(data)) for _ in range(200)]:
            got = L.sfq_last_record_start(buf, cut)
            # a start only counts if its '+' line begins inside the buffer (otherwise it cannot be told from a quality line)
            ok = [s for s in starts if 0 < s < cut and data.find(b"\n", data.find(b"\n", s) + 1) + 1 < cut]
            assert got == (ok[-1] if ok else 0), (cut, got)
