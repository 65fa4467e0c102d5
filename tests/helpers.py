"""Shared checks: a container produced by the CUDA path versus the oracle, chunk by chunk."""
from __future__ import annotations

import hashlib

from slimfastq_b200 import container as K


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def check_container_against_oracle(O, data: bytes, blob: bytes, level: int):
    """Every chunk's info keys and every named stream must equal what the oracle (== the reference
    run on that chunk as a standalone file) produces.  Returns the parsed container."""
    ct = K.parse(blob)
    pos = 0
    for i, ch in enumerate(ct.chunks):
        part = data[pos:pos + ch.text_len]
        pos += ch.text_len
        o = O.encode(part, level)
        assert o.info_tuple() == ch.info_tuple(), f"chunk {i}: info keys differ: {o.info_tuple()} vs {ch.info_tuple()}"
        assert sorted(o.streams) == sorted(ch.streams), f"chunk {i}: stream sets differ: {sorted(o.streams)} vs {sorted(ch.streams)}"
        for nm in o.streams:
            assert o.streams[nm] == ch.streams[nm], f"chunk {i}: stream {nm} differs ({len(o.streams[nm])} vs {len(ch.streams[nm])} bytes)"
    assert pos == len(data), "chunks do not tile the input"
    return ct


def summarize(data: bytes, blob: bytes) -> dict:
    """Size-independent fingerprint of a container: md5 over the per-chunk stream md5s."""
    ct = K.parse(blob)
    h = hashlib.md5()
    for ch in ct.chunks:
        for nm in K.STREAM_NAMES:
            if nm in ch.streams:
                h.update(nm.encode() + b":" + md5(ch.streams[nm]).encode())
        h.update(repr(ch.info_tuple()).encode())
    return {"nchunks": len(ct.chunks), "stream_bytes": ct.stream_bytes, "digest": h.hexdigest()}


def container_from_oracle(enc, orig_size: int, flags: int = 1) -> bytes:
    """A single-chunk container around the streams of an oracle / reference `Encoded` (what sfq_import_reference builds
    from a reference file): plane sizes are upper bounds (flag 1 = SFQ_BLOB_IMPORTED), flag 2 = SFQ_BLOB_PRE5."""
    import struct

    ss = [len(enc.streams.get(nm, b"")) for nm in K.STREAM_NAMES]
    hdr = K.BLOB_HDR.pack(K.BLOB_MAGIC, enc.level, orig_size, orig_size + 16 * enc.num_records + 4096, enc.num_records,
                          orig_size, orig_size, orig_size, enc.llen, enc.solid, enc.two_id, enc.n_byte if enc.n_byte != ord("N") else 0, flags,
                          0, len(enc.rec_first), 0, 0, 0, 0, 0, 0, *ss)
    body = hdr + enc.rec_first + b"".join(enc.streams.get(nm, b"") for nm in K.STREAM_NAMES)
    index_off = K.FILE_HDR.size + len(body)
    head = K.FILE_HDR.pack(K.STAMP, K.KIND, 6, enc.level, orig_size, 1, orig_size, index_off, orig_size + 16 * enc.num_records + 4096)
    return head + body + struct.pack("<Q", K.FILE_HDR.size)
