"""Reader for the chunked .sfq container (slimfastq_b200/csrc/sfq_container.h).

Pure host-side bookkeeping used by the Python API, the tests and bench.py: it only slices
bytes, it never codes anything.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

STAMP = b"whoami=slimfastq"
KIND = b"\nformat=b200.c2\n"
STREAM_NAMES = ["rec", "gen", "qlt", "gen.Ns", "gen.Nn", "rec.x", "usr.x", "usr.x.q", "usr.pfg", "usr.pfq",
                "usr.lrec", "usr.lgen", "usr.lqlt"]
FILE_HDR = struct.Struct("<16s16sIIQQQQQ")
BLOB_HDR = struct.Struct("<IIQQIIIIiBBBBIIIIIIII13I")
BLOB_MAGIC = 0x43514653


@dataclass
class Chunk:
    level: int
    text_len: int
    out_len: int
    num_records: int
    nbases: int
    nquals: int
    hdr_bytes: int
    llen: int
    solid: int
    two_id: int
    n_byte: int
    extra_hi: int
    rec_first: bytes
    streams: dict[str, bytes] = field(default_factory=dict)
    blob_bytes: int = 0

    def info_tuple(self):
        return (self.level, self.llen, self.solid, self.two_id, self.n_byte, self.num_records, self.rec_first)


@dataclass
class Container:
    level: int
    orig_size: int
    chunk_bytes: int
    chunks: list[Chunk]
    out_size: int = 0

    @property
    def stream_bytes(self) -> int:
        return sum(len(s) for c in self.chunks for s in c.streams.values())


def is_container(blob: bytes) -> bool:
    return len(blob) >= FILE_HDR.size and blob[:16] == STAMP and blob[16:32] == KIND


def parse_blob(blob: bytes, off: int = 0) -> Chunk:
    """One chunk blob (header, rec.first, streams) starting at `off`."""
    f = BLOB_HDR.unpack_from(blob, off)
    if f[0] != BLOB_MAGIC:
        raise ValueError("bad chunk magic")
    (_, lvl, text_len, out_len, nrec, nb, nq, hb, llen, solid, two_id, n_byte, _pad, extra_hi, rfl, _qu, _gu,
     _nbig, _bb, _bq, _bh) = f[:21]
    ssize = f[21:]
    p = off + BLOB_HDR.size
    rec_first = blob[p:p + rfl]
    p += rfl
    streams = {}
    for nm, sz in zip(STREAM_NAMES, ssize):
        if sz:
            streams[nm] = blob[p:p + sz]
        p += sz
    return Chunk(lvl, text_len, out_len, nrec, nb, nq, hb, llen, solid, two_id, n_byte, extra_hi, rec_first, streams, p - off)


def parse(blob: bytes) -> Container:
    if not is_container(blob):
        raise ValueError("not a b200 chunked .sfq container")
    _, _, version, level, orig, nchunks, chunk_bytes, index_off, out_size = FILE_HDR.unpack_from(blob, 0)
    offs = struct.unpack_from(f"<{nchunks}Q", blob, index_off)
    return Container(level, orig, chunk_bytes, [parse_blob(blob, off) for off in offs], out_size)
