"""TEST INFRASTRUCTURE ONLY - ctypes loader for oracle/libsfq_oracle.so and a driver for the
unmodified reference binary oracle/_ref/slimfastq.  Imported by tests/, bench.py's CPU legs
and __graft_entry__.smoke(); never by slimfastq_b200/ (the product path).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass, field

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsfq_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "slimfastq")
REF_SAMPLES = os.path.join(HERE, "_ref", "samples")

STREAM_NAMES = ["rec", "gen", "qlt", "gen.Ns", "gen.Nn", "rec.x", "usr.x", "usr.x.q", "usr.pfg", "usr.pfq",
                "usr.lrec", "usr.lgen", "usr.lqlt"]
NSTREAMS = len(STREAM_NAMES)


class _Chunk(C.Structure):
    _fields_ = [
        ("level", C.c_int32), ("llen", C.c_int32), ("solid", C.c_int32), ("two_id", C.c_int32),
        ("n_byte", C.c_int32), ("extra_hi_qlt", C.c_uint32), ("num_records", C.c_uint64),
        ("version", C.c_int32), ("rec_first", C.c_char * 0x200), ("rec_first_len", C.c_uint32),
        ("data", C.POINTER(C.c_uint8) * NSTREAMS), ("size", C.c_size_t * NSTREAMS),
    ]


@dataclass
class Encoded:
    """The reference's view of one standalone file/chunk: semantic info keys + named streams."""
    level: int
    llen: int
    solid: int
    two_id: int
    n_byte: int
    num_records: int
    rec_first: bytes
    streams: dict[str, bytes] = field(default_factory=dict)
    version: int = 6            # info key `version`; < 5 = pre-v5 header stream (recs.cpp:397-398)

    def info_tuple(self):
        return (self.level, self.llen, self.solid, self.two_id, self.n_byte, self.num_records, self.rec_first)


class OracleError(RuntimeError):
    pass


_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.sfq_oracle_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_Chunk), C.c_char_p]
        _lib.sfq_oracle_encode.restype = C.c_int
        _lib.sfq_oracle_encode_pre5.argtypes = _lib.sfq_oracle_encode.argtypes
        _lib.sfq_oracle_encode_pre5.restype = C.c_int
        _lib.sfq_oracle_decode.argtypes = [C.POINTER(_Chunk), C.POINTER(C.POINTER(C.c_uint8)),
                                           C.POINTER(C.c_size_t), C.c_char_p]
        _lib.sfq_oracle_decode.restype = C.c_int
        _lib.sfq_oracle_free_chunk.argtypes = [C.POINTER(_Chunk)]
        _lib.sfq_oracle_free.argtypes = [C.c_void_p]
    return _lib


def encode(fastq: bytes, level: int, pre5: bool = False) -> Encoded:
    ck = _Chunk()
    err = C.create_string_buffer(256)
    if (lib().sfq_oracle_encode_pre5 if pre5 else lib().sfq_oracle_encode)(fastq, len(fastq), level, C.byref(ck), err):
        raise OracleError(err.value.decode("latin1"))
    streams = {}
    for i, nm in enumerate(STREAM_NAMES):
        if ck.size[i]:
            streams[nm] = C.string_at(ck.data[i], ck.size[i])
    enc = Encoded(ck.level, ck.llen, ck.solid, ck.two_id, ck.n_byte, ck.num_records,
                  C.string_at(ck.rec_first, ck.rec_first_len), streams, ck.version)
    lib().sfq_oracle_free_chunk(C.byref(ck))
    return enc


def decode(enc: Encoded) -> bytes:
    ck = _Chunk()
    ck.level, ck.llen, ck.solid, ck.two_id = enc.level, enc.llen, enc.solid, enc.two_id
    ck.n_byte, ck.num_records = enc.n_byte, enc.num_records
    ck.version = enc.version
    ck.rec_first = enc.rec_first
    ck.rec_first_len = len(enc.rec_first)
    keep = []
    for i, nm in enumerate(STREAM_NAMES):
        b = enc.streams.get(nm, b"")
        if b:
            arr = (C.c_uint8 * len(b)).from_buffer_copy(b)
            keep.append(arr)
            ck.data[i] = C.cast(arr, C.POINTER(C.c_uint8))
            ck.size[i] = len(b)
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    err = C.create_string_buffer(256)
    if lib().sfq_oracle_decode(C.byref(ck), C.byref(out), C.byref(n), err):
        raise OracleError(err.value.decode("latin1"))
    res = C.string_at(out, n.value)
    lib().sfq_oracle_free(out)
    return res


# ------------------------------------------------------------------ the real reference binary
def have_ref() -> bool:
    return os.access(REF_BIN, os.X_OK)


def ref_encode(fastq: bytes, level: int, tmpdir: str | None = None) -> Encoded:
    """Run the unmodified reference on `fastq` as a standalone file and split its .sfq."""
    from . import sfq_extract  # noqa: PLC0415

    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        src, dst = os.path.join(d, "in.fq"), os.path.join(d, "out.sfq")
        with open(src, "wb") as f:
            f.write(fastq)
        r = subprocess.run([REF_BIN, "-u", src, "-f", dst, "-O", "-q", "-l", str(level)],
                           capture_output=True)
        if r.returncode:
            raise OracleError(r.stderr.decode("latin1").strip())
        info, streams = sfq_extract.extract(open(dst, "rb").read())
    return Encoded(int(info["config.level"]), int(info.get("llen", 0)), int(info.get("usr.solid", 0)),
                   int(info.get("usr.2id", 0)), int(info.get("gen.N_byte", 0)),
                   int(info.get("num_records", 0)), info.get("rec.first", "").encode("latin1"), streams)


def ref_roundtrip(fastq: bytes, level: int) -> bytes:
    with tempfile.TemporaryDirectory() as d:
        src, dst, back = (os.path.join(d, x) for x in ("in.fq", "out.sfq", "back.fq"))
        with open(src, "wb") as f:
            f.write(fastq)
        subprocess.run([REF_BIN, "-u", src, "-f", dst, "-O", "-q", "-l", str(level)], check=True)
        subprocess.run([REF_BIN, "-d", "-f", dst, "-u", back, "-O"], check=True)
        return open(back, "rb").read()
