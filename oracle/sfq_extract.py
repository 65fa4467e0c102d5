"""TEST INFRASTRUCTURE ONLY - never imported by the product path (slimfastq_b200/).

Reader for the reference's `.sfq` WORM page container, so that a file written by the
unmodified reference binary (oracle/_ref/slimfastq) can be split into its named byte
streams and compared with ours.

Format restated from /root/reference/filer.cpp:41-53 (file table entry), :94-106 (table in
page 1, entry 0 = info stream whose `first` holds the entry count), filer.hpp:34-41 and
filer.cpp:210-242, 273-303 (a stream = its `first` page, then the pages listed in its node
page; a node page is u32[2048]: 2047 data pages + the next node page).
"""
from __future__ import annotations

import struct

PAGE = 0x2000
NODE_MAX = PAGE // 4 - 1  # 2047 data-page slots per node page
ENTRY = struct.Struct("<QQII")  # name[8], size, first, node (packed, 24 bytes)


def _read_stream(blob: bytes, size: int, first: int, node: int) -> bytes:
    out = bytearray()
    if size == 0:
        return bytes(out)
    out += blob[first * PAGE:(first + 1) * PAGE]
    while len(out) < size and node:
        idx = struct.unpack_from("<2048I", blob, node * PAGE)
        for k in range(NODE_MAX):
            if len(out) >= size:
                break
            pg = idx[k]
            out += blob[pg * PAGE:(pg + 1) * PAGE]
        node = idx[NODE_MAX]
    return bytes(out[:size])


def extract(blob: bytes) -> tuple[dict[str, str], dict[str, bytes]]:
    """Return (info key->value, stream name->bytes) of a reference-written .sfq file."""
    if not blob.startswith(b"whoami=slimfastq"):
        raise ValueError("not a reference .sfq file (stamp missing)")
    table = blob[PAGE:2 * PAGE]
    name0, size0, count, node0 = ENTRY.unpack_from(table, 0)
    info_raw = _read_stream(blob, size0, 0, node0)
    info: dict[str, str] = {}
    for line in info_raw.split(b"\n"):
        if b"=" in line:
            k, v = line.split(b"=", 1)
            # std::map::insert keeps the FIRST value of a duplicated key (config.cpp:155)
            info.setdefault(k.decode("latin1"), v.decode("latin1"))
    streams: dict[str, bytes] = {}
    for i in range(1, count):
        name, size, first, node = ENTRY.unpack_from(table, i * ENTRY.size)
        nm = struct.pack("<Q", name).rstrip(b"\0").decode("latin1")
        streams[nm] = _read_stream(blob, size, first, node)
    return info, streams


if __name__ == "__main__":
    import hashlib
    import sys

    info, streams = extract(open(sys.argv[1], "rb").read())
    for k, v in info.items():
        print(f"{k}={v}")
    for k, v in streams.items():
        print(f"[{k}] {len(v)} {hashlib.md5(v).hexdigest()}")
