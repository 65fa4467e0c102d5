"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

The reference drives the path from two loops, `UsrSave().encode()` (usrs.cpp:392-407) and
`UsrLoad().decode()` (usrs.cpp:539-574), selected by `Config::encode` (main.cpp:53-57); `Codec`
offers the same pair as `compress` / `decompress` with the reference's level semantics
(`-l 1..4`, clamped, config.cpp:231-236) and its error behaviour (a croak message, here raised as
`SfqError` instead of exit(1)).  Everything is computed by libsfq_b200.so's CUDA kernels; this
module is ctypes glue and fails loudly when the library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFQ_B200_LIB") or os.path.join(HERE, "libsfq_b200.so")    # override = A/B kernel variants
DEFAULT_CHUNK = 1 << 20


class SfqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"slimfastq: {msg} (code {code})")
        self.code = code


class Stats(C.Structure):
    _fields_ = [
        ("in_bytes", C.c_uint64), ("out_bytes", C.c_uint64), ("nchunks", C.c_uint64), ("nrecords", C.c_uint64),
        ("nbases", C.c_uint64), ("nquals", C.c_uint64), ("stream_bytes", C.c_uint64),
        ("waves", C.c_uint32), ("resident_chunks", C.c_uint32), ("kernel_launches", C.c_uint32), ("retries", C.c_uint32),
        ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_scan", C.c_float),
        ("ms_plan", C.c_float), ("ms_clear", C.c_float), ("ms_code", C.c_float), ("ms_pack", C.c_float),
        ("workspace_bytes", C.c_uint64), ("ms_gen", C.c_float), ("ms_qlt", C.c_float), ("ms_rec", C.c_float),
        ("gen_stream_bytes", C.c_uint64), ("qlt_stream_bytes", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "sfq_create", "sfq_destroy", "sfq_last_error", "sfq_version", "sfq_set_max_resident", "sfq_host_alloc",
    "sfq_host_free", "sfq_compress", "sfq_compress_device", "sfq_compress_bound", "sfq_decompress",
    "sfq_decompress_device", "sfq_decompressed_size", "sfq_get_stats",
    "sfq_is_reference_file", "sfq_export_reference_bound", "sfq_export_reference",
    "sfq_import_reference_bound", "sfq_import_reference",
    "sfq_record_start_at_or_after", "sfq_last_record_start", "sfq_set_chunk_phase", "sfq_stream_cut",
    "sfq_encode_gen_chunks", "sfq_encode_qlt_chunks", "sfq_encode_rec_chunks",
    "sfq_decode_gen_chunks", "sfq_decode_qlt_chunks", "sfq_decode_rec_chunks", "sfq_trim",
    "sfq_device_numa_node",
]

_lib = None


def load_library():
    """dlopen libsfq_b200.so (no fallback: a missing build is an error, not a slow path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfqError(1, f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, u8p, sz = C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t
    L.sfq_create.argtypes = [C.POINTER(vp), C.c_int]; L.sfq_create.restype = C.c_int
    L.sfq_destroy.argtypes = [vp]; L.sfq_destroy.restype = None
    L.sfq_last_error.argtypes = [vp]; L.sfq_last_error.restype = C.c_char_p
    L.sfq_version.argtypes = []; L.sfq_version.restype = C.c_char_p
    L.sfq_set_max_resident.argtypes = [vp, C.c_uint32]; L.sfq_set_max_resident.restype = C.c_int
    L.sfq_trim.argtypes = [vp]; L.sfq_trim.restype = C.c_int
    L.sfq_device_numa_node.argtypes = [C.c_int]; L.sfq_device_numa_node.restype = C.c_int
    L.sfq_set_chunk_phase.argtypes = [vp, C.c_uint64]; L.sfq_set_chunk_phase.restype = C.c_int
    L.sfq_host_alloc.argtypes = [sz]; L.sfq_host_alloc.restype = vp
    L.sfq_host_free.argtypes = [vp]; L.sfq_host_free.restype = None
    L.sfq_compress.argtypes = [vp, vp, sz, C.c_int, C.c_uint64, C.POINTER(u8p), C.POINTER(sz)]; L.sfq_compress.restype = C.c_int
    L.sfq_compress_device.argtypes = [vp, vp, sz, C.c_int, C.c_uint64, vp, sz, C.POINTER(sz)]; L.sfq_compress_device.restype = C.c_int
    L.sfq_compress_bound.argtypes = [sz, C.c_uint64]; L.sfq_compress_bound.restype = sz
    L.sfq_decompress.argtypes = [vp, vp, sz, C.POINTER(u8p), C.POINTER(sz)]; L.sfq_decompress.restype = C.c_int
    L.sfq_decompress_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]; L.sfq_decompress_device.restype = C.c_int
    L.sfq_decompressed_size.argtypes = [vp, sz, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]; L.sfq_decompressed_size.restype = C.c_int
    L.sfq_get_stats.argtypes = [vp, C.POINTER(Stats)]; L.sfq_get_stats.restype = C.c_int
    L.sfq_is_reference_file.argtypes = [vp, sz]; L.sfq_is_reference_file.restype = C.c_int
    L.sfq_export_reference_bound.argtypes = [vp, sz]; L.sfq_export_reference_bound.restype = sz
    L.sfq_export_reference.argtypes = [vp, sz, C.c_char_p, vp, sz, C.POINTER(sz)]; L.sfq_export_reference.restype = C.c_int
    L.sfq_import_reference_bound.argtypes = [sz]; L.sfq_import_reference_bound.restype = sz
    L.sfq_import_reference.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]; L.sfq_import_reference.restype = C.c_int
    L.sfq_record_start_at_or_after.argtypes = [vp, sz, sz]; L.sfq_record_start_at_or_after.restype = sz
    L.sfq_last_record_start.argtypes = [vp, sz]; L.sfq_last_record_start.restype = sz
    L.sfq_stream_cut.argtypes = [vp, sz, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]; L.sfq_stream_cut.restype = sz
    for pl in ("gen", "qlt", "rec"):
        f = getattr(L, f"sfq_encode_{pl}_chunks"); f.argtypes = L.sfq_compress.argtypes; f.restype = C.c_int
        f = getattr(L, f"sfq_decode_{pl}_chunks"); f.argtypes = L.sfq_decompress.argtypes; f.restype = C.c_int
    _lib = L
    return L


def _host_ptr(buf):
    """(address, nbytes, keepalive) of a bytes-like / numpy / CPU torch uint8 buffer."""
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr(), buf.numel() * buf.element_size(), buf
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value, len(buf), buf
    if isinstance(buf, bytearray):
        arr = (C.c_char * len(buf)).from_buffer(buf)
        return C.addressof(arr), len(buf), arr
    import numpy as np  # noqa: PLC0415

    a = np.ascontiguousarray(buf)
    return a.ctypes.data, a.nbytes, a


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_device_node(device: int):
    """Bind the calling process to the cores of the NUMA node the GPU hangs off, so that pinned buffers allocated from now
    on are placed there (first touch) and host<->device copies stay off the inter-socket link.  Returns
    (node, previous affinity set) or (None, None) when the node is unknown or has no core this process may use; restore with
    os.sched_setaffinity(0, previous) before starting CPU work that should use every core."""
    try:
        node = load_library().sfq_device_numa_node(device)
        if node < 0:
            return None, None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        prev = os.sched_getaffinity(0)
        want = cpus & prev
        if not want or want == prev:
            return (node, prev) if want else (None, None)
        os.sched_setaffinity(0, want)
        return node, prev
    except (OSError, ValueError, AttributeError):
        return None, None


class Codec:
    """One context = one GPU = one host thread (mirrors the reference's one-process-per-file model)."""

    def __init__(self, device: int = -1, max_resident: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.sfq_create(C.byref(h), device)
        if rc:
            raise SfqError(rc, "cannot create a context: no usable CUDA device (there is no CPU fallback)")
        self._h = h
        if max_resident:
            self._L.sfq_set_max_resident(h, max_resident)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sfq_destroy(self._h)
            self._h = None

    __del__ = close

    def trim(self):
        """Release the device / pinned memory the context caches between calls (views returned earlier become invalid)."""
        self._check(self._L.sfq_trim(self._h))

    def _check(self, rc: int):
        if rc:
            raise SfqError(rc, self._L.sfq_last_error(self._h).decode("latin1"))

    # ---- host buffers in, host bytes out (the call a user of the reference makes)
    def compress_view(self, fastq, level: int = 3, chunk_bytes: int = DEFAULT_CHUNK, phase: int = 0):
        """Returns (address, nbytes) of the context-owned pinned result, valid until the next call.
        `phase`: chunk-grid offset when `fastq` is a part of a larger file (see split_on_grid)."""
        addr, n, keep = _host_ptr(fastq)
        if phase:
            self._L.sfq_set_chunk_phase(self._h, phase)
        out = C.POINTER(C.c_uint8)()
        on = C.c_size_t()
        self._check(self._L.sfq_compress(self._h, addr, n, level, chunk_bytes, C.byref(out), C.byref(on)))
        del keep
        return C.cast(out, C.c_void_p).value, on.value

    def compress(self, fastq, level: int = 3, chunk_bytes: int = DEFAULT_CHUNK, phase: int = 0) -> bytes:
        addr, n = self.compress_view(fastq, level, chunk_bytes, phase)
        return C.string_at(addr, n)

    def decompress_view(self, sfq):
        addr, n, keep = _host_ptr(sfq)
        out = C.POINTER(C.c_uint8)()
        on = C.c_size_t()
        self._check(self._L.sfq_decompress(self._h, addr, n, C.byref(out), C.byref(on)))
        del keep
        return C.cast(out, C.c_void_p).value, on.value

    def decompress_addr(self, addr: int, n: int):
        """decompress_view for a container already in host memory at `addr` (e.g. compress_view's result)."""
        out = C.POINTER(C.c_uint8)()
        on = C.c_size_t()
        self._check(self._L.sfq_decompress(self._h, addr, n, C.byref(out), C.byref(on)))
        return C.cast(out, C.c_void_p).value, on.value

    def decompress(self, sfq) -> bytes:
        addr, n = self.decompress_view(sfq)
        return C.string_at(addr, n)

    # ---- per-plane test hooks (include/sfq_b200.h): one plane's coders only
    def encode_plane(self, plane: str, fastq, level: int = 3, chunk_bytes: int = DEFAULT_CHUNK) -> bytes:
        addr, n, keep = _host_ptr(fastq)
        out = C.POINTER(C.c_uint8)()
        on = C.c_size_t()
        self._check(getattr(self._L, f"sfq_encode_{plane}_chunks")(self._h, addr, n, level, chunk_bytes, C.byref(out), C.byref(on)))
        del keep
        return C.string_at(out, on.value)

    def decode_plane(self, plane: str, sfq) -> bytes:
        addr, n, keep = _host_ptr(sfq)
        out = C.POINTER(C.c_uint8)()
        on = C.c_size_t()
        self._check(getattr(self._L, f"sfq_decode_{plane}_chunks")(self._h, addr, n, C.byref(out), C.byref(on)))
        del keep
        return C.string_at(out, on.value)

    # ---- device-resident variants (torch CUDA uint8 tensors; caller synchronises its own stream first)
    def compress_device(self, d_fastq, d_out, level: int = 3, chunk_bytes: int = DEFAULT_CHUNK, nbytes: int | None = None) -> int:
        on = C.c_size_t()
        n = d_fastq.numel() if nbytes is None else nbytes
        self._check(self._L.sfq_compress_device(self._h, d_fastq.data_ptr(), n, level, chunk_bytes,
                                                d_out.data_ptr(), d_out.numel(), C.byref(on)))
        return on.value

    def decompress_device(self, d_sfq, nbytes: int, d_out) -> int:
        on = C.c_size_t()
        self._check(self._L.sfq_decompress_device(self._h, d_sfq.data_ptr(), nbytes, d_out.data_ptr(), d_out.numel(), C.byref(on)))
        return on.value

    def compress_bound(self, n: int, chunk_bytes: int = DEFAULT_CHUNK) -> int:
        return self._L.sfq_compress_bound(n, chunk_bytes)

    def stats(self) -> dict:
        st = Stats()
        self._L.sfq_get_stats(self._h, C.byref(st))
        return st.as_dict()


def decompressed_size(sfq: bytes) -> tuple[int, int]:
    n = C.c_uint64()
    lv = C.c_int()
    rc = load_library().sfq_decompressed_size(C.cast(C.c_char_p(sfq[:128]), C.c_void_p), len(sfq), C.byref(n), C.byref(lv))
    if rc:
        raise SfqError(rc, "not a b200 chunked .sfq container")
    return n.value, lv.value


# ---------------------------------------------------------------------------- the reference's own file format
# Host-side format conversions (no coding, no GPU): see include/sfq_b200.h.
def is_reference_file(blob: bytes) -> bool:
    return bool(load_library().sfq_is_reference_file(C.cast(C.c_char_p(blob), C.c_void_p), len(blob)))


def export_reference(sfq: bytes, orig_filename: str = "") -> bytes:
    """Single-chunk container -> a file the unmodified reference binary decodes."""
    L = load_library()
    src = C.cast(C.c_char_p(sfq), C.c_void_p)
    cap = L.sfq_export_reference_bound(src, len(sfq))
    out = C.create_string_buffer(cap)
    on = C.c_size_t()
    rc = L.sfq_export_reference(src, len(sfq), orig_filename.encode(), C.cast(out, C.c_void_p), cap, C.byref(on))
    if rc:
        raise SfqError(rc, "cannot export: not a single-chunk b200 container")
    return out.raw[:on.value]


def import_reference(ref: bytes) -> bytes:
    """Reference-written .sfq file -> single-chunk container for Codec.decompress."""
    L = load_library()
    cap = L.sfq_import_reference_bound(len(ref))
    out = C.create_string_buffer(cap)
    on = C.c_size_t()
    rc = L.sfq_import_reference(C.cast(C.c_char_p(ref), C.c_void_p), len(ref), C.cast(out, C.c_void_p), cap, C.byref(on))
    if rc:
        raise SfqError(rc, "cannot import: not a reference .sfq file, or one without orig.size / with oversized-record streams")
    return out.raw[:on.value]


# ---------------------------------------------------------------------------- sharding (multi-GPU)
def record_start_at_or_after(fastq: bytes, pos: int) -> int:
    """Smallest record start >= pos (host-side, no coding).  A line that starts with '@' is a header
    iff the line two below starts with '+': a quality line beginning with '@' is followed by a header
    and then bases, never by a '+' line two below."""
    n = len(fastq)
    if pos <= 0:
        return 0
    p = fastq.find(b"\n", pos - 1)
    while 0 <= p < n - 1:
        s = p + 1
        if fastq[s:s + 1] == b"@":
            l2 = fastq.find(b"\n", s)
            l3 = fastq.find(b"\n", l2 + 1) if l2 >= 0 else -1
            if l3 >= 0 and fastq[l3 + 1:l3 + 2] == b"+":
                return s
        p = fastq.find(b"\n", s)
    return n


def chunk_lengths(fastq: bytes, chunk_bytes: int = DEFAULT_CHUNK, phase: int = 0) -> list[int]:
    """Byte lengths of the chunks the device planner forms (records starting in [c*B, (c+1)*B) of the
    grid shifted by `phase`, sfq_plan.cuh) - used to hand the CPU reference the very same chunks."""
    n, cuts = len(fastq), [0]
    c = 1
    while cuts[-1] < n:
        s = record_start_at_or_after(fastq, c * chunk_bytes - phase) if c * chunk_bytes - phase < n else n
        c += 1
        if s > cuts[-1]:
            cuts.append(s)
    return [cuts[i + 1] - cuts[i] for i in range(len(cuts) - 1)]


def split_records(fastq: bytes, parts: int) -> list[tuple[int, int]]:
    """Cut FASTQ text into `parts` byte ranges on record boundaries (shards for `parts` GPUs)."""
    n = len(fastq)
    cuts = [0]
    for p in range(1, parts):
        cuts.append(max(cuts[-1], record_start_at_or_after(fastq, n * p // parts)))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(parts)]


def split_on_grid(fastq: bytes, parts: int, chunk_bytes: int = DEFAULT_CHUNK) -> list[tuple[int, int, int]]:
    """Cut FASTQ text into up to `parts` ranges that respect the chunk grid: (start, end, phase) per range,
    every range beginning with the first record at or after a grid line k*chunk_bytes (phase = start - k*B).
    Compressing the ranges with their phases and merging gives the one-call container byte for byte -
    "chunks partitioned by index across the GPUs".  Ranges without records are dropped."""
    n = len(fastq)
    nslots = (n + chunk_bytes - 1) // chunk_bytes
    lines = sorted({min(nslots, round(nslots * p / parts)) for p in range(parts)} | {0})
    cuts = []
    for k in lines:
        g = record_start_at_or_after(fastq, k * chunk_bytes) if k else 0
        cuts.append((g, g - k * chunk_bytes))
    out = []
    for i, (g, ph) in enumerate(cuts):
        end = cuts[i + 1][0] if i + 1 < len(cuts) else n
        if end > g:
            out.append((g, end, ph))
    return out


def merge_containers(parts: list[bytes]) -> bytes:
    """Lay several rank-local containers out as one: blobs back to back in rank order, then the
    index rebuilt from the exchanged sizes (the only cross-rank step of the data path)."""
    import struct  # noqa: PLC0415

    from . import container as K  # noqa: PLC0415

    hdrs = [K.FILE_HDR.unpack_from(p, 0) for p in parts]
    level, chunk_bytes = hdrs[0][3], hdrs[0][6]
    body = bytearray()
    index: list[int] = []
    orig = out_size = 0
    for p, h in zip(parts, hdrs):
        _, _, _version, lv, o, nchunks, cb, index_off, osz = h
        if lv != level:
            raise ValueError("containers of different levels")
        offs = struct.unpack_from(f"<{nchunks}Q", p, index_off)
        index.extend(off + len(body) for off in offs)      # blobs keep their order, shifted by what precedes
        body += p[K.FILE_HDR.size:index_off]
        orig += o
        out_size += osz
    index_off = K.FILE_HDR.size + len(body)
    head = K.FILE_HDR.pack(K.STAMP, K.KIND, hdrs[0][2], level, orig, len(index), chunk_bytes, index_off, out_size)
    return bytes(head) + bytes(body) + struct.pack(f"<{len(index)}Q", *index)
