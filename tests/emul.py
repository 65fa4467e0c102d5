"""Loader for tests/emul/libsfq_emul.so: the kernels' per-chunk routines compiled for the CPU so
their logic can be checked without a GPU (test tooling, never shipped)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "sfq_emul.cpp")
LIB = os.path.join(HERE, "emul", "libsfq_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        csrc = os.path.join(os.path.dirname(HERE), "slimfastq_b200", "csrc")
        deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", LIB], check=True)
        L = C.CDLL(LIB)
        L.sfq_emul_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint64, C.POINTER(C.POINTER(C.c_uint8)),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
        L.sfq_emul_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t),
                                          C.POINTER(C.c_uint32)]
        L.sfq_emul_free.argtypes = [C.c_void_p]
        L.sfq_emul_set_two_phase.argtypes = [C.c_int]
        L.sfq_emul_set_chunk_phase.argtypes = [C.c_uint64]
        _lib = L
    return _lib


class EmulError(RuntimeError):
    pass


def compress(data: bytes, level: int, chunk_bytes: int = 1 << 20, two_phase: bool = False, phase: int = 0) -> bytes:
    out, n, st = C.POINTER(C.c_uint8)(), C.c_size_t(), C.c_uint32()
    lib().sfq_emul_set_two_phase(1 if two_phase else 0)
    lib().sfq_emul_set_chunk_phase(phase)
    if lib().sfq_emul_compress(data, len(data), level, chunk_bytes, C.byref(out), C.byref(n), C.byref(st)):
        raise EmulError(f"status {st.value}")
    r = C.string_at(out, n.value)
    lib().sfq_emul_free(out)
    return r


def decompress(blob: bytes) -> bytes:
    out, n, st = C.POINTER(C.c_uint8)(), C.c_size_t(), C.c_uint32()
    if lib().sfq_emul_decompress(blob, len(blob), C.byref(out), C.byref(n), C.byref(st)):
        raise EmulError(f"status {st.value}")
    r = C.string_at(out, n.value)
    lib().sfq_emul_free(out)
    return r
