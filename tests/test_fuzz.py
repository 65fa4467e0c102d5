"""CPU: seeded random FASTQ files - odd header layouts, changing field types, variable lengths, N runs, lower case, '.' bases,
qualities up to '~', '!' under real bases, 2nd ids, SOLiD - through three implementations that must agree bit for bit:
the unmodified reference binary, the oracle (plain-C restatement) and the CPU build of the kernels' per-chunk routines
(tests/emul; the two-phase encoder form too, whole file and in chunks).  The fixed samples pin the common paths; this
walks combinations nobody wrote a fixture for."""
import numpy as np
import pytest

import emul
from helpers import check_container_against_oracle
from oracle import oracle as O

SEEDS = list(range(40))


def random_fastq(seed: int) -> bytes:
    rng = np.random.default_rng(0xF00D + seed)
    nrec = int(rng.integers(1, 140))
    solid = seed % 8 == 5
    two_id = rng.random() < 0.3
    fixed_len = rng.random() < 0.5
    L0 = int(rng.integers(1 if not solid else 2, 260))
    qlo = int(rng.integers(33, 60))
    qhi = int(rng.integers(qlo + 1, 127 if rng.random() < 0.3 else 75))
    style = int(rng.integers(0, 5))
    nbyte = b"." if (solid or rng.random() < 0.15) else b"N"
    out = []
    lane, tile, x, y = 1, 1101, 1000, 2000
    for i in range(nrec):
        # ---- header: a few families, with the occasional change of layout (rec.x) and of field type
        if rng.random() < 0.06:
            style = int(rng.integers(0, 5))
        x += int(rng.integers(0, 300)); y = int(rng.integers(0, 99999))
        if rng.random() < 0.1:
            tile += 1
        if style == 0:
            h = b"@HWI-ST%d:%d:C0FJ%dACXX:%d:%d:%d:%d %d:N:0:%s" % (700 + seed, 100 + seed, seed, lane, tile, x, y, 1 + (i & 1), b"ACGT"[: 1 + i % 4])
        elif style == 1:
            h = b"@SRR%07d.%d %d/%d" % (seed, i + 1, i + 1, 1 + (i & 1))
        elif style == 2:
            h = b"@read_%04x_%03d  len=%d\tflag=0x%X;q=%s" % (i * 7919 & 0xFFFF, i % 1000, L0, i * 31, b"%.2f" % (i / 7.0))
        elif style == 3:
            h = b"@%d" % (10 ** int(rng.integers(0, 19)) + i)                      # bare numbers up to 19 digits, leading-digit changes
        else:
            h = b"@" + bytes(rng.integers(33, 127, int(rng.integers(1, 60))).astype(np.uint8)).replace(b"\n", b"_")
        if rng.random() < 0.05:
            h += b" " + b"0" * int(rng.integers(1, 5)) + b"%d" % i                  # leading zeros
        # ---- bases
        L = L0 if fixed_len else int(rng.integers(1 if not solid else 2, 260))
        if solid:
            s = np.frombuffer(b"0123", dtype=np.uint8)[rng.integers(0, 4, L)].copy()
            s[0] = ord("T") if i % 11 else ord("G")
            if rng.random() < 0.2 and L > 3:
                s[1 + int(rng.integers(0, L - 1))] = ord(".")
        else:
            s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)].copy()
            if rng.random() < 0.3:
                a = int(rng.integers(0, L)); b = min(L, a + int(rng.integers(1, 12)))
                s[a:b] = nbyte[0]
            if rng.random() < 0.1:
                s[int(rng.integers(0, L))] |= 0x20                                   # lower case (the reference folds it)
        q = rng.integers(qlo, qhi + 1, L).astype(np.uint8)
        if rng.random() < 0.25:
            q[int(rng.integers(0, L))] = ord("!")                                    # '!' under whatever base is there
        if rng.random() < 0.3:
            q[np.flatnonzero(s == nbyte[0])] = ord("!") if rng.random() < 0.5 else ord("#")
        plus = b"+" + (h[1:] if two_id else b"")
        out.append(h + b"\n" + s.tobytes() + b"\n" + plus + b"\n" + q.tobytes() + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("seed", SEEDS)
def test_reference_oracle_and_kernel_routines_agree(oracle, seed):
    if not O.have_ref():
        pytest.skip("oracle/_ref/slimfastq is not built")
    data = random_fastq(seed)
    for level in ((1, 3) if seed % 2 else (2, 4)):
        try:
            ref = O.ref_encode(data, level)
        except Exception:                      # the reference refuses the file (e.g. "switched N_byte"): so must the other two
            with pytest.raises(Exception):
                oracle.encode(data, level)
            with pytest.raises(emul.EmulError):
                emul.compress(data, level, 1 << 40)
            continue
        mine = oracle.encode(data, level)
        assert mine.info_tuple() == ref.info_tuple()
        assert mine.streams == ref.streams
        whole = emul.compress(data, level, 1 << 40, two_phase=bool(seed & 2))
        check_container_against_oracle(oracle, data, whole, level)
        assert emul.decompress(whole) == oracle.decode(mine)
        chunks = emul.compress(data, level, 4096, two_phase=not bool(seed & 2))
        check_container_against_oracle(oracle, data, chunks, level)
        assert emul.decompress(chunks) == oracle.decode(mine)


@pytest.mark.parametrize("seed", SEEDS[::3])
def test_interchange_with_the_reference_file_format(oracle, seed):
    """What the kernels' routines write, exported to the reference's page file, the unmodified binary must decode; what the
    binary writes, imported, the kernels' decoder routines must decode (format conversions are host code: no GPU needed)."""
    import tempfile

    import slimfastq_b200 as S
    from test_reference_format import ref_compress, ref_decompress

    if not O.have_ref():
        pytest.skip("oracle/_ref/slimfastq is not built")
    data = random_fastq(seed)
    level = 1 + seed % 4
    try:
        want = oracle.decode(oracle.encode(data, level))          # the reference's own round trip of this file (lossy cases included)
    except Exception:
        pytest.skip("rejected input (covered above)")
    with tempfile.TemporaryDirectory() as d:
        ours = emul.compress(data, level, 1 << 40, two_phase=bool(seed & 1))
        assert ref_decompress(S.export_reference(ours, "in.fq"), d) == want
        theirs = ref_compress(data, level, d)
        assert emul.decompress(S.import_reference(theirs)) == want


CORRUPT = r"""
import random, resource, sys
resource.setrlimit(resource.RLIMIT_AS, (16 << 30, 16 << 30))      # a wild size field must fail an allocation, not take the machine
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[3])
import slimfastq_b200 as S
import emul, test_fuzz
from slimfastq_b200 import container as K
rnd = random.Random(int(sys.argv[1]))
data = test_fuzz.random_fastq(int(sys.argv[1]))
blob = emul.compress(data, 3, 1 << 40)
ref = S.export_reference(blob, "in.fq")
ok = err = 0
for trial in range(120):
    for name, buf, fn in (("import", ref, S.import_reference), ("export", blob, lambda b: S.export_reference(b, "x.fq")),
                          ("size", blob, S.decompressed_size), ("parse", blob, K.parse), ("decode", blob, emul.decompress)):
        b = bytearray(buf)
        mode = rnd.randrange(5)
        if mode == 0:
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        elif mode == 1:
            b = b[: rnd.randrange(0, len(b))]
        elif mode == 2:
            p = rnd.randrange(len(b)); b[p:p + 8] = rnd.randbytes(8)
        elif mode == 3:
            p = rnd.randrange(min(len(b), 300)); b[p:p + 4] = (0xFFFFFFFF).to_bytes(4, "little")
        else:
            p = 80 + rnd.randrange(132); b[p] = rnd.randrange(256)        # a field of the first blob header / of the info page
        try:
            r = fn(bytes(b)); ok += 1
            if name == "import":
                try:
                    emul.decompress(r)
                except Exception:
                    pass
        except Exception:
            err += 1
print("survived", ok, err)
"""


@pytest.mark.parametrize("seed", [0, 7])
def test_corrupt_containers_never_crash_the_host_code(seed):
    """Bit flips, truncations and wild counts in a container / a reference-format file: the host-side format code (import,
    export, size query, parser) and the decoder routines behind the library's own framing checks (sfq_index_check /
    sfq_blob_check, shared with the test emulation) must answer with an error or an answer, never with a crash - run in a
    child process so that a crash is a test failure instead of the end of the test run."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-c", CORRUPT, str(seed), os.path.dirname(here), here], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "survived" in r.stdout, (r.returncode, r.stderr[-1500:])
