"""Deterministic synthetic FASTQ of the shapes BASELINE.json names.

There is no network and no dataset: every benchmark and golden-vector input comes from here.
All randomness is numpy's PCG64 seeded explicitly (bit-stable across platforms for the integer
and uniform-double draws used), so tests/golden/make_golden.py can regenerate the exact bytes
whose reference stream md5s are committed.
"""
from __future__ import annotations

import numpy as np

SEED0 = 0x5F51


def _assemble(headers: list[bytes], seq: list[np.ndarray] | np.ndarray, qual: list[np.ndarray] | np.ndarray,
              plus: list[bytes] | None = None) -> bytes:
    """Join per-record pieces into FASTQ text with one bulk numpy scatter per plane."""
    n = len(headers)
    hl = np.fromiter((len(h) for h in headers), dtype=np.int64, count=n)
    if isinstance(seq, np.ndarray):
        sl = np.full(n, seq.shape[1], dtype=np.int64)
        ql = np.full(n, qual.shape[1], dtype=np.int64)
    else:
        sl = np.fromiter((len(s) for s in seq), dtype=np.int64, count=n)
        ql = np.fromiter((len(q) for q in qual), dtype=np.int64, count=n)
    pl = np.ones(n, dtype=np.int64) if plus is None else np.fromiter((len(p) for p in plus), dtype=np.int64, count=n)
    rec_len = hl + 1 + sl + 1 + pl + 1 + ql + 1
    start = np.concatenate(([0], np.cumsum(rec_len)))
    out = np.full(int(start[-1]), ord("\n"), dtype=np.uint8)

    def scatter(offs, lens, flat):
        if flat.size == 0:
            return
        idx = np.repeat(offs - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(flat.size)
        out[idx] = flat

    scatter(start[:-1], hl, np.frombuffer(b"".join(headers), dtype=np.uint8))
    s_off = start[:-1] + hl + 1
    scatter(s_off, sl, seq.reshape(-1) if isinstance(seq, np.ndarray) else np.concatenate(seq))
    p_off = s_off + sl + 1
    if plus is None:
        out[p_off] = ord("+")
    else:
        scatter(p_off, pl, np.frombuffer(b"".join(plus), dtype=np.uint8))
    q_off = p_off + pl + 1
    scatter(q_off, ql, qual.reshape(-1) if isinstance(qual, np.ndarray) else np.concatenate(qual))
    return out.tobytes()


def _markov_quals(rng: np.random.Generator, n: int, length: int, lo: int, hi: int, start_mean: float) -> np.ndarray:
    """First-order Markov chain over Phred lo..hi, drifting down with position, with a sticky
    floor state (the '####' tails of real Illumina reads). Returns Phred values (n, length)."""
    q = np.clip(rng.normal(start_mean, 2.5, n).round(), lo, hi).astype(np.int16)
    dead = np.zeros(n, dtype=bool)
    out = np.empty((n, length), dtype=np.int16)
    for t in range(length):
        out[:, t] = np.where(dead, lo, q)
        u = rng.random(n)
        target = hi - 3 - 8.0 * (t / length) ** 2          # slow position-dependent decay
        low = q < target
        step = np.zeros(n, dtype=np.int16)
        step[u < 0.20] = -1
        step[u < 0.06] = -4
        step[u < 0.010] = -14
        step[(u > 0.55) & low] = 1
        step[(u > 0.80) & low] = 3
        step[(u > 0.85) & ~low] = 1
        q = np.clip(q + step, lo + 1, hi).astype(np.int16)
        dead |= rng.random(n) < (0.0002 + 0.00002 * t)
    return out


_BIN8 = np.zeros(64, dtype=np.int16)
for _lo, _hi, _v in ((0, 1, 2), (2, 9, 6), (10, 19, 15), (20, 24, 22), (25, 29, 27), (30, 34, 33), (35, 39, 37), (40, 63, 40)):
    _BIN8[_lo:_hi + 1] = _v      # Illumina 8-level binning {2,6,15,22,27,33,37,40}


def illumina(n_reads: int, seed: int = SEED0 + 1, read_len: int = 150, bins8: bool = False,
             first_read: int = 0) -> bytes:
    """Illumina-style 2x150 interleaved /1,/2 records, 40-level Phred ('#'..'J'), i.i.d. uniform
    ACGT bases (2 bit/base worst case for gens), 0.1 % N under '#' quality plus a few N under
    '!' quality; header @<instr>:<run>:<flowcell>:<lane>:<tile>:<x>:<y> <read>:N:0:<index>."""
    rng = np.random.default_rng(seed)
    phred = _markov_quals(rng, n_reads, read_len, 2, 41, 37.0)
    if bins8:
        phred = _BIN8[phred]
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n_reads, read_len))]
    isn = rng.random((n_reads, read_len)) < 0.001
    bang = isn & (rng.random((n_reads, read_len)) < 0.2)
    bases[isn] = ord("N")
    phred[isn] = 2
    qual = (phred + 33).astype(np.uint8)
    qual[bang] = ord("!")
    pair = np.arange(first_read, first_read + n_reads) // 2
    tile = 1101 + (pair // 4000) % 78 + 100 * ((pair // 312000) % 2)
    y = 1000 + (pair % 4000) * 25 + rng.integers(0, 25, n_reads) // 2 * 2
    y[1::2] = y[0::2][: len(y[1::2])]
    x = rng.integers(1000, 32000, n_reads)
    x[1::2] = x[0::2][: len(x[1::2])]
    lane = 1 + (pair // 2496000) % 4
    headers = [b"@A00123:45:HXXXXXXXX:%d:%d:%d:%d %d:N:0:ACGTACGT" % (lane[i], tile[i], x[i], y[i], 1 + (i + first_read) % 2)
               for i in range(n_reads)]
    return _assemble(headers, bases, qual)


def ont(n_reads: int, seed: int = SEED0 + 4, max_len: int = 50000, mu: float = 8.9) -> bytes:
    """ONT-style long reads: log-normal lengths clipped to 1 000..max_len (above 65 534 a read is an oversized record), UUID/hex
    headers like fast5.to.fq, qualities over Phred 1..60 with occasional values >= 63 (escape path,
    qlts.cpp:120-125), N runs of 1..50 under non-'!' quality (gen.Ns)."""
    rng = np.random.default_rng(seed)
    lens = np.clip(rng.lognormal(mu, 0.8, n_reads), 1000, max_len).astype(np.int64)
    run_id = "%040x" % int(rng.integers(0, 2**62))
    headers, seqs, quals = [], [], []
    t0 = 1000
    for i in range(n_reads):
        L = int(lens[i])
        uu = "%08x-%04x-%04x-%04x-%012x" % (int(rng.integers(0, 2**32)), int(rng.integers(0, 2**16)),
                                            int(rng.integers(0, 2**16)), int(rng.integers(0, 2**16)),
                                            int(rng.integers(0, 2**48)))
        t0 += int(rng.integers(1, 900))
        headers.append(("@%s runid=%s read=%d ch=%d start_time=%d" % (uu, run_id, 10 + 3 * i, int(rng.integers(1, 513)), t0)).encode())
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)]
        ph = np.clip(np.cumsum(rng.integers(-3, 4, L)) // 4 + rng.integers(8, 30), 1, 60).astype(np.int16)
        hi = rng.random(L) < 0.0005
        ph[hi] = rng.integers(63, 90, int(hi.sum()))
        for _ in range(int(rng.integers(0, 3))):
            a = int(rng.integers(0, L - 60))
            b = a + int(rng.integers(1, 51))
            s[a:b] = ord("N")
            ph[a:b] = np.maximum(ph[a:b], 1)
        seqs.append(s)
        quals.append((ph + 33).astype(np.uint8))
    return _assemble(headers, seqs, quals)


def edge_cases(seed: int = SEED0 + 9) -> dict[str, bytes]:
    """Small inputs that walk the reference's special paths (one per fixture family in SURVEY.md
    section 4): SOLiD colour space with prefix changes, '.' as the N byte, 2nd-id '+' lines,
    variable lengths, qlen != llen, qualities >= 63, header layout changes (rec.x), hex / zero /
    leading-zero header fields, N under '#' and under '!' quality, real bases under '!'."""
    rng = np.random.default_rng(seed)
    out: dict[str, bytes] = {}

    def rand_seq(n, alphabet=b"ACGT"):
        return np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)]

    def rand_q(n, lo=35, hi=74):
        return rng.integers(lo, hi, n).astype(np.uint8)

    # SOLiD: first base char and first quality char are prefixes (usrs.cpp:140-152,324-329,358-363)
    hs, ss, qs = [], [], []
    for i in range(300):
        hs.append(b"@ERR0489.%d solid0743_2011_PE_bc_%d_%d_%d/1" % (i + 1, 1 + i // 100, 2 + i // 7, 225 + 31 * i % 1700))
        pf = b"T" if i < 250 else b"G"
        body = rand_seq(50, b"0123")
        if i % 17 == 3:
            body[5:9] = ord(".")
        ss.append(np.concatenate((np.frombuffer(pf, dtype=np.uint8), body)))
        q = rand_q(50, 34, 70)
        if i % 17 == 3:
            q[5:7] = ord("!")
        qs.append(np.concatenate((np.frombuffer(b"!" if i < 280 else b"#", dtype=np.uint8), q)))
    out["solid"] = _assemble(hs, ss, qs)

    # 2nd id on the '+' line, variable length (454-like), N under '#' and '!' (usr.2id, usr.x, gen.Ns, gen.Nn)
    hs, ss, qs, ps = [], [], [], []
    for i in range(200):
        h = b"@SRR0012.%d FX9ABC01%s length=%d" % (i + 1, bytes(rand_seq(5, b"ABCDEFGH")), 0)
        L = int(rng.integers(40, 400))
        h = h[:-1] + b"%d" % L
        s = rand_seq(L)
        q = rand_q(L, 35, 74)
        if i % 5 == 0:
            s[3] = ord("N"); q[3] = ord("#")
            s[7] = ord("N"); q[7] = ord("!")
            q[11] = ord("!")
        hs.append(h); ss.append(s); qs.append(q); ps.append(b"+" + h[1:])
    out["twoid_varlen"] = _assemble(hs, ss, qs, ps)

    # qlen != llen, qualities >= 63 (escape), lower-case bases are NOT reproduced by the reference
    hs, ss, qs = [], [], []
    for i in range(120):
        hs.append(b"@run7_%d:%d:%d#0/1" % (3 + i // 50, 100 + i, 2000 - 3 * i))
        L = 60
        s = rand_seq(L)
        ql = L if i % 9 else L - 4 - i % 3
        q = rand_q(ql, 40, 110)
        hs[-1] = hs[-1]
        ss.append(s); qs.append(q)
    out["badqlt"] = _assemble(hs, ss, qs)

    # header model: layout change (rec.x), hex fields lower/upper, zero fields, leading zeros, >20 digits
    hs, ss, qs = [], [], []
    hexv = 0xfe12
    for i in range(160):
        hexv += int(rng.integers(-3, 40))
        if i % 40 == 39:
            h = b"@odd layout %d|%d" % (i, i * i)
        elif i % 4 == 0:
            h = b"@NB501:0:HV2:%d:%x:%X:0%d:%d:00%d" % (i % 3, hexv, hexv * 3, i % 11, 9 * 10**18 + i, i)
        else:
            h = b"@NB501:0:HV2:%d:%x:%X:0%d:%d:%d" % (i % 3, hexv, hexv * 3, i % 11, 123456789012345678 + i, 0 if i % 5 == 0 else i)
        hs.append(h)
        ss.append(rand_seq(36)); qs.append(rand_q(36, 35, 74))
    out["headers"] = _assemble(hs, ss, qs)

    # '+' line with trailing id that is only spaces -> usr.2id stays 0 but text is lossy in the reference;
    # tiny single-record and two-record files
    out["one_record"] = b"@r1\nACGTN\n+\nIIII#\n"
    out["two_records"] = b"@r.1 a:1\nACGTNACGT\n+\nIIII!IIII\n@r.2 a:2\nACGTTACGTAA\n+\nIIIIIIIII##\n"
    return out


def oversized_cases(seed: int = SEED0 + 11) -> dict[str, bytes]:
    """Inputs with oversized records (id of 8 191+ chars, base or quality line of 65 535+ chars: usrs.hpp:34-36,
    usrs.cpp:269-301), which the reference stores verbatim in usr.lrec / usr.lgen / usr.lqlt: in the middle of a
    file, at its start (determine_record, usrs.cpp:186-267), at the exact limits, with only the quality line too
    long (the length exception of the record is written before the record is put away), in a SOLiD file, alone."""
    rng = np.random.default_rng(seed)

    def rec(i, L, ql=None, hdr=None, alphabet=b"ACGT", pfx=b""):
        ql = L if ql is None else ql
        h = hdr if hdr is not None else b"@ont.%d ch=%d start=%d" % (i, 1 + i % 512, 1000 + 37 * i)
        s = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), L)].tobytes()
        q = rng.integers(35, 74, ql).astype(np.uint8).tobytes()
        return h + b"\n" + pfx + s + b"\n+\n" + (pfx and b"!") + q + b"\n"

    long_id = b"@" + b"x".join(b"%d" % (i * 7919) for i in range(1600))[:8700]
    out = {}
    out["oversize_mid"] = b"".join(
        [rec(i, 150) for i in range(20)] + [rec(20, 70000)] + [rec(i, 150) for i in range(21, 31)] + [rec(31, 100, hdr=long_id)] +
        [rec(i, 151) for i in range(32, 37)] + [rec(37, 65535), rec(38, 65534), rec(39, 150), rec(40, 65534, hdr=b"@" + b"y" * 8190),
                                               rec(41, 150, hdr=b"@" + b"z" * 8191), rec(42, 150)])
    out["oversize_first"] = b"".join([rec(0, 70000), rec(1, 90, hdr=long_id), rec(2, 65536)] + [rec(i, 120) for i in range(3, 30)])
    out["oversize_qual_only"] = b"".join([rec(i, 150) for i in range(8)] + [rec(8, 120, ql=66000)] + [rec(i, 150) for i in range(9, 16)] +
                                         [rec(16, 150, ql=65535)] + [rec(i, 150) for i in range(17, 20)])
    out["oversize_all"] = rec(0, 66000) + rec(1, 80, hdr=long_id)
    out["oversize_solid"] = b"".join([rec(i, 50, alphabet=b"0123", pfx=b"T") for i in range(12)] + [rec(12, 65535, alphabet=b"0123", pfx=b"T"),
                                      rec(13, 65534, alphabet=b"0123", pfx=b"G")] + [rec(i, 50, alphabet=b"0123", pfx=b"T") for i in range(14, 20)])
    return out
