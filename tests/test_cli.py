"""The host CLI keeps the reference's command-line surface (config.cpp:158-201, 239-379)."""
import os
import subprocess

import pytest

from slimfastq_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "slimfastq_b200", "bin", "slimfastq-b200")


def run(*args, **kw):
    return subprocess.run([CLI, *args], capture_output=True, **kw)


def test_version_help_and_missing_f():
    r = run("-v")
    assert r.returncode == 0 and r.stdout == b"Version 2.04\nInternal format version=6\n"      # config.cpp:268-270
    r = run("-h")
    assert r.returncode == 0 and b"-l level" in r.stdout and b"DWIM" in r.stdout
    r = run("-d")
    assert r.returncode == 1 and b"Missing essential argument: -f" in r.stderr                # config.cpp:326


@pytest.mark.gpu
def test_cli_roundtrip_pipes_dwim_and_overwrite_guard(tmp_path):
    data = synth.illumina(3000)
    fq, sfq, back = tmp_path / "a.fq", tmp_path / "a.sfq", tmp_path / "b.fq"
    fq.write_bytes(data)
    assert run("-u", str(fq), "-f", str(sfq), "-l", "3").returncode == 0
    assert sfq.read_bytes()[:16] == b"whoami=slimfastq"                                      # stamp, config.cpp:295
    r = run(str(sfq))                                                                         # DWIM: decode to stdout
    assert r.returncode == 0 and r.stdout == data
    assert run(str(sfq), str(back)).returncode == 0 and back.read_bytes() == data            # DWIM: decode to file
    r = run("-u", str(fq), "-f", str(sfq))                                                   # no -O: refuse to overwrite
    assert r.returncode == 1 and b"Can't write file" in r.stderr
    assert run("-u", str(fq), "-f", str(sfq), "-O", "-1").returncode == 0
    r = run("-f", str(tmp_path / "p.sfq"), "-4", input=data)                                 # stdin -> file
    assert r.returncode == 0
    r = run("-d", "-f", str(tmp_path / "p.sfq"))
    assert r.returncode == 0 and r.stdout == data
    r = run("-s", str(sfq))
    assert r.returncode == 0 and b"config.level     = 1" in r.stderr and b"num_records      = 3000" in r.stderr
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACXT\n+\nIIII\n")
    r = run("-u", str(bad), "-f", str(tmp_path / "bad.sfq"))
    assert r.returncode == 1 and b"slimfastq: encoding" in r.stderr and b"unexpected genome char: X" in r.stderr
    assert not (tmp_path / "bad.sfq").exists()


@pytest.mark.gpu
def test_cli_mixes_with_the_original_slimfastq(tmp_path):
    """-R writes the reference's own page format; a reference-written file is decoded without a flag."""
    from oracle import oracle as O

    if not O.have_ref():
        pytest.skip("oracle/_ref/slimfastq is not built")
    data = synth.illumina(2500)
    fq, ours, theirs, back = tmp_path / "a.fq", tmp_path / "ours.sfq", tmp_path / "theirs.sfq", tmp_path / "b.fq"
    fq.write_bytes(data)
    assert run("-R", "-u", str(fq), "-f", str(ours)).returncode == 0
    assert subprocess.run([O.REF_BIN, "-d", "-f", str(ours), "-u", str(back), "-O"]).returncode == 0   # theirs reads ours
    assert back.read_bytes() == data
    assert subprocess.run([O.REF_BIN, "-u", str(fq), "-f", str(theirs), "-O", "-q"]).returncode == 0
    r = run(str(theirs))                                                                              # ours reads theirs
    assert r.returncode == 0 and r.stdout == data
    r = run("-s", str(theirs))
    assert r.returncode == 0 and b"num_records      = 2500" in r.stderr


@pytest.mark.gpu
def test_cli_streams_in_bounded_segments(tmp_path):
    """-m: the input is coded in segments (each cut on a record boundary, carry kept), blobs appended as they
    come, one index at the end; decoding walks groups of chunks.  Same records, same order, same bytes back."""
    from slimfastq_b200 import container as K

    data = synth.illumina(9000) + synth.ont(30) + synth.illumina(2000, seed=77)       # ~5 MB, mixed record sizes
    fq, sfq, back = tmp_path / "a.fq", tmp_path / "a.sfq", tmp_path / "b.fq"
    fq.write_bytes(data)
    r = run("-m", "1", "-c", "262144", "-f", str(sfq), input=data)                  # stdin, 1 MiB segments
    assert r.returncode == 0, r.stderr
    ct = K.parse(sfq.read_bytes())
    assert ct.orig_size == len(data) and sum(c.text_len for c in ct.chunks) == len(data) and len(ct.chunks) > 8
    assert run("-m", "0", "-c", "262144", "-u", str(fq), "-f", str(tmp_path / "oneshot.sfq")).returncode == 0
    assert sfq.read_bytes() == (tmp_path / "oneshot.sfq").read_bytes()              # parts on the chunk grid: same container
    r = run("-m", "1", "-d", "-f", str(sfq))                                         # grouped decode to stdout
    assert r.returncode == 0 and r.stdout == data
    assert run("-m", "0", str(sfq), str(back)).returncode == 0 and back.read_bytes() == data   # ... equals one-shot decode
    r = run("-m", "1", "-f", str(tmp_path / "t.sfq"), input=data[:-7])              # truncated last record: croak, no file
    assert r.returncode == 1 and b"truncated" in r.stderr and not (tmp_path / "t.sfq").exists()
