"""tools/slimfastq-b200.multi keeps the command line of the reference's tools/slimfastq.multi."""
import os
import subprocess
import sys

import pytest

from slimfastq_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MULTI = os.path.join(ROOT, "tools", "slimfastq-b200.multi")


def run(*args):
    return subprocess.run([sys.executable, MULTI, *args], capture_output=True, text=True)


def test_dry_run_names_targets_like_the_reference(tmp_path):
    src = tmp_path / "FQ"
    (src / "sub").mkdir(parents=True)
    for name in ("a.fastq", "b.fq", "c.txt", "sub/d.fq"):
        (src / name).write_text("@r\nA\n+\nI\n")
    tgt = tmp_path / "SFQ"
    r = run("-n", "-e", "slimfastq-b200", "-g", "0,1", "-t", str(tgt), str(src))
    assert r.returncode == 0 and tgt.is_dir()
    cmds = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("slimfastq-b200")]
    assert sorted(c[c.index("-f") + 1] for c in cmds) == [str(tgt / "a.sfq"), str(tgt / "b.sfq")]     # suffix list, no recursion
    assert all("-O" in c and "-u" in c for c in cmds)                                                   # compress overwrites, as the reference's driver does
    assert {c[c.index("-g") + 1] for c in cmds} <= {"0", "1"}
    r = run("-n", "-e", "x", "-r", "-g", "0", str(src))
    assert sum("d.sfq" in ln for ln in r.stdout.splitlines()) == 1                                      # -r descends
    good = tmp_path / "z.sfq"
    good.write_bytes(b"whoami=slimfastq\n" + b"\0" * 64)
    (tmp_path / "bad.sfq").write_bytes(b"not one")
    r = run("-n", "-d", "-e", "x", "-g", "3", "-f", ".fq2,.fq", str(tmp_path))
    lines = r.stdout.splitlines()
    assert any(ln.startswith("x -g 3 -d -u " + str(tmp_path / "z.fq2")) for ln in lines)                # first suffix of the list
    assert not any(ln.startswith("x ") and " -O" in ln for ln in lines)                                  # decompress never overwrites
    assert any("bad.sfq doesn't seem to be a valid slimfastq file" in ln for ln in lines)
    assert run("-n", str(tmp_path / "nothing-here")).returncode != 0


@pytest.mark.gpu
def test_batch_round_trip(tmp_path):
    src, mid, back = tmp_path / "FQ", tmp_path / "SFQ", tmp_path / "BACK"
    src.mkdir()
    data = {f"s{i}.fq": synth.illumina(1500, seed=100 + i) for i in range(3)}
    for k, v in data.items():
        (src / k).write_bytes(v)
    assert run("-g", "0", "-c", "2", "-t", str(mid), str(src)).returncode == 0
    assert run("-d", "-g", "0", "-t", str(back), "-f", ".fq", str(mid)).returncode == 0
    for k, v in data.items():
        assert (back / k).read_bytes() == v


@pytest.mark.gpu
def test_workers_keep_one_context_and_never_overwrite_on_decompress(tmp_path):
    """The in-process path (one Codec per worker), the CLI path (-e), and the reference's rule that an existing FASTQ is
    not overwritten on decompress."""
    src, mid, back = tmp_path / "FQ", tmp_path / "SFQ", tmp_path / "BACK"
    src.mkdir(); back.mkdir()
    data = {f"s{i}.fq": synth.illumina(1200, seed=200 + i) for i in range(4)}
    for k, v in data.items():
        (src / k).write_bytes(v)
    assert run("-g", "0", "-c", "1", "-t", str(mid), str(src)).returncode == 0                        # one worker, four files, one context
    cli = os.path.join(ROOT, "slimfastq_b200", "bin", "slimfastq-b200")
    via_cli = tmp_path / "SFQ2"
    assert run("-g", "0", "-c", "1", "-e", cli, "-t", str(via_cli), str(src)).returncode == 0
    for k in data:
        assert (mid / k.replace(".fq", ".sfq")).read_bytes() == (via_cli / k.replace(".fq", ".sfq")).read_bytes()
    (back / "s0.fq").write_bytes(b"precious")
    r = run("-d", "-g", "0", "-t", str(back), "-f", ".fq", str(mid))
    assert r.returncode != 0 and (back / "s0.fq").read_bytes() == b"precious"
    for k, v in data.items():
        if k != "s0.fq":
            assert (back / k).read_bytes() == v


@pytest.mark.gpu
def test_batch_over_two_gpus(tmp_path):
    """-g 0,1: one worker (one context) per device, files from one shared queue."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    src, mid, back = tmp_path / "FQ", tmp_path / "SFQ", tmp_path / "BACK"
    src.mkdir()
    data = {f"s{i}.fq": synth.illumina(1500, seed=300 + i) for i in range(6)}
    for k, v in data.items():
        (src / k).write_bytes(v)
    r = run("-v", "-g", "0,1", "-c", "2", "-t", str(mid), str(src))
    assert r.returncode == 0
    assert run("-d", "-g", "0,1", "-c", "2", "-t", str(back), "-f", ".fq", str(mid)).returncode == 0
    for k, v in data.items():
        assert (back / k).read_bytes() == v
