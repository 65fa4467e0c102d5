#!/usr/bin/env python
"""Benchmark of the slimfastq hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gb G] [--level L] [--chunk B]

One step = one compress pass + one decompress pass over the workload (BASELINE.json configs[1]:
synthetic Illumina 2x150, 40-level Phred, 10 GB, level 3, 1 MiB chunks).  `value` is FASTQ GB
moved through the codec per second with inputs resident in HBM (2 x workload / (t_c + t_d));
`e2e` is the same through sfq_compress/sfq_decompress with pinned HOST buffers, copies inside the
timed region.  N > 1: every rank codes its own shard of chunks (weak scaling, no collective on
the data path; only sizes are exchanged), launched by torchrun.  `--single-file` instead cuts ONE
file on the chunk grid across the ranks and merges the rank containers (strong scaling).

At N = 1 the default run also reports, outside the timed headline region:
  parity_sampled   chunks of the benched container compared with the oracle
  chain            link time of every coder chain at the benched residency and with one resident warp
  chunk_pareto     throughput and ratio loss against the reference's whole-file ratio per chunk size
  configs          short legs for BASELINE configs[2] (level 4, and its 8-bin variant) and configs[3] (ONT)
  cpu_baseline     the unmodified reference on the host cores: over the same chunks (tools/slimfastq.multi),
                   and its best case (one whole-file process per core)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import random
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fastq_compress_decompress_throughput"
UNIT = "GB/s"
UNIQUE_READS = 360_000          # ~129 MB unique synthetic block, tiled to the workload size
ONT_READS = 6400                # ~129 MB of ONT-style reads (log-normal 1-50 kb), tiled like the Illumina block
TRAFFIC_PROFILE = os.path.join(ROOT, "profiles", "r2_traffic_10gb.json")    # written by tools/ncu_traffic.py from an ncu run of this file


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gb", type=float, default=10.0, help="workload size per GPU in decimal GB")
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=1 << 20)
    ap.add_argument("--bins8", action="store_true", help="Illumina 8-bin quantised qualities (BASELINE configs[2] variant)")
    ap.add_argument("--workload", default="illumina", choices=["illumina", "ont"],
                    help="illumina = BASELINE configs[1] (the metric's config); ont = configs[3], long reads 1-50 kb with N runs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip parity sample, chain probe, chunk Pareto and the configs[2]/[3] legs")
    ap.add_argument("--pareto", default="2,4,7,8", help="chunk sizes (MiB) of the chunk_pareto legs besides --chunk ('' = none)")
    ap.add_argument("--single-file", action="store_true", help="one file cut on the chunk grid across the ranks, containers merged (strong scaling)")
    ap.add_argument("--cpu-sample-mb", type=int, default=0, help="CPU sample per pass in MiB (0 = sized from --ref-budget-s)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="wall-clock budget of the whole --impl reference run")
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu,utilization.memory")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons, ug, um, pw = [], 0, set(), [], [], []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                try:
                    pw.append(float(r[3])); ug.append(float(r[9])); um.append(float(r[10]))
                except (ValueError, IndexError):
                    pass
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None,
                "nvidia_smi_utilization_pct": {"gpu": statistics.median(ug) if ug else None, "memory": statistics.median(um) if um else None}}


# ------------------------------------------------------------------------------ CPU reference legs
class ReferenceRunner:
    """The UNMODIFIED reference (oracle/_ref/slimfastq) over a set of FASTQ files on tmpfs, one process per file on
    `cores` workers through the reference's own tools/slimfastq.multi (which cannot pass a level, hence the two-line
    wrapper).  The input files are written once; every pass() compresses all of them and decompresses the results."""

    def __init__(self, level: int, cores: int):
        from oracle import oracle as O

        if not O.have_ref():
            raise RuntimeError("oracle/_ref/slimfastq is not built")
        self.O, self.level, self.cores = O, level, cores
        self.dir = tempfile.mkdtemp(prefix="sfqref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.wrapper = os.path.join(self.dir, "sfq_l.sh")
        with open(self.wrapper, "w") as f:
            f.write(f"#!/bin/sh\nexec {O.REF_BIN} -l {level} \"$@\"\n")
        os.chmod(self.wrapper, 0o755)
        self.multi = os.path.join(os.path.dirname(O.REF_BIN), "tools", "slimfastq.multi")
        self.sets = {}
        self.how = "tools/slimfastq.multi"

    def add_files(self, name: str, data: bytes, lens: list[int], batch_files: int = 0):
        """The files of a set live in batch directories (one tools/slimfastq.multi run each): a pass that is out of time
        stops between batches, so no pass can overrun its deadline by more than one batch."""
        d = os.path.join(self.dir, name)
        os.mkdir(d)
        batch_files = batch_files or max(1, len(lens))
        pos, batches = 0, []
        for i, ln in enumerate(lens):
            if i % batch_files == 0:
                bd = os.path.join(d, f"b{len(batches):04d}")
                os.mkdir(bd)
                batches.append([bd, 0, 0])
            with open(os.path.join(batches[-1][0], f"c{i:06d}.fq"), "wb") as f:
                f.write(data[pos:pos + ln])
            pos += ln
            batches[-1][1] += 1
            batches[-1][2] += ln
        self.sets[name] = batches

    def _batch(self, src: str, nfiles: int, comp: str, back: str):
        for x in (comp, back):
            shutil.rmtree(x, ignore_errors=True); os.mkdir(x)
        ok = False
        tc = td = 0.0
        if self.how == "tools/slimfastq.multi":
            t0 = time.perf_counter()
            r = subprocess.run(["perl", self.multi, "-c", str(self.cores), "-e", self.wrapper, "-t", comp, src], capture_output=True, text=True)
            tc = time.perf_counter() - t0
            ok = r.returncode == 0 and len(os.listdir(comp)) == nfiles
            if ok:
                t0 = time.perf_counter()
                r = subprocess.run(["perl", self.multi, "-d", "-c", str(self.cores), "-e", self.wrapper, "-t", back, comp], capture_output=True, text=True)
                td = time.perf_counter() - t0
                ok = r.returncode == 0 and len(os.listdir(back)) == nfiles
        if not ok:                                   # same work without perl ithreads
            from concurrent.futures import ThreadPoolExecutor

            self.how = "thread pool of reference processes"
            for x in (comp, back):
                shutil.rmtree(x); os.mkdir(x)
            paths = sorted(os.path.join(src, f) for f in os.listdir(src))

            def enc(p):
                subprocess.run([self.O.REF_BIN, "-l", str(self.level), "-q", "-O", "-u", p, "-f", os.path.join(comp, os.path.basename(p) + ".sfq")], check=True)

            def dec(p):
                subprocess.run([self.O.REF_BIN, "-d", "-O", "-f", os.path.join(comp, os.path.basename(p) + ".sfq"), "-u", os.path.join(back, os.path.basename(p))], check=True)

            with ThreadPoolExecutor(self.cores) as ex:
                t0 = time.perf_counter(); list(ex.map(enc, paths)); tc = time.perf_counter() - t0
                t0 = time.perf_counter(); list(ex.map(dec, paths)); td = time.perf_counter() - t0
        shutil.rmtree(back, ignore_errors=True)
        shutil.rmtree(comp, ignore_errors=True)
        return tc, td

    def pass_(self, name: str, deadline_s: float = 0.0) -> dict:
        """Compress then decompress the set, batch by batch; with a deadline the pass ends after the first batch that
        crosses it and reports the bytes it did code."""
        comp, back = os.path.join(self.dir, name + ".sfq"), os.path.join(self.dir, name + ".back")
        tc = td = 0.0
        nbytes = done = 0
        for src, nfiles, nb in self.sets[name]:
            a, b = self._batch(src, nfiles, comp, back)
            tc += a; td += b; nbytes += nb; done += 1
            if deadline_s and tc + td >= deadline_s:
                break
        return {"t_compress": tc, "t_decompress": td, "bytes": nbytes, "value": 2 * nbytes / (tc + td) / 1e9,
                "compress_GBps": nbytes / tc / 1e9, "decompress_GBps": nbytes / td / 1e9, "batches": done, "of_batches": len(self.sets[name])}

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def _spread(vals: list[float]) -> dict:
    return {"median": round(statistics.median(vals), 5), "min": round(min(vals), 5), "max": round(max(vals), 5), "passes": len(vals)}


def reference_timed(block: bytes, level: int, chunk: int, cores: int, steps: int, warmup: int, budget_s: float, sample_mb: int = 0) -> dict:
    """W warm + K timed passes of the reference over the product's chunks of a bounded sample of the workload.  The
    sample is sized from a calibration pass so that the whole run fits `budget_s` at any core count (the first pass
    of a fresh box pages perl and the binary in and is several times slower: it is always discarded)."""
    from slimfastq_b200.api import chunk_lengths, record_start_at_or_after

    def sample_of(nbytes: int):
        reps = (nbytes + len(block) - 1) // len(block)
        s = block * max(1, reps)
        s = s[: record_start_at_or_after(s, min(len(s), nbytes))] or s
        return s, chunk_lengths(s, chunk)        # the product's chunking rule (sfq_plan.cuh)

    t_begin = time.perf_counter()
    run = ReferenceRunner(level, cores)
    try:
        cal, cal_lens = sample_of(max(2 * cores, 8) * chunk)
        run.add_files("cal", cal, cal_lens)
        run.pass_("cal")                                            # discarded: page-in
        c = run.pass_("cal")
        rate = c["bytes"] / (c["t_compress"] + c["t_decompress"])    # bytes per second of (compress + decompress)
        npass = max(1, steps + warmup)
        left = max(5.0, budget_s - (time.perf_counter() - t_begin) - 5.0)
        per_pass = left / npass                                      # seconds a pass may take: enforced, not hoped for
        want = (sample_mb << 20) if sample_mb else int(rate * per_pass * 0.8)
        want = max(len(cal), min(want, 512 << 20))     # (2 GiB passes showed 2x pass-to-pass swings on the 16-core box - 300 GB of table
                                                       # zeroing per pass in the kernel's page allocator - where ~400 MiB passes stay within 1 %)
        sample, lens = sample_of(want)
        # batches of ~4 s of work: a pass stops at the first batch boundary past its deadline (a pass of the reference is
        # now and then several times slower than its neighbours - page zeroing of its 77 MB of tables per process, tmpfs
        # reclaim - and the calibration cannot know that)
        per_batch = max(2 * cores, min(len(lens), int(rate * 4.0 / max(1, chunk))))
        run.add_files("main", sample, lens, per_batch)
        for _ in range(warmup):
            run.pass_("main", per_pass)
        res = [run.pass_("main", per_pass) for _ in range(steps)]
        tot = sum(r["t_compress"] + r["t_decompress"] for r in res)
        coded = sum(r["bytes"] for r in res)
        vals = [r["value"] for r in res]
        return {"value": statistics.median(vals), "value_total": 2 * coded / tot / 1e9, "ms_per_step": tot / len(res) * 1e3,
                "compress_GBps": _spread([r["compress_GBps"] for r in res]), "decompress_GBps": _spread([r["decompress_GBps"] for r in res]),
                "value_spread": _spread(vals), "sample_bytes": max(r["bytes"] for r in res), "files": len(lens), "how": run.how,
                "batches_done": [r["batches"] for r in res], "batches": res[0]["of_batches"],
                "calibration_GBps": round(c["value"], 5), "wall_s": round(time.perf_counter() - t_begin, 1)}
    finally:
        run.close()


def reference_best_case(block: bytes, level: int, cores: int, file_mb: int = 48) -> dict:
    """the reference's best case - one WHOLE-FILE process per core (no per-chunk
    start-up, warm models), `cores` files of `file_mb` MiB through tools/slimfastq.multi."""
    from slimfastq_b200.api import record_start_at_or_after

    s = block[: record_start_at_or_after(block, min(len(block), file_mb << 20))] or block
    run = ReferenceRunner(level, cores)
    try:
        run.add_files("whole", s * cores, [len(s)] * cores)
        r = run.pass_("whole")
        return {"value": round(r["value"], 5), "unit": UNIT, "compress_GBps": round(r["compress_GBps"], 5), "decompress_GBps": round(r["decompress_GBps"], 5),
                "sample": "%d whole files of %d MiB, one reference process per core via %s (one pass)" % (cores, len(s) >> 20, run.how)}
    finally:
        run.close()


def reference_whole_file(block: bytes, level: int) -> dict:
    """Whole-file run of the reference on the unique block: its stream bytes are the ratio yardstick."""
    from oracle import oracle as O

    t0 = time.perf_counter()
    enc = O.ref_encode(block, level, tmpdir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    dt = time.perf_counter() - t0
    sb = sum(len(v) for v in enc.streams.values())
    return {"stream_bytes": sb, "ratio": len(block) / sb, "one_core_compress_MBps": len(block) / dt / 1e6}


# ------------------------------------------------------------------------------ workload
def workload_block(kind: str, rank: int, bins8: bool = False) -> bytes:
    from slimfastq_b200 import synth

    if kind == "ont":
        return synth.ont(ONT_READS, seed=synth.SEED0 + 4 + 1000 * rank)
    return synth.illumina(UNIQUE_READS, seed=synth.SEED0 + 1 + 1000 * rank, bins8=bins8)


def workload_label(kind: str, bins8: bool, gb: float) -> str:
    """The same string in both arms (nominal size; the exact byte count is config.bytes_per_gpu)."""
    if kind == "ont":
        return "ont_1-50kb_%gGB_per_gpu" % gb
    return "illumina_2x150_%s_%gGB_per_gpu" % ("8bin" if bins8 else "phred40", gb)


class Bench:
    """Device-resident and end-to-end timing of one (workload, level, chunk size) on this rank's GPU."""

    def __init__(self, torch, dist, codec, local: int, world: int):
        self.torch, self.dist, self.codec, self.local, self.world = torch, dist, codec, local, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def upload(self, block: bytes, tiles: int):
        torch = self.torch
        d_block = torch.frombuffer(bytearray(block), dtype=torch.uint8).cuda()
        d_text = d_block.repeat(tiles) if tiles > 1 else d_block
        return d_text

    def device_steps(self, d_text, level: int, chunk: int, steps: int, warmup: int, clocks: bool = False, parity_block=None) -> dict:
        torch, codec = self.torch, self.codec
        n = d_text.numel()
        d_sfq = torch.empty(codec.compress_bound(n, chunk), dtype=torch.uint8, device="cuda")
        d_back = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()

        def step():
            csz = codec.compress_device(d_text, d_sfq, level, chunk)
            sc = codec.stats()
            osz = codec.decompress_device(d_sfq, csz, d_back)
            sd = codec.stats()
            assert osz == n
            return csz, sc, sd

        for _ in range(max(1, warmup)):
            csz, sc, sd = step()
        assert torch.equal(d_back[:n], d_text), "device round trip differs"
        parity = sample_parity(torch, d_text, d_sfq, csz, level, 8) if parity_block is not None else None
        sampler = ClockSampler(self.local) if clocks else None
        self.barrier()
        if sampler:
            sampler.start()
        acc = {k: 0.0 for k in ("t_c", "t_d", "launches", "code_c", "code_d", "scan", "waves_c", "waves_d", "c_gen", "c_qlt", "c_rec", "d_gen", "d_qlt", "d_rec")}
        per_step = []
        t0 = time.perf_counter()
        for _ in range(steps):
            csz, sc, sd = step()
            acc["t_c"] += sc["ms_total"]; acc["t_d"] += sd["ms_total"]
            acc["launches"] += sc["kernel_launches"] + sd["kernel_launches"]
            acc["code_c"] += sc["ms_code"]; acc["code_d"] += sd["ms_code"]; acc["scan"] += sc["ms_scan"]
            acc["waves_c"] += sc["waves"]; acc["waves_d"] += sd["waves"]
            for k in ("gen", "qlt", "rec"):
                acc["c_" + k] += sc["ms_" + k]; acc["d_" + k] += sd["ms_" + k]
            per_step.append([round(sc["ms_code"], 1), round(sd["ms_gen"], 1), round(sd["ms_qlt"], 1), round(sd["ms_rec"], 1)])
        self.barrier()
        wall = time.perf_counter() - t0
        out = {"n": n, "csz": csz, "sc": sc, "sd": sd, "acc": acc, "per_step": per_step, "wall": wall, "steps": steps,
               "clocks": sampler.stop() if sampler else None, "parity": parity}
        del d_sfq, d_back
        torch.cuda.empty_cache()
        return out

    def e2e_steps(self, d_text, csz_hint: int, level: int, chunk: int, steps: int, warmup: int):
        """Host buffers in, host buffers out, copies inside the timed region.  Frees nothing of the caller's."""
        torch, codec, S = self.torch, self.codec, sys.modules["slimfastq_b200"]
        n = d_text.numel()
        err, res = None, None
        try:
            need = self.world * (2 * n + csz_hint + (1 << 29))
            try:
                import psutil  # noqa: PLC0415

                avail = psutil.virtual_memory().available
            except ImportError:
                avail = None
            if avail is not None and need > 0.85 * avail:
                raise MemoryError("e2e leg needs %.0f GB of pinned host memory on this node, %.0f GB available" % (need / 1e9, avail / 1e9))
            h_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            h_text.copy_(d_text)
            torch.cuda.synchronize()

            def step():
                a, cn = codec.compress_view(h_text, level, chunk)        # container in the context's pinned buffer
                s1 = codec.stats()
                b, on = codec.decompress_addr(a, cn)                     # (decompress returns into a buffer of its own)
                s2 = codec.stats()
                return cn, on, s1, s2, b

            for _ in range(max(1, min(warmup, 3))):
                cn, on, s1, s2, b = step()
            assert on == n
            m = min(n, 64 << 20)
            assert ctypes.string_at(b, m) == bytes(h_text[:m].numpy()), "e2e round trip differs"
            tail = max(0, n - m)
            assert ctypes.string_at(b + tail, n - tail) == bytes(h_text[tail:].numpy()), "e2e round trip differs (tail)"
        except (RuntimeError, S.SfqError, MemoryError) as ex:      # e.g. not enough pinned host memory for N ranks
            err = str(ex)[:200]
        ok = 0.0 if err else 1.0                                   # every rank must take the same path through the barriers
        if self.world > 1:
            tok = torch.tensor([ok], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(tok, op=self.dist.ReduceOp.MIN)
            ok = float(tok.item())
        if not ok:
            return None, err or "another rank could not set up its host buffers"
        self.barrier()
        e_ms = e_copy = e_hidden = 0.0
        e_steps, launches, parts = [], 0, {}
        t0 = time.perf_counter()
        for _ in range(steps):
            cn, on, s1, s2, b = step()
            e_ms += s1["ms_total"] + s2["ms_total"]
            e_copy += s1["ms_h2d"] + s1["ms_d2h"] + s2["ms_h2d"] + s2["ms_d2h"]
            e_steps.append([round(s1["ms_total"], 1), round(s1["ms_code"], 1), round(s2["ms_total"], 1), round(s2["ms_gen"], 1), round(s2["ms_qlt"], 1), round(s2["ms_rec"], 1)])
            parts = {"c_total": s1["ms_total"], "c_h2d": s1["ms_h2d"], "c_code": s1["ms_code"], "c_d2h": s1["ms_d2h"], "c_waves": s1["waves"],
                     "d_total": s2["ms_total"], "d_h2d": s2["ms_h2d"], "d_code": s2["ms_code"], "d_d2h": s2["ms_d2h"], "d_waves": s2["waves"],
                     "d_resident": s2["resident_chunks"],
                     "c_scan_plan_clear_pack": s1["ms_scan"] + s1["ms_plan"] + s1["ms_clear"] + s1["ms_pack"],
                     "d_plan_clear_pack": s2["ms_plan"] + s2["ms_clear"] + s2["ms_pack"]}
            launches += s1["kernel_launches"] + s2["kernel_launches"]
        self.barrier()
        res = {"ms": e_ms, "wall_s": time.perf_counter() - t0, "h2d": n + cn, "d2h": cn + n, "copy_ms": e_copy, "steps": e_steps,
               "launches": launches, "parts": {k: round(float(v), 2) for k, v in parts.items()}}
        del h_text
        return res, None


def sample_parity(torch, d_text, d_sfq, csz: int, level: int, k: int) -> dict:
    """k random chunks of the container just written by the timed configuration against the oracle (== the reference run
    on that chunk as a standalone file): every stream and info key, bit for bit.  Outside the timed region."""
    from oracle import oracle as O
    from slimfastq_b200 import container as K

    fh = K.FILE_HDR.unpack_from(bytes(d_sfq[:K.FILE_HDR.size].cpu().numpy()), 0)
    nchunks, index_off = fh[5], fh[7]
    offs = d_sfq[index_off:index_off + 8 * nchunks].clone().view(torch.int64)
    tl = d_sfq[(offs[:, None] + torch.arange(8, 16, device=offs.device)[None, :]).reshape(-1)].view(torch.int64)   # text_len of every blob
    starts = (torch.cumsum(tl, 0) - tl).cpu().tolist()
    offs_h = offs.cpu().tolist() + [index_off]
    rng = random.Random(0x5F51)
    picks = sorted(rng.sample(range(nchunks), min(k, nchunks)))
    for c in picks:
        blob = bytes(d_sfq[offs_h[c]:offs_h[c + 1]].cpu().numpy())
        ch = K.parse_blob(blob, 0)
        text = bytes(d_text[starts[c]:starts[c] + ch.text_len].cpu().numpy())
        o = O.encode(text, level)
        assert o.info_tuple() == ch.info_tuple(), "chunk %d of the benched container: info keys differ from the oracle" % c
        assert o.streams == ch.streams, "chunk %d of the benched container: streams differ from the oracle" % c
    return {"chunks": picks, "of": nchunks, "checker": "oracle/sfq_oracle.c (pinned to the reference binary), all streams and info keys"}


def chain_numbers(sc: dict, sd: dict, acc: dict, steps: int) -> dict:
    """ns per link of every serial chain at the run's residency (a link = one coded symbol of a chunk-stream;
    header chains: one record).  Kernels of a wave run concurrently, each timed with events on its own stream."""
    nch = max(1, sc["nchunks"])
    wc, wd = max(1, sc["waves"]), max(1, sd["waves"])
    sym_b, sym_q, recs = sc["nbases"] / nch, sc["nquals"] / nch, sc["nrecords"] / nch

    def ns(ms, per, waves):
        return round(ms / steps / waves * 1e6 / max(1.0, per), 2)

    return {"symbols_per_chunk_stream": round((sym_b + sym_q) / 2), "records_per_chunk": round(recs),
            "compress": {"resident_chunks": sc["resident_chunks"], "waves": sc["waves"],
                         "gen_path_ns_per_base": ns(acc["c_gen"], sym_b, wc), "qlt_path_ns_per_quality": ns(acc["c_qlt"], sym_q, wc),
                         "rec_ns_per_record": ns(acc["c_rec"], recs, wc)},
            "decompress": {"resident_chunks": sd["resident_chunks"], "waves": sd["waves"],
                           "gen_ns_per_base": ns(acc["d_gen"], sym_b, wd), "qlt_ns_per_quality": ns(acc["d_qlt"], sym_q, wd),
                           "rec_ns_per_record": ns(acc["d_rec"], recs, wd)}}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        run_reference_arm(args, cores)
        return 0

    import torch
    import torch.distributed as dist

    import slimfastq_b200 as S
    from slimfastq_b200 import container as K

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # pinned staging buffers on the GPU's own NUMA node (first touch): N ranks then do not all copy through node 0
    from slimfastq_b200.api import bind_to_device_node
    numa_node, all_cpus = bind_to_device_node(local)
    codec = S.Codec(local)
    B = Bench(torch, dist, codec, local, world)
    if args.single_file:
        rc = run_single_file(args, B, rank, world)
        codec.close()
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return rc

    nbytes = int(args.gb * 1e9)
    block = workload_block(args.workload, rank, args.bins8)
    tiles = max(1, round(nbytes / len(block)))
    extras = world == 1 and not args.no_extras
    d_text = B.upload(block, tiles)
    n = d_text.numel()

    # ---------------- device-resident steps (value)
    dev = B.device_steps(d_text, args.level, args.chunk, args.steps, args.warmup, clocks=True, parity_block=block if extras else None)
    sc, sd, acc, csz = dev["sc"], dev["sd"], dev["acc"], dev["csz"]
    t_c, t_d = acc["t_c"], acc["t_d"]
    dev_ms = t_c + t_d            # CUDA-event time on the library's stream, summed over the K steps
    launches = acc["launches"]

    # ---------------- end-to-end steps (host buffers, copies inside the timed region)
    e2e = e2e_err = None
    if not args.no_e2e:
        e2e, e2e_err = B.e2e_steps(d_text, csz, args.level, args.chunk, args.steps, args.warmup)
        if e2e:
            launches += e2e["launches"]

    # ---------------- reduce over ranks: max time, sum bytes
    def allred(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(x):
        return allred(x, dist.ReduceOp.MAX)

    def allsum(x):
        return allred(x, dist.ReduceOp.SUM)

    tot_bytes = allsum(float(n))
    tot_csz = allsum(float(csz))
    dev_ms_max = allmax(dev_ms)
    t_c_max, t_d_max = allmax(t_c), allmax(t_d)
    wall_max = allmax(dev["wall"])
    e2e_wall_max = allmax(e2e["wall_s"]) if e2e else None
    launches_sum = allsum(float(launches))

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        K_ = args.steps
        value = 2 * tot_bytes * K_ / (dev_ms_max / 1e3) / 1e9
        # Dominant kernel = the quality decoder (k_qlt_decode<LPC>, the longest launch of a step).  Algorithmic bytes
        # per launch: one byte out per quality + the qlt stream bytes in.  Its duration is
        # measured with CUDA events on the stream it is launched on (sfq_stats.ms_qlt of the decompress call).
        qd_ms = acc["d_qlt"]
        qd_bytes = sd["nquals"] + sd["qlt_stream_bytes"]
        achieved = qd_bytes * K_ / (qd_ms / 1e3) / 1e9
        hdr_bytes = n - sc["nbases"] - sc["nquals"] - 6 * sc["nrecords"]
        lpc = 4 if sd["resident_chunks"] >= 4096 else 8
        kname = "k_qlt_decode<%d>" % lpc
        traffic, traffic_src = None, None
        try:
            tp = json.load(open(TRAFFIC_PROFILE))
            kk = next((v for k, v in tp["kernels"].items() if k.startswith("k_qlt_decode<%d" % lpc)), None)   # (ncu prints every template argument)
            same = tp["config"]["chunk_bytes"] == args.chunk and tp["config"]["level"] == args.level and abs(tp["config"]["nchunks"] - sc["nchunks"]) <= 0.02 * sc["nchunks"]
            if kk and same:
                traffic = int(kk["dram_bytes_read"] + kk["dram_bytes_write"])
                traffic_src = "%s (%s)" % (os.path.relpath(TRAFFIC_PROFILE, ROOT), tp["how"])
        except (OSError, KeyError, ValueError):
            pass

        def gbps(nb, ms):
            return round(nb / (ms / 1e3) / 1e9, 3) if ms > 0 else None

        coder_kernels = {   # every coder kernel group of one step against its own algorithmic bytes (last step's timings)
            "compress_gen": gbps(sc["nbases"] + sc["gen_stream_bytes"], sc["ms_gen"]),
            "compress_qlt": gbps(sc["nquals"] + sc["qlt_stream_bytes"], sc["ms_qlt"]),
            "compress_rec": gbps(hdr_bytes, sc["ms_rec"]),
            "decompress_gen": gbps(sd["nbases"] + sd["gen_stream_bytes"], sd["ms_gen"]),
            "decompress_qlt": gbps(sd["nquals"] + sd["qlt_stream_bytes"], sd["ms_qlt"]),
            "decompress_rec": gbps(hdr_bytes, sd["ms_rec"]),
        }
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": args.warmup,
            "ms_per_step": round(dev_ms_max / K_, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic (%s block of %d MB generated by slimfastq_b200/synth.py, tiled x%d: every %dth chunk repeats content; chunks are coded independently, so neither ratio nor per-chunk work changes)"
                                    % ("ONT-style long-read" if args.workload == "ont" else "Markov-quality Illumina", len(block) // 10**6, tiles, max(1, round(len(block) / args.chunk))),
            "config": {"workload": workload_label(args.workload, args.bins8, args.gb),
                       "level": args.level, "chunk_bytes": args.chunk, "bytes_per_gpu": n, "l2": "inputs (%.1f GB) exceed L2 (126 MB); no flush needed" % (n / 1e9),
                       "step": "compress + decompress", "sharding": "chunks by rank, no data-path collective",
                       "host_numa_node": numa_node},
            "compress_GBps": round(tot_bytes * K_ / (t_c_max / 1e3) / 1e9, 4),
            "decompress_GBps": round(tot_bytes * K_ / (t_d_max / 1e3) / 1e9, 4),
            "ratio": round(tot_bytes / tot_csz, 4), "stream_ratio": round(n / sc["stream_bytes"], 4),
            "wall_s_per_step": round(wall_max / K_, 4),
            "clocks": dev["clocks"],
            "gpu_launches": int(launches_sum),
            "roofline": {"bound": "hbm", "kernel": "%s (quality decoder: %d lanes per chunk, %d chunks per warp, warp-converged)" % (kname, lpc, 32 // lpc),
                         "achieved": round(achieved, 3), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 6),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes_per_launch": int(qd_bytes / max(1, sd["waves"])),
                         "launch_ms_avg": round(qd_ms / K_ / max(1, sd["waves"]), 3),
                         "note": "HBM is the contract's denominator; what bounds this kernel is its serial chain (one adaptive-coder link per quality per chunk): see chain.frac"},
            "coder_kernels_GBps": coder_kernels,
            "per_step_ms[c_code,d_gen,d_qlt,d_rec]": dev["per_step"],
            "chain": chain_numbers(sc, sd, acc, K_),
            "roofline_scan": {"bound": "hbm", "kernel": "k_count_newlines+k_scan_tiles+k_fill_lines", "achieved": round((n + 8 * 4 * sc["nrecords"]) * K_ / (acc["scan"] / 1e3) / 1e9, 2),
                              "peak": hbm_peak, "unit": "GB/s", "frac": round((n + 32 * sc["nrecords"]) * K_ / (acc["scan"] / 1e3) / 1e9 / hbm_peak, 4)},
            "phases_ms_per_step": {"c_scan": round(acc["scan"] / K_, 3), "c_code": round(acc["code_c"] / K_, 3), "d_code": round(acc["code_d"] / K_, 3),
                                   "c_total": round(t_c / K_, 3), "d_total": round(t_d / K_, 3), "c_clear": round(sc["ms_clear"], 3),
                                   "c_pack": round(sc["ms_pack"], 3), "d_clear": round(sd["ms_clear"], 3), "d_pack": round(sd["ms_pack"], 3),
                                   "c_plan": round(sc["ms_plan"], 3),
                                   "c_gen": round(acc["c_gen"] / K_, 3), "c_qlt": round(acc["c_qlt"] / K_, 3), "c_rec": round(acc["c_rec"] / K_, 3),
                                   "d_gen": round(acc["d_gen"] / K_, 3), "d_qlt": round(acc["d_qlt"] / K_, 3), "d_rec": round(acc["d_rec"] / K_, 3)},
        }
        if dev["parity"]:
            line["parity_sampled"] = len(dev["parity"]["chunks"])
            line["parity"] = dev["parity"]
        if e2e:
            ev = 2 * tot_bytes * K_ / e2e_wall_max / 1e9
            line["e2e"] = {"value": round(ev, 4), "unit": UNIT, "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                           "timing": "wall clock around K steps of sfq_compress+sfq_decompress on pinned host buffers, barrier+sync both sides",
                           "device_event_ms_per_step": round(e2e["ms"] / K_, 3),
                           "copy_ms_per_step": round(e2e["copy_ms"] / K_, 3), "last_step_ms": e2e["parts"], "per_step_ms[c_total,c_code,d_total,d_gen,d_qlt,d_rec]": e2e["steps"],
                           "note": "copy_ms_per_step = copy time NOT hidden behind coding (sfq_stats.ms_h2d + ms_d2h: the exposed head and tail of the pipelined copies)"}
        elif e2e_err:
            line["e2e"] = {"value": None, "unit": UNIT, "error": e2e_err}

    # ---------------- extras (N = 1 only; all outside the timed headline region)
    if extras:
        try:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)          # the CPU legs below use every core the process was given
            run_extras(args, B, K, block, d_text, line, sc, sd, acc, cores)
        except Exception as ex:                          # the headline numbers stand on their own
            line["extras_error"] = "%s: %s" % (type(ex).__name__, str(ex)[:300])
    del d_text
    if rank == 0:
        print(json.dumps(line))
    codec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_extras(args, B: Bench, K, block: bytes, d_text, line: dict, sc: dict, sd: dict, acc: dict, cores: int):
    torch, codec = B.torch, B.codec
    from slimfastq_b200.api import record_start_at_or_after

    # the context's grow-only workspace was sized for the 10 GB calls next to the tensors alive then (~90 % of what was
    # free): hand it back before the legs below bring tensors of their own
    codec.trim()
    torch.cuda.empty_cache()

    # ---- chain micro-benchmark: two chunks resident (one quality-decoder warp on the whole GPU) = the shortest link this build reaches
    two = block[: record_start_at_or_after(block, 2 * args.chunk)]
    d_two = B.upload(two, 1)
    p = B.device_steps(d_two, args.level, args.chunk, 3, 1)
    del d_two
    probe = chain_numbers(p["sc"], p["sd"], p["acc"], 3)
    line["chain"]["one_warp_resident"] = {"chunks": p["sc"]["nchunks"], "compress": probe["compress"], "decompress": probe["decompress"]}
    full, one = line["chain"]["decompress"]["qlt_ns_per_quality"], probe["decompress"]["qlt_ns_per_quality"]
    line["chain"]["frac"] = round(one / full, 4) if full else None
    line["chain"]["frac_note"] = ("dominant chain (quality decoder): link time with one resident warp / link time at the benched residency; "
                                  "1.0 = the chains do not slow each other down, the rest is issue-slot and memory-system contention")

    # ---- ratio yardstick (reference, whole file, one core) and the cost of chunking per chunk size
    whole = None
    if not args.no_cpu:
        whole = reference_whole_file(block, args.level)
        line["whole_file_reference"] = {"sample": "the %d MB unique block as ONE file through the unmodified reference" % (len(block) // 10**6),
                                        "stream_bytes": whole["stream_bytes"], "stream_ratio": round(whole["ratio"], 4),
                                        "one_core_compress_MBps": round(whole["one_core_compress_MBps"], 2)}
    sizes = [args.chunk] + [int(float(x) * (1 << 20)) for x in args.pareto.split(",") if x.strip()]
    pareto = []
    for cb in sizes:
        if cb == args.chunk:
            r = {"value": line["value"], "compress_GBps": line["compress_GBps"], "decompress_GBps": line["decompress_GBps"]}
            if "e2e" in line and line["e2e"].get("value"):
                r["e2e"] = line["e2e"]["value"]
        else:
            m = B.device_steps(d_text, args.level, cb, 2, 1)
            a = m["acc"]
            r = {"value": round(2 * m["n"] * 2 / ((a["t_c"] + a["t_d"]) / 1e3) / 1e9, 4),
                 "compress_GBps": round(m["n"] * 2 / (a["t_c"] / 1e3) / 1e9, 4), "decompress_GBps": round(m["n"] * 2 / (a["t_d"] / 1e3) / 1e9, 4),
                 "waves": [m["sc"]["waves"], m["sd"]["waves"]]}
        ours = K.parse(codec.compress(block, args.level, cb)).stream_bytes
        r.update({"chunk_MiB": round(cb / (1 << 20), 3), "stream_ratio": round(len(block) / ours, 4)})
        if whole:
            r["chunking_loss_pct"] = round(100 * (ours / whole["stream_bytes"] - 1), 3)
        pareto.append(r)
    line["chunk_pareto"] = {"rows": pareto, "note": "device-resident GB/s on the same %.1f GB (2 steps after 1 warm-up, except the headline row); loss = our stream bytes on the unique block vs the reference coding that block as one file; target <= 2 %%" % (d_text.numel() / 1e9)}
    ok = [r for r in pareto if r.get("chunking_loss_pct") is not None and r["chunking_loss_pct"] <= 2.0]
    if ok:
        best = max(ok, key=lambda r: r["value"])
        line["chunk_pareto"]["best_within_2pct"] = {"chunk_MiB": best["chunk_MiB"], "value": best["value"], "chunking_loss_pct": best["chunking_loss_pct"]}

    # ---- BASELINE configs[2] / configs[3]: short legs at the same size
    legs = {}
    for name, kind, bins8, level in (("level4", args.workload, args.bins8, 4), ("level4_8bin", "illumina", True, 4), ("ont", "ont", False, 3)):
        if name == "level4":
            blk, dt = block, d_text
        else:
            del dt
            codec.trim()
            torch.cuda.empty_cache()
            blk = workload_block(kind, 0, bins8)
            dt = B.upload(blk, max(1, round(d_text.numel() / len(blk))))
        m = B.device_steps(dt, level, args.chunk, 2, 1)
        a = m["acc"]
        leg = {"workload": workload_label(kind, bins8, args.gb), "level": level, "chunk_bytes": args.chunk, "steps": 2, "warmup": 1,
               "value": round(2 * m["n"] * 2 / ((a["t_c"] + a["t_d"]) / 1e3) / 1e9, 4),
               "compress_GBps": round(m["n"] * 2 / (a["t_c"] / 1e3) / 1e9, 4), "decompress_GBps": round(m["n"] * 2 / (a["t_d"] / 1e3) / 1e9, 4),
               "ratio": round(m["n"] / m["csz"], 4), "stream_ratio": round(m["n"] / m["sc"]["stream_bytes"], 4), "waves": [m["sc"]["waves"], m["sd"]["waves"]]}
        if not args.no_e2e:
            e, err = B.e2e_steps(dt, m["csz"], level, args.chunk, 2, 2)
            leg["e2e"] = round(2 * m["n"] * 2 / e["wall_s"] / 1e9, 4) if e else None
            if e:
                leg["e2e_per_step_ms[c_total,c_code,d_total,d_gen,d_qlt,d_rec]"] = e["steps"]
            if err:
                leg["e2e_error"] = err
        legs[name] = leg
    del dt
    torch.cuda.empty_cache()
    line["configs"] = legs

    # ---- the reference on this box's host cores
    if not args.no_cpu:
        try:
            r = reference_timed(block, args.level, args.chunk, cores, steps=3, warmup=1, budget_s=30.0, sample_mb=args.cpu_sample_mb)
            cb = {"value": round(r["value"], 5), "unit": UNIT, "cores": cores, "kind": "reference",
                  "sample": "%d MiB of the workload as %d chunk files on tmpfs, level %d, compress then decompress via %s; 1 discarded + 3 timed passes"
                            % (r["sample_bytes"] >> 20, r["files"], args.level, r["how"]),
                  "spread": r["value_spread"], "compress_GBps": r["compress_GBps"], "decompress_GBps": r["decompress_GBps"]}
            cb["best_case"] = reference_best_case(block, args.level, cores)
            line["cpu_baseline"] = cb
        except Exception as ex:
            line["cpu_baseline"] = {"error": str(ex)[:200]}


# ------------------------------------------------------------------------------ one file over N GPUs
def run_single_file(args, B: Bench, rank: int, world: int) -> int:
    """north_star: "chunks are partitioned by index across the GPUs ... only a host-side exchange of compressed-size offsets to
    lay out the container".  Every rank takes the ranges of ONE file that split_on_grid assigns to it, codes them through
    sfq_compress with their grid phase, the blob sizes go through one all_gather, every rank writes its blobs at its offset of
    a shared file and rank 0 adds header and index.  The merged file must equal the 1-GPU container byte for byte."""
    import hashlib
    import struct

    torch, dist, codec = B.torch, B.dist, B.codec
    from slimfastq_b200 import api, container as K

    nbytes = int(args.gb * 1e9)
    block = workload_block(args.workload, 0, args.bins8)          # the same file on every rank (deterministic generator)
    tiles = max(1, round(nbytes / len(block)))
    n = len(block) * tiles
    # split_on_grid on the tiled file without materialising it on the host: grid lines at k*B, record starts from the block
    nslots = (n + args.chunk - 1) // args.chunk
    lines = sorted({min(nslots, round(nslots * p / world)) for p in range(world)} | {0})

    def rec_start(pos):                                           # first record start >= pos in the tiled file
        if pos <= 0:
            return 0
        t, o = divmod(pos, len(block))
        s = api.record_start_at_or_after(block, o)
        return min(n, t * len(block) + s)

    cuts = [(rec_start(k * args.chunk), k) for k in lines]
    my = None
    if rank < len(cuts):
        g, k = cuts[rank]
        end = cuts[rank + 1][0] if rank + 1 < len(cuts) else n
        my = (g, end, g - k * args.chunk)

    def part_bytes(a, b):
        out = bytearray()
        while a < b:
            t, o = divmod(a, len(block))
            take = min(b - a, len(block) - o)
            out += block[o:o + take]
            a += take
        return out

    h_part = None
    if my and my[1] > my[0]:
        raw = part_bytes(my[0], my[1])
        h_part = torch.empty(len(raw), dtype=torch.uint8, pin_memory=True)
        h_part.copy_(torch.frombuffer(raw, dtype=torch.uint8))
        del raw
    shm = "/dev/shm/sfq_single_%d.sfq" % int(os.environ.get("MASTER_PORT", "0"))
    times = []
    merged_md5 = None
    for it in range(args.warmup + args.steps):
        B.barrier()
        t0 = time.perf_counter()
        body = b""
        sizes = []
        if h_part is not None:
            addr, cn = codec.compress_view(h_part, args.level, args.chunk, phase=my[2])
            fh = K.FILE_HDR.unpack_from(ctypes.string_at(addr, K.FILE_HDR.size), 0)
            idx = struct.unpack_from("<%dQ" % fh[5], ctypes.string_at(addr + fh[7], 8 * fh[5]), 0)
            ends = list(idx[1:]) + [fh[7]]
            sizes = [e - s for s, e in zip(idx, ends)]
            body = (addr + K.FILE_HDR.size, fh[7] - K.FILE_HDR.size)
            meta = (fh[4], fh[8])
        else:
            meta = (0, 0)
        # the one exchange of the data path: blob sizes (and the byte counts) of every rank
        cnt = torch.tensor([len(sizes), meta[0], meta[1]], dtype=torch.int64, device="cuda")
        allcnt = [torch.zeros_like(cnt) for _ in range(world)]
        if world > 1:
            dist.all_gather(allcnt, cnt)
        else:
            allcnt = [cnt]
        counts = [int(c[0]) for c in allcnt]
        mx = max(counts + [1])
        sz = torch.zeros(mx, dtype=torch.int64, device="cuda")
        if sizes:
            sz[:len(sizes)] = torch.tensor(sizes, dtype=torch.int64)
        allsz = [torch.zeros_like(sz) for _ in range(world)]
        if world > 1:
            dist.all_gather(allsz, sz)
        else:
            allsz = [sz]
        flat = []
        for r in range(world):
            flat += allsz[r][:counts[r]].cpu().tolist()
        body_off = [K.FILE_HDR.size]
        for r in range(world):
            body_off.append(body_off[-1] + sum(allsz[r][:counts[r]].cpu().tolist()))
        if rank == 0:
            with open(shm, "wb") as f:
                f.truncate(body_off[-1] + 8 * len(flat))
        B.barrier()
        if body:
            with open(shm, "r+b") as f:
                f.seek(body_off[rank])
                f.write(ctypes.string_at(body[0], body[1]))
        B.barrier()
        if rank == 0:
            index, o = [], K.FILE_HDR.size
            for s in flat:
                index.append(o); o += s
            with open(shm, "r+b") as f:
                f.write(K.FILE_HDR.pack(K.STAMP, K.KIND, 6, args.level, sum(int(c[1]) for c in allcnt), len(index), args.chunk, o, sum(int(c[2]) for c in allcnt)))
                f.seek(o)
                f.write(struct.pack("<%dQ" % len(index), *index))
        B.barrier()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ok = None
    if rank == 0:
        merged = open(shm, "rb").read()
        merged_md5 = hashlib.md5(merged).hexdigest()
        # the 1-GPU container of the same file, and the round trip of the merged one
        d_text = B.upload(block, tiles)
        d_sfq = torch.empty(codec.compress_bound(n, args.chunk), dtype=torch.uint8, device="cuda")
        csz = codec.compress_device(d_text, d_sfq, args.level, args.chunk)
        one_md5 = hashlib.md5(bytes(d_sfq[:csz].cpu().numpy())).hexdigest()
        d_m = torch.frombuffer(bytearray(merged), dtype=torch.uint8).cuda()
        d_back = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        osz = codec.decompress_device(d_m, len(merged), d_back)
        ok = {"merged_equals_single_gpu_container": merged_md5 == one_md5, "merged_md5": merged_md5, "single_gpu_md5": one_md5,
              "merged_round_trip": bool(osz == n and torch.equal(d_back[:n], d_text))}
        os.unlink(shm)
    t = torch.tensor([sum(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall = float(t.item())
        line = {"metric": "fastq_single_file_compress_throughput", "value": round(n * args.steps / wall / 1e9, 4), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic (one %.1f GB file: the %d MB block tiled x%d)" % (n / 1e9, len(block) // 10**6, tiles),
                "config": {"workload": workload_label(args.workload, args.bins8, args.gb) + "_single_file", "level": args.level, "chunk_bytes": args.chunk, "file_bytes": n,
                           "sharding": "split_on_grid over %d ranks; sizes all_gather'ed; every rank writes its blobs at its offset of one tmpfs file" % world,
                           "timing": "wall clock per step, max over ranks: host part -> sfq_compress (pinned) -> size exchange -> write into the shared file"},
                "check": ok}
        print(json.dumps(line))
        return 0 if ok["merged_equals_single_gpu_container"] and ok["merged_round_trip"] else 1
    return 0


def run_reference_arm(args, cores):
    """--impl reference: the reference's own CPU implementation on all host cores, same metric/config; every step codes
    a bounded sample of the workload sized so that the whole run fits --ref-budget-s."""
    block = workload_block(args.workload, 0, args.bins8)
    r = reference_timed(block, args.level, args.chunk, cores, args.steps, args.warmup, args.ref_budget_s, args.cpu_sample_mb)
    value = r["value"]
    sample = "%d MiB as %d chunk files on tmpfs, level %d, via %s; sample sized from a calibration pass for a %.0f s run" % (
        r["sample_bytes"] >> 20, r["files"], args.level, r["how"], args.ref_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic (same generator and seed as the GPU arm)",
            "config": {"workload": workload_label(args.workload, args.bins8, args.gb),
                       "level": args.level, "chunk_bytes": args.chunk,
                       "step": "compress + decompress", "note": "each step codes a bounded sample of the workload"},
            "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             "spread": r["value_spread"]},
            "compress_GBps": r["compress_GBps"], "decompress_GBps": r["decompress_GBps"], "wall_s": r["wall_s"],
            "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:
        if r["wall_s"] < args.ref_budget_s:
            line["cpu_baseline"]["best_case"] = reference_best_case(block, args.level, cores)
    except Exception as ex:
        line["cpu_baseline"]["best_case"] = {"error": str(ex)[:200]}
    print(json.dumps(line))


if __name__ == "__main__":
    sys.exit(main())
