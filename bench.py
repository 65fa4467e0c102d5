#!/usr/bin/env python
"""Benchmark of the slimfastq hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gb G] [--level L]

One step = one compress pass + one decompress pass over the workload (BASELINE.json configs[1]:
synthetic Illumina 2x150, 40-level Phred, 10 GB, level 3, 1 MiB chunks).  `value` is FASTQ GB
moved through the codec per second with inputs resident in HBM (2 x workload / (t_c + t_d));
`e2e` is the same through sfq_compress/sfq_decompress with pinned HOST buffers, copies inside the
timed region.  N > 1: every rank codes its own shard of chunks (weak scaling, no collective on
the data path; only sizes are exchanged), launched by torchrun.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
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
    ap.add_argument("--cpu-sample-mb", type=int, default=0, help="CPU baseline sample (0 = 32 MiB x cores, <= 1 GiB)")
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
def _write_chunks(data: bytes, chunk_lens: list[int], d: str) -> list[str]:
    paths, pos = [], 0
    for i, ln in enumerate(chunk_lens):
        p = os.path.join(d, f"c{i:06d}.fq")
        with open(p, "wb") as f:
            f.write(data[pos:pos + ln])
        paths.append(p)
        pos += ln
    return paths


def reference_cpu_pass(sample: bytes, chunk_lens: list[int], level: int, cores: int) -> dict:
    """Times the UNMODIFIED reference (oracle/_ref/slimfastq) over the same chunks the GPU path codes,
    one process per chunk file on `cores` workers via the reference's own tools/slimfastq.multi
    (which cannot pass a level, hence the 2-line wrapper), falling back to a thread pool of
    subprocesses if perl ithreads are missing.  tmpfs files; compress then decompress."""
    from oracle import oracle as O

    if not O.have_ref():
        raise RuntimeError("oracle/_ref/slimfastq is not built")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="sfqref_", dir=base)
    try:
        src = os.path.join(d, "fq"); comp = os.path.join(d, "sfq"); back = os.path.join(d, "back")
        for x in (src, comp, back):
            os.mkdir(x)
        paths = _write_chunks(sample, chunk_lens, src)
        wrapper = os.path.join(d, "sfq_l.sh")
        with open(wrapper, "w") as f:
            f.write(f"#!/bin/sh\nexec {O.REF_BIN} -l {level} \"$@\"\n")
        os.chmod(wrapper, 0o755)
        multi = os.path.join(os.path.dirname(O.REF_BIN), "tools", "slimfastq.multi")
        how = "tools/slimfastq.multi"
        t0 = time.perf_counter()
        r = subprocess.run(["perl", multi, "-c", str(cores), "-e", wrapper, "-t", comp, src], capture_output=True, text=True)
        tc = time.perf_counter() - t0
        ok = r.returncode == 0 and len(os.listdir(comp)) == len(paths)
        if ok:
            t0 = time.perf_counter()
            r = subprocess.run(["perl", multi, "-d", "-c", str(cores), "-e", wrapper, "-t", back, comp], capture_output=True, text=True)
            td = time.perf_counter() - t0
            ok = r.returncode == 0 and len(os.listdir(back)) == len(paths)
        if not ok:                                   # same work without perl
            from concurrent.futures import ThreadPoolExecutor

            how = "thread pool of reference processes"
            for x in (comp, back):
                shutil.rmtree(x); os.mkdir(x)

            def enc(p):
                subprocess.run([O.REF_BIN, "-l", str(level), "-q", "-O", "-u", p, "-f", os.path.join(comp, os.path.basename(p) + ".sfq")], check=True)

            def dec(p):
                subprocess.run([O.REF_BIN, "-d", "-O", "-f", os.path.join(comp, os.path.basename(p) + ".sfq"), "-u", os.path.join(back, os.path.basename(p))], check=True)

            with ThreadPoolExecutor(cores) as ex:
                t0 = time.perf_counter(); list(ex.map(enc, paths)); tc = time.perf_counter() - t0
                t0 = time.perf_counter(); list(ex.map(dec, paths)); td = time.perf_counter() - t0
        comp_bytes = sum(os.path.getsize(os.path.join(comp, f)) for f in os.listdir(comp))
        n = len(sample)
        return {"t_compress": tc, "t_decompress": td, "compress_GBps": n / tc / 1e9, "decompress_GBps": n / td / 1e9,
                "value": 2 * n / (tc + td) / 1e9, "how": how, "sfq_file_bytes": comp_bytes}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_whole_file_ratio(sample: bytes, level: int) -> dict:
    """Whole-file run of the reference on the sample: its stream bytes are the ratio yardstick."""
    from oracle import oracle as O

    t0 = time.perf_counter()
    enc = O.ref_encode(sample, level, tmpdir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    dt = time.perf_counter() - t0
    sb = sum(len(v) for v in enc.streams.values())
    return {"stream_bytes": sb, "ratio": len(sample) / sb, "one_core_compress_MBps": len(sample) / dt / 1e6}


# ------------------------------------------------------------------------------ workload
ONT_READS = 6400                # ~129 MB of ONT-style reads (log-normal 1-50 kb), tiled like the Illumina block


def workload_block(kind: str, rank: int, bins8: bool = False) -> bytes:
    from slimfastq_b200 import synth

    if kind == "ont":
        return synth.ont(ONT_READS, seed=synth.SEED0 + 4 + 1000 * rank)
    return synth.illumina(UNIQUE_READS, seed=synth.SEED0 + 1 + 1000 * rank, bins8=bins8)


def make_workload(nbytes: int, rank: int, bins8: bool = False, kind: str = "illumina"):
    """(unique block bytes, tiles): one synthetic block, tiled to `nbytes`.  Chunks are coded independently,
    so tiling changes neither the ratio nor the per-chunk work."""
    block = workload_block(kind, rank, bins8)
    tiles = max(1, round(nbytes / len(block)))
    return block, tiles


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nbytes = int(args.gb * 1e9)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        run_reference_arm(args, nbytes, cores)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    block, tiles = make_workload(nbytes, rank, args.bins8, args.workload)
    n = len(block) * tiles
    d_block = torch.frombuffer(bytearray(block), dtype=torch.uint8).cuda()
    d_text = d_block.repeat(tiles)
    del d_block
    codec = S.Codec(local)
    d_sfq = torch.empty(codec.compress_bound(n, args.chunk), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    # ---------------- device-resident steps (value)
    def step_device():
        csz = codec.compress_device(d_text, d_sfq, args.level, args.chunk)
        sc = codec.stats()
        osz = codec.decompress_device(d_sfq, csz, d_back)
        sd = codec.stats()
        assert osz == n
        return csz, sc, sd

    for _ in range(args.warmup):
        csz, sc, sd = step_device()
    assert torch.equal(d_back[:n], d_text), "device round trip differs"
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_c = t_d = 0.0
    launches = 0
    code_ms_c = code_ms_d = scan_ms = qd_ms = 0.0
    waves_c = waves_d = 0
    per_step = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        csz, sc, sd = step_device()
        t_c += sc["ms_total"]; t_d += sd["ms_total"]
        launches += sc["kernel_launches"] + sd["kernel_launches"]
        code_ms_c += sc["ms_code"]; code_ms_d += sd["ms_code"]; scan_ms += sc["ms_scan"]; qd_ms += sd["ms_qlt"]
        waves_c += sc["waves"]; waves_d += sd["waves"]
        per_step.append([round(sc["ms_code"], 1), round(sd["ms_gen"], 1), round(sd["ms_qlt"], 1), round(sd["ms_rec"], 1)])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = t_c + t_d            # CUDA-event time on the library's stream, summed over the K steps

    # ---------------- end-to-end steps (host buffers, copies inside the timed region)
    e2e = None
    e2e_err = None
    if not args.no_e2e:
        try:
            # pinned host memory of the leg: input + the library's container and output buffers, on every rank of the node
            need = world * (2 * n + csz + (1 << 29))
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
            del d_text, d_sfq, d_back                  # the host-buffer entry points stage through the library's own buffers
            torch.cuda.empty_cache()

            def step_e2e():
                a, cn = codec.compress_view(h_text, args.level, args.chunk)     # container in the context's pinned buffer
                s1 = codec.stats()
                b, on = codec.decompress_addr(a, cn)                            # (decompress returns into a buffer of its own)
                s2 = codec.stats()
                return cn, on, s1, s2, b

            for _ in range(min(args.warmup, 3)):
                cn, on, s1, s2, b = step_e2e()
            assert on == n
            m = min(n, 64 << 20)
            assert ctypes.string_at(b, m) == bytes(h_text[:m].numpy()), "e2e round trip differs"
        except (RuntimeError, S.SfqError, MemoryError) as ex:      # e.g. not enough pinned host memory for N ranks
            e2e_err = str(ex)[:200]
        # every rank must take the same path through the barriers below
        ok = 0.0 if e2e_err else 1.0
        if world > 1:
            tok = torch.tensor([ok], dtype=torch.float64, device="cuda")
            dist.all_reduce(tok, op=dist.ReduceOp.MIN)
            ok = float(tok.item())
        if ok:
            barrier()
            e_ms = e_copy = 0.0
            e_steps = []
            t0 = time.perf_counter()
            for _ in range(args.steps):
                cn, on, s1, s2, b = step_e2e()
                e_ms += s1["ms_total"] + s2["ms_total"]
                e_copy += s1["ms_h2d"] + s1["ms_d2h"] + s2["ms_h2d"] + s2["ms_d2h"]
                e_steps.append([round(s1["ms_total"], 1), round(s1["ms_code"], 1), round(s2["ms_total"], 1), round(s2["ms_gen"], 1), round(s2["ms_qlt"], 1), round(s2["ms_rec"], 1)])
                e_parts = {"c_total": s1["ms_total"], "c_h2d": s1["ms_h2d"], "c_code": s1["ms_code"], "c_d2h": s1["ms_d2h"], "c_waves": s1["waves"],
                           "d_total": s2["ms_total"], "d_h2d": s2["ms_h2d"], "d_code": s2["ms_code"], "d_d2h": s2["ms_d2h"], "d_waves": s2["waves"],
                           "d_resident": s2["resident_chunks"]}
                launches += s1["kernel_launches"] + s2["kernel_launches"]
            barrier()
            e_wall = time.perf_counter() - t0
            e2e = {"ms": e_ms, "wall_s": e_wall, "h2d": n + cn, "d2h": cn + n, "copy_ms": e_copy, "steps": e_steps, "parts": {k: round(float(v), 2) for k, v in e_parts.items()}}
            del h_text
        elif not e2e_err:
            e2e_err = "another rank could not set up its host buffers"

    # ---------------- reduce over ranks: max time, sum bytes
    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    tot_bytes = allsum(float(n))
    tot_csz = allsum(float(csz))
    dev_ms_max = allmax(dev_ms)
    t_c_max, t_d_max = allmax(t_c), allmax(t_d)
    wall_max = allmax(wall)
    e2e_wall_max = allmax(e2e["wall_s"]) if e2e else None
    launches_sum = allsum(float(launches))

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
        plane_bytes = sc["nbases"] + sc["nquals"] + (n - sc["nbases"] - sc["nquals"] - 6 * sc["nrecords"])
        code_ms = code_ms_c + code_ms_d
        qd_bytes = sd["nquals"] + sd["qlt_stream_bytes"]
        achieved = qd_bytes * K_ / (qd_ms / 1e3) / 1e9
        symbols = sc["nbases"] + sc["nquals"]
        hdr_bytes = n - sc["nbases"] - sc["nquals"] - 6 * sc["nrecords"]

        def gbps(nbytes, ms):
            return round(nbytes / (ms / 1e3) / 1e9, 3) if ms > 0 else None

        coder_kernels = {   # every coder kernel group of one step against its own algorithmic bytes (last step's timings)
            "compress_gen (k_gen_model + k_rc_encode<0>)": gbps(sc["nbases"] + sc["gen_stream_bytes"], sc["ms_gen"]),
            "compress_qlt (k_qlt_keys/scan/scatter/model + k_rc_encode<1>)": gbps(sc["nquals"] + sc["qlt_stream_bytes"], sc["ms_qlt"]),
            "compress_rec (k_encode<2>)": gbps(hdr_bytes, sc["ms_rec"]),
            "decompress_gen (k_decode<0>)": gbps(sd["nbases"] + sd["gen_stream_bytes"], sd["ms_gen"]),
            "decompress_qlt (k_qlt_decode<LPC>)": gbps(sd["nquals"] + sd["qlt_stream_bytes"], sd["ms_qlt"]),
            "decompress_rec (k_decode<2>)": gbps(hdr_bytes, sd["ms_rec"]),
        }
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": args.warmup,
            "ms_per_step": round(dev_ms_max / K_, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic (%s block of %d MB, tiled x%d; chunks are coded independently)" % ("ONT-style long-read" if args.workload == "ont" else "Markov-quality Illumina", len(block) // 10**6, tiles),
            "config": {"workload": ("ont_1-50kb_%.1fGB_per_gpu" % (n / 1e9)) if args.workload == "ont" else "illumina_2x150_%s_%.1fGB_per_gpu" % ("8bin" if args.bins8 else "phred40", n / 1e9),
                       "level": args.level, "chunk_bytes": args.chunk, "bytes_per_gpu": n, "l2": "inputs (%.1f GB) exceed L2 (126 MB); no flush needed" % (n / 1e9),
                       "step": "compress + decompress", "sharding": "chunks by rank, no data-path collective"},
            "compress_GBps": round(tot_bytes * K_ / (t_c_max / 1e3) / 1e9, 4),
            "decompress_GBps": round(tot_bytes * K_ / (t_d_max / 1e3) / 1e9, 4),
            "ratio": round(tot_bytes / tot_csz, 4), "stream_ratio": round(n / sc["stream_bytes"], 4),
            "wall_s_per_step": round(wall_max / K_, 4),
            "clocks": clocks,
            "gpu_launches": int(launches_sum),
            "roofline": {"bound": "hbm", "kernel": "k_qlt_decode<%d> (quality decoder: %d lanes per chunk, %d chunks per warp, warp-converged)" % ((4, 4, 8) if sd["resident_chunks"] >= 4096 else (8, 8, 4)),
                         "achieved": round(achieved, 3), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 6),
                         "traffic": None,
                         "traffic_profiled": {"source": "profiles/r1e_ncu_full_summary.txt (ncu --set full, k_qlt_decode<8>, 247 chunks = 108.3 M qualities per launch)",
                                              "dram_bytes_per_launch": 8436278000, "algorithmic_bytes_per_launch": 135400000,
                                              "note": "a 10 GB launch cannot be replayed by ncu; at full residency every model visit misses L2 (256 B read + up to 256 B written back per quality)"},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes_per_launch": int(qd_bytes / max(1, sd["waves"])),
                         "launch_ms_avg": round(qd_ms / K_ / max(1, sd["waves"]), 3),
                         "note": "a serial adaptive-coder chain per chunk: bound by issue slots and dependent latency, not by HBM (see chain and DESIGN.md section 3)"},
            "coder_kernels_GBps": coder_kernels,
            "per_step_ms[c_code,d_gen,d_qlt,d_rec]": per_step,
            "chain": {"symbols_per_chunk_stream": round(symbols / 2 / max(1, sc["nchunks"])),
                      "compress": {"resident_chunks": sc["resident_chunks"], "waves": sc["waves"],
                                   "ns_per_symbol_per_wave": round(code_ms_c / K_ * 1e6 / (symbols / 2 / sc["nchunks"]) / max(1, sc["waves"]), 2)},
                      "decompress": {"resident_chunks": sd["resident_chunks"], "waves": sd["waves"],
                                     "ns_per_symbol_per_wave": round(code_ms_d / K_ * 1e6 / (symbols / 2 / sc["nchunks"]) / max(1, sd["waves"]), 2)}},
            "roofline_scan": {"bound": "hbm", "kernel": "k_count_newlines+k_scan_tiles+k_fill_lines", "achieved": round((n + 8 * 4 * sc["nrecords"]) * K_ / (scan_ms / 1e3) / 1e9, 2),
                              "peak": hbm_peak, "unit": "GB/s", "frac": round((n + 32 * sc["nrecords"]) * K_ / (scan_ms / 1e3) / 1e9 / hbm_peak, 4)},
            "phases_ms_per_step": {"c_scan": round(scan_ms / K_, 3), "c_code": round(code_ms_c / K_, 3), "d_code": round(code_ms_d / K_, 3),
                                   "c_total": round(t_c / K_, 3), "d_total": round(t_d / K_, 3), "c_clear": round(sc["ms_clear"], 3),
                                   "c_pack": round(sc["ms_pack"], 3), "d_clear": round(sd["ms_clear"], 3), "d_pack": round(sd["ms_pack"], 3),
                                   "c_plan": round(sc["ms_plan"], 3),
                                   "c_gen": round(sc["ms_gen"], 3), "c_qlt": round(sc["ms_qlt"], 3), "c_rec": round(sc["ms_rec"], 3),
                                   "d_gen": round(sd["ms_gen"], 3), "d_qlt": round(sd["ms_qlt"], 3), "d_rec": round(sd["ms_rec"], 3)},
        }
        if e2e:
            ev = 2 * tot_bytes * K_ / e2e_wall_max / 1e9
            line["e2e"] = {"value": round(ev, 4), "unit": UNIT, "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                           "timing": "wall clock around K steps of sfq_compress+sfq_decompress on pinned host buffers, barrier+sync both sides",
                           "device_event_ms_per_step": round(e2e["ms"] / K_, 3),
                           "copy_ms_per_step": round(e2e["copy_ms"] / K_, 3), "last_step_ms": e2e["parts"], "per_step_ms[c_total,c_code,d_total,d_gen,d_qlt,d_rec]": e2e["steps"],
                           "note": "copies and coding run back to back (no overlap yet): e2e = value's kernels + PCIe time"}
        elif e2e_err:
            line["e2e"] = {"value": None, "unit": UNIT, "error": e2e_err}
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline(args, block, codec, cores, K)
            except Exception as ex:                      # the GPU numbers stand on their own
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        print(json.dumps(line))
    codec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_sample(args, block: bytes, cores: int) -> bytes:
    mb = args.cpu_sample_mb or min(1024, 32 * cores)
    want = mb << 20
    reps = (want + len(block) - 1) // len(block)
    from slimfastq_b200.api import record_start_at_or_after

    s = block * reps
    return s[: record_start_at_or_after(s, want)]


def cpu_baseline(args, block, codec, cores, K) -> dict:
    """Reference CPU path on this box's host cores over a bounded sample of the same workload."""
    sample = cpu_sample(args, block, cores)
    ct = K.parse(codec.compress(sample, args.level, args.chunk))
    lens = [c.text_len for c in ct.chunks]
    r = reference_cpu_pass(sample, lens, args.level, cores)
    from slimfastq_b200.api import record_start_at_or_after

    small = sample[: record_start_at_or_after(sample, min(len(sample), 64 << 20))]
    whole = reference_whole_file_ratio(small, args.level)
    ours_small = K.parse(codec.compress(small, args.level, args.chunk))
    return {"value": round(r["value"], 5), "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d MiB of the workload as %d chunk files on tmpfs, level %d, compress then decompress via %s" % (len(sample) >> 20, len(lens), args.level, r["how"]),
            "compress_GBps": round(r["compress_GBps"], 5), "decompress_GBps": round(r["decompress_GBps"], 5),
            "whole_file_reference": {"sample_MiB": len(small) >> 20, "stream_ratio": round(whole["ratio"], 4),
                                     "one_core_compress_MBps": round(whole["one_core_compress_MBps"], 2),
                                     "ours_stream_ratio_same_sample": round(len(small) / ours_small.stream_bytes, 4),
                                     "chunking_loss_pct": round(100 * (ours_small.stream_bytes / whole["stream_bytes"] - 1), 3)}}


def run_reference_arm(args, nbytes, cores):
    """--impl reference: the reference's own CPU implementation on all host cores, same metric/config."""
    block = workload_block(args.workload, 0, args.bins8)
    sample = cpu_sample(args, block, cores)
    from slimfastq_b200.api import chunk_lengths

    lens = chunk_lengths(sample, args.chunk)         # the product's chunking rule (sfq_plan.cuh)
    for _ in range(args.warmup):
        reference_cpu_pass(sample[: sum(lens[:cores])], lens[:cores], args.level, cores)
    tot = 0.0
    last = None
    for _ in range(args.steps):
        last = reference_cpu_pass(sample, lens, args.level, cores)
        tot += last["t_compress"] + last["t_decompress"]
    value = 2 * len(sample) * args.steps / tot / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(tot / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic (same generator and seed as the GPU arm)",
            "config": {"workload": ("ont_1-50kb_%.1fGB_per_gpu" % (nbytes / 1e9)) if args.workload == "ont" else "illumina_2x150_%s_%.1fGB_per_gpu" % ("8bin" if args.bins8 else "phred40", nbytes / 1e9),
                       "level": args.level, "chunk_bytes": args.chunk,
                       "step": "compress + decompress", "note": "each step codes a bounded sample of the workload"},
            "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "%d MiB as %d chunk files on tmpfs, level %d, via %s" % (len(sample) >> 20, len(lens), args.level, last["how"])},
            "compress_GBps": round(last["compress_GBps"], 5), "decompress_GBps": round(last["decompress_GBps"], 5),
            "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    sys.exit(main())
