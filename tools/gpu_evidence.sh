# Round evidence: GPU parity tests, the full default bench line, the reference arm, an ncu launch list of the
# same command at 1 GB, and one full ncu capture of the coder kernels (small input: ncu replays every kernel).
# Usage: TAG=r1c bash tools/gpu_evidence.sh
TAG=${TAG:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_10gb.json 2> gpurun_out/${TAG}_bench_10gb.err; tail -c 6000 gpurun_out/${TAG}_bench_10gb.json; tail -3 gpurun_out/${TAG}_bench_10gb.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cat gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_1gb.csv python bench.py --gb 1 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
python - <<PY
import csv, re, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches_1gb.csv")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    if name.startswith("at::"): continue
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
for k, (n, ms) in agg.items(): print(f"{k:28s} x{n:3d} {ms:10.3f} ms")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gen_model|k_qlt_scatter|k_rc_encode|k_qlt_model|k_qlt_decode|k_decode' -c 8 -o gpurun_out/${TAG}_ncu_full -f python bench.py --gb 0.25 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
