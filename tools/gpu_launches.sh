# per-kernel durations of one compress+decompress step (ncu serialises the launches)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --gb ${GB:-2} --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
python - <<'PY'
import csv, re, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]); name = name.replace("void ", "")
    if name.startswith("at::"): continue
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
for k, (n, ms) in agg.items(): print(f"{k:28s} x{n:3d} {ms:10.3f} ms")
PY
