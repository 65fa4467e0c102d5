# round 2, pass ah: 60 device-resident steps with the 100 us hold in front of the base decoder: is the packed-CTA outlier gone?
mkdir -p gpurun_out
timeout 600 python bench.py --steps 60 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10 > gpurun_out/r2ah_hold.json 2> gpurun_out/r2ah_hold.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2ah_hold.json").read().strip().splitlines()[-1])
ps = d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"]
g = sorted(p[1] for p in ps)
print({k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}, "steps", len(ps), "d_gen min/median/max", g[0], g[len(g)//2], g[-1], "outliers(>900):", sum(1 for x in g if x > 900))
print("d_qlt range", min(p[2] for p in ps), max(p[2] for p in ps), "d_rec range", min(p[3] for p in ps), max(p[3] for p in ps))
PY
