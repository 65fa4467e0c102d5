# round 2, pass ai: can a shared-memory carve-out hint keep the base decoder at one CTA per SM?  (hold = the deterministic bad case)
mkdir -p gpurun_out
for k in "SFQ_DEC_HOLD_US=100 SFQ_DEC_FIT=1 SFQ_GEN_CARVEOUT=40" "SFQ_DEC_FIT=1 SFQ_GEN_CARVEOUT=40" "SFQ_DEC_HOLD_US=100 SFQ_DEC_FIT=1 SFQ_GEN_CARVEOUT=40 SFQ_GEN_RESERVE_KB=48"; do
  tag=$(echo $k | tr ' =' '__')
  env $k timeout 300 python bench.py --steps 8 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10 > gpurun_out/r2ai_$tag.json 2> gpurun_out/r2ai_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2ai_$tag.json").read().strip().splitlines()[-1])
    ps = d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"]
    print("$k", {k: d[k] for k in ("value", "decompress_GBps")}, "d_gen", sorted(p[1] for p in ps), "d_qlt", ps[0][2], "d_rec", ps[0][3])
except Exception as ex:
    print("$k ERR", ex); print(open("gpurun_out/r2ai_$tag.err").read()[-600:])
PY
done
