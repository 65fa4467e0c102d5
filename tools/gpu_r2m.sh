mkdir -p gpurun_out
B="timeout 400 python bench.py --warmup 1 --no-cpu --no-extras --no-e2e"
run() { # name, env, extra args
  env $2 $B $3 --gb 10 > gpurun_out/r2m_$1.json 2> gpurun_out/r2m_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m_$1.json").read().strip().splitlines()[-1])
    print("$1", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}); p = d["phases_ms_per_step"]
    print("  ", {k: p[k] for k in ("c_code", "d_code", "d_gen", "d_qlt", "d_rec")})
    print("   d_gen per step:", [x[1] for x in d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"]], "d_qlt:", [x[2] for x in d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"]][:4])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2m_$1.err").read()[-1500:])
PY
}
run genfirst "A=1" "--steps 10"
run genfirst_nospread "SFQ_SPREAD=0" "--steps 10"
run genfirst_nospread_w1 "SFQ_SPREAD=0 SFQ_DEC_WARPS=1" "--steps 6"
run round1order "SFQ_DEC_ORDER=0" "--steps 10"
