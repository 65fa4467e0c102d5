mkdir -p gpurun_out
B="timeout 400 python bench.py --warmup 1 --no-cpu --no-extras"
run() { # name, env, extra args
  env $2 $B $3 --gb 10 > gpurun_out/r2j_$1.json 2> gpurun_out/r2j_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2j_$1.json").read().strip().splitlines()[-1])
    print("$1", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}); p = d["phases_ms_per_step"]
    print("  ", {k: p[k] for k in ("c_code", "d_code", "d_gen", "d_qlt", "d_rec")})
    print("   per step [c_code,d_gen,d_qlt,d_rec]:", d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"])
    if "e2e" in d: print("   e2e", d["e2e"]["value"], d["e2e"]["copy_ms_per_step"], d["e2e"]["last_step_ms"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2j_$1.err").read()[-1500:])
PY
}
run default "A=1" "--steps 8 --no-e2e"
run reserve112 "SFQ_GEN_RESERVE_KB=112" "--steps 8 --no-e2e"
run reserve140 "SFQ_GEN_RESERVE_KB=140" "--steps 8 --no-e2e"
run e2e_parts "A=1" "--steps 3"
