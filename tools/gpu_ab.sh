# A/B runs of bench.py under environment knobs (SFQ_LANES, SFQ_RC_LANES, SFQ_QDEC, SFQ_QGPW, SFQ_SERIAL_ROLES,
# SFQ_MAX_RESIDENT, SFQ_GEN_AHEAD2, SFQ_GM_VARIANT, SFQ_ENC_SERIAL).  Usage: GB=10 bash tools/gpu_ab.sh "A=1" "SFQ_LANES=4" ...
mkdir -p gpurun_out
B="timeout 600 python bench.py --gb ${GB:-4} --steps 1 --warmup 1 --no-cpu --no-e2e"
i=0
for knob in "$@"; do
  i=$((i+1)); env $knob $B > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err; echo "== $knob"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$i.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}); p = d["phases_ms_per_step"]
    print({k: p[k] for k in ("c_code", "c_gen", "c_qlt", "c_rec", "d_code", "d_gen", "d_qlt", "d_rec")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/ab_$i.err").read()[-1500:])
PY
done
