# A/B harness: TAG=<name> KNOBS="A=1 SFQ_X=1 ..." [TESTS="tests/..."] [ARGS="--gb 10 ..."] bash tools/gpu_ab2.sh
mkdir -p gpurun_out
T=${TAG:-ab}
if [ -n "$TESTS" ]; then timeout 1200 python -m pytest $TESTS -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.log; fi
B="timeout 400 python bench.py ${ARGS:---steps 3 --warmup 1 --no-cpu --no-extras --gb 10}"
for knob in ${KNOBS:-A=1}; do
  k=$(echo $knob | tr ',' ' ')
  env $k $B > gpurun_out/${T}_$knob.json 2> gpurun_out/${T}_$knob.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_$knob.json").read().strip().splitlines()[-1])
    p = d["phases_ms_per_step"]
    print("$knob", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}, "waves", d["chain"]["compress"]["waves"], d["chain"]["decompress"]["waves"])
    print("   ", {k: p[k] for k in ("c_total", "c_code", "c_gen", "c_qlt", "c_rec", "d_total", "d_code", "d_gen", "d_qlt", "d_rec")})
    print("    per step", d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"])
    e = d.get("e2e")
    if e: print("    e2e", e["value"], e.get("copy_ms_per_step"), e.get("last_step_ms"))
except Exception as ex:
    print("$knob ERR", ex); print(open("gpurun_out/${T}_$knob.err").read()[-1500:])
PY
  grep "sfq trace\|sfq host" gpurun_out/${T}_$knob.err | tail -${TAILN:-0}
done
