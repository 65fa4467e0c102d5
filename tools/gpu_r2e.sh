# full GPU test-suite + 10 GB traces under a few knobs
mkdir -p gpurun_out
timeout 1200 python -m pytest ${TESTS:-tests} -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2e_pytest_gpu.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras"
for knob in ${KNOBS:-A=1}; do
  env SFQ_TRACE=1 $knob $B --gb ${GB:-10} > gpurun_out/r2e_$knob.json 2> gpurun_out/r2e_$knob.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2e_$knob.json").read().strip().splitlines()[-1])
    print("$knob", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps", "stream_ratio")}); p = d["phases_ms_per_step"]
    print({k: p[k] for k in ("c_code", "c_gen", "c_qlt", "c_rec", "d_code", "d_gen", "d_qlt", "d_rec")}, d["chain"]["compress"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2e_$knob.err").read()[-1500:])
PY
  grep "sfq trace" gpurun_out/r2e_$knob.err | tail -${TAILN:-12}
done
