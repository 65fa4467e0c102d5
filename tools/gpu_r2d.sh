# parity (test_gpu_parity only) + 10 GB trace of the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2d_pytest_gpu.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras"
for knob in ${KNOBS:-A=1}; do
  env SFQ_TRACE=1 $knob $B --gb ${GB:-10} > gpurun_out/r2d_$knob.json 2> gpurun_out/r2d_$knob.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2d_$knob.json").read().strip().splitlines()[-1])
print("$knob", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps", "stream_ratio")}); p = d["phases_ms_per_step"]
print({k: p[k] for k in ("c_code", "c_gen", "c_qlt", "c_rec", "d_code", "d_gen", "d_qlt", "d_rec")}, d["chain"]["compress"])
PY
  grep "sfq trace" gpurun_out/r2d_$knob.err | tail -${TAILN:-18}
done
