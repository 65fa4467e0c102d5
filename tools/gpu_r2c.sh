# round 2, pass c: parity of the partitioned base-model replay + its effect at 10 GB
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2c_pytest_gpu.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras"
for knob in "A=1" "SFQ_GM_TABLE=1"; do
  env SFQ_TRACE=1 $knob $B --gb 10 > gpurun_out/r2c_$knob.json 2> gpurun_out/r2c_$knob.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2c_$knob.json").read().strip().splitlines()[-1])
print("$knob", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps", "stream_ratio")}); p = d["phases_ms_per_step"]
print({k: p[k] for k in ("c_code", "c_gen", "c_qlt", "c_rec", "d_code", "d_gen", "d_qlt", "d_rec")}, d["chain"]["compress"])
PY
  grep "sfq trace" gpurun_out/r2c_$knob.err | tail -16
done
