# shapes + parts tests, decode A/B (speculative prefetch, lanes per chunk), e2e with and without pipelined parts
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2i_pytest_gpu.log
B="timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
for knob in "A=1" "SFQ_QSPEC=1" "SFQ_QLPC=8" "SFQ_PARTS=1"; do
  env $knob $B --gb 10 > gpurun_out/r2i_$knob.json 2> gpurun_out/r2i_$knob.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2i_$knob.json").read().strip().splitlines()[-1])
    print("$knob", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}, "e2e", d["e2e"]["value"], d["e2e"]["copy_ms_per_step"], d["e2e"]["last_step_ms"]); p = d["phases_ms_per_step"]
    print({k: p[k] for k in ("c_code", "c_gen", "c_qlt", "c_rec", "d_code", "d_gen", "d_qlt", "d_rec")}, d["chain"]["compress"]["waves"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2i_$knob.err").read()[-1500:])
PY
done
