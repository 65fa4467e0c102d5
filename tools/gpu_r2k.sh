mkdir -p gpurun_out
SFQ_TRACE=1 timeout 400 python bench.py --warmup 1 --steps 1 --no-cpu --no-extras --gb 10 > gpurun_out/r2k_e2e.json 2> gpurun_out/r2k_e2e.err
grep "sfq host" gpurun_out/r2k_e2e.err | tail -40
python - <<PY
import json
d = json.loads(open("gpurun_out/r2k_e2e.json").read().strip().splitlines()[-1])
print(d["e2e"]["value"], d["e2e"]["last_step_ms"])
PY
