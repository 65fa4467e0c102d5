# round 2, last sanity pass on the committed build: smoke(), one short bench step, the quick GPU tests
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_shapes.py tests/test_plane_hooks.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --gb 10 > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2aj_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}, d["e2e"]["value"], d["per_step_ms[c_code,d_gen,d_qlt,d_rec]"])
PY
