mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_batch_driver.py tests/test_cli.py tests/test_reference_format.py -m gpu -x -q > gpurun_out/r2l_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2l_pytest_gpu.log
timeout 400 python bench.py --warmup 1 --steps 3 --no-cpu --no-extras --gb 10 > gpurun_out/r2l_e2e.json 2> gpurun_out/r2l_e2e.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2l_e2e.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "compress_GBps", "decompress_GBps")}, d["phases_ms_per_step"])
print(d["e2e"]["value"], d["e2e"]["copy_ms_per_step"], d["e2e"]["last_step_ms"])
PY
