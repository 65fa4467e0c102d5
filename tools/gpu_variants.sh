timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
B="timeout 300 python bench.py --gb 2 --steps 2 --warmup 1 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; echo "== $name"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/v_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value","compress_GBps","decompress_GBps","chain")}); print(d["phases_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/v_$name.err").read()[-800:])
PY
}
run default A=1
run lanes8 SFQ_LANES=8
run lanes1 SFQ_LANES=1
