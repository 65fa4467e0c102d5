mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
B="timeout 400 python bench.py --gb ${GB:-4} --steps 1 --warmup 1 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; echo "== $name"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/v_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value","compress_GBps","decompress_GBps","chain","ratio")}); print(d["phases_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/v_$name.err").read()[-1500:])
PY
}
run serial4 SFQ_SERIAL_ROLES=1
B="timeout 400 python bench.py --gb 10 --steps 1 --warmup 1 --no-cpu --no-e2e"
run d10_lanes4 A=1
run d10_lanes8 SFQ_LANES=8
run d10_lanes16 SFQ_LANES=16
