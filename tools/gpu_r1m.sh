mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
B="timeout 400 python bench.py --gb 10 --steps 1 --warmup 1 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; echo "== $name"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/v_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value","compress_GBps","decompress_GBps","ratio")}); print(d["phases_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/v_$name.err").read()[-1500:])
PY
}
run d10 A=1
run d10_noahead2 SFQ_GEN_AHEAD2=0
B="timeout 400 python bench.py --gb 4 --steps 1 --warmup 1 --no-cpu --no-e2e"
run d4_serial SFQ_SERIAL_ROLES=1
