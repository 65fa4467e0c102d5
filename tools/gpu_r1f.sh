mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
B="timeout 400 python bench.py --gb ${GB:-4} --steps 2 --warmup 1 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; echo "== $name"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/v_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value","compress_GBps","decompress_GBps","chain")}); print(d["phases_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/v_$name.err").read()[-1500:])
PY
}
run base A=1
KERN=k_decode SKIP=0 COUNT=3 TAG=r1g GB=0.25 CHUNK=131072 bash tools/gpu_ncu_sections.sh 2>&1 | tail -2
