mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
B="timeout 400 python bench.py --gb 10 --steps 1 --warmup 1 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; echo "== $name"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/v_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value","compress_GBps","decompress_GBps")}, d["clocks"]["nvidia_smi_utilization_pct"]); p=d["phases_ms_per_step"]; print({k:p[k] for k in ("c_code","c_gen","c_qlt","c_rec","d_code","d_gen","d_qlt","d_rec")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/v_$name.err").read()[-1500:])
PY
}
run gm0 SFQ_GM_VARIANT=0
run gm1 SFQ_GM_VARIANT=1
run gm2 SFQ_GM_VARIANT=2
run gm3 SFQ_GM_VARIANT=3
run rc4 SFQ_RC_LANES=4
run rc16 SFQ_RC_LANES=16
