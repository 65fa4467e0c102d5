# residency / concurrency / groups-per-warp sweep of the decoders (diagnosis)
mkdir -p gpurun_out
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
run base A=1
run r1907 SFQ_MAX_RESIDENT=1907
run r954 SFQ_MAX_RESIDENT=954
run serial SFQ_SERIAL_ROLES=1
run serial_r954 SFQ_SERIAL_ROLES=1 SFQ_MAX_RESIDENT=954
run qgpw1 SFQ_QGPW=1
run qgpw2 SFQ_QGPW=2
EXTRA="--chunk 4194304" run chunk4m A=1
