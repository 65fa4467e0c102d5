# round 2, pass t: the decoders one after another (SFQ_SERIAL_ROLES=1): each one's own time, compact header against lane-owned entries
TAG=r2t TAILN=0 KNOBS="SFQ_SERIAL_ROLES=1 SFQ_SERIAL_ROLES=1,SFQ_QCH=0" ARGS="--steps 2 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
ARGS2="--steps 2 --warmup 1 --no-cpu --no-extras --no-e2e --gb 0.13"
for k in A=1 SFQ_QLPC=4 SFQ_QLPC=4,SFQ_QCH=0; do
  env $(echo $k | tr ',' ' ') timeout 300 python bench.py $ARGS2 > gpurun_out/r2t_small_$k.json 2> gpurun_out/r2t_small_$k.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2t_small_$k.json").read().strip().splitlines()[-1]); p = d["phases_ms_per_step"]
print("small $k", {k: p[k] for k in ("d_code", "d_gen", "d_qlt", "d_rec")})
PY
done
