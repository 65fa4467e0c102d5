# full-size bench + launch list + one full ncu capture of the coder kernels (small input: ncu replays ~40x)
timeout 900 python bench.py > gpurun_out/bench_10gb.json 2> gpurun_out/bench_10gb.err; tail -c 4000 gpurun_out/bench_10gb.json; tail -3 gpurun_out/bench_10gb.err
for L in 2 8; do SFQ_LANES=$L timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_10gb_lanes$L.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_10gb_lanes$L.json").read().strip().splitlines()[-1]); print("lanes$L", d["value"], d["compress_GBps"], d["decompress_GBps"], d["phases_ms_per_step"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --gb 1 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
