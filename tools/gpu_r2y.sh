# round 2, evidence pass y: all GPU tests, the default invocation of both bench arms under the wall clock, single-pass ncu
# metrics of the benched 10 GB launches (-> profiles/r2_traffic_10gb.json via tools/ncu_traffic.py)
mkdir -p gpurun_out
T=r2y
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.log
( time timeout 600 python bench.py --impl reference ) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -4 gpurun_out/${T}_bench_reference.err
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -4 gpurun_out/${T}_bench_default.err
python - <<PY
import json
for f in ("reference", "default"):
    try:
        d = json.loads(open("gpurun_out/${T}_bench_%s.json" % f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    print(f, {k: d.get(k) for k in ("value", "compress_GBps", "decompress_GBps", "ms_per_step")}, (d.get("e2e") or {}).get("value"))
    for k in ("roofline", "chain", "chunk_pareto", "configs", "cpu_baseline", "whole_file_reference", "parity_sampled", "extras_error"):
        if k in d: print("  ", k, json.dumps(d[k])[:1500])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_' --csv --log-file gpurun_out/${T}_ncu_10gb_metrics.csv python bench.py --gb 10 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
python tools/ncu_traffic.py gpurun_out/${T}_ncu_10gb_metrics.csv gpurun_out/${T}_traffic_10gb.json --nchunks 9481
