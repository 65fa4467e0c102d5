# round 2, final pass: smoke(), the default invocation of both bench arms on the final build (-> profiles/r2_final_*)
mkdir -p gpurun_out
T=r2ag
( time python __graft_entry__.py --smoke ) > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
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
    for k in ("chunk_pareto", "configs", "cpu_baseline", "parity_sampled", "extras_error"):
        if k in d: print("  ", k, json.dumps(d[k])[:1500])
    if "e2e" in d: print("   e2e", json.dumps(d["e2e"])[:900])
PY
