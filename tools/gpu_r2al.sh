# round 2, closing pass on the final commit: smoke() and the default bench invocation (-> profiles/r2_final_bench_default.json)
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 600 python bench.py ) > gpurun_out/r2al_bench_default.json 2> gpurun_out/r2al_bench_default.err; tail -4 gpurun_out/r2al_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2al_bench_default.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "compress_GBps", "decompress_GBps", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["per_step_ms[c_total,c_code,d_total,d_gen,d_qlt,d_rec]"])
print([ (r["chunk_MiB"], r["value"], r.get("chunking_loss_pct")) for r in d["chunk_pareto"]["rows"]], {k: (v["value"], v.get("e2e")) for k, v in d["configs"].items()}, d["cpu_baseline"]["value"], d.get("extras_error"))
PY
