# round 2, pass a: where the round-1 kernels stand at larger chunks, at one resident chunk, and what the
# benched 10 GB launches really move (single-pass ncu metrics: no replay)
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
for ch in 1 2 4 8; do
  SFQ_TRACE=1 $B --gb 10 --chunk $((ch << 20)) > gpurun_out/r2a_chunk${ch}m.json 2> gpurun_out/r2a_chunk${ch}m.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2a_chunk${ch}m.json").read().strip().splitlines()[-1])
print("chunk ${ch} MiB", {k: d[k] for k in ("value", "compress_GBps", "decompress_GBps", "stream_ratio")}); p = d["phases_ms_per_step"]
print({k: p[k] for k in ("c_code", "c_gen", "c_qlt", "c_rec", "d_code", "d_gen", "d_qlt", "d_rec")}, d["chain"])
PY
  grep "sfq trace" gpurun_out/r2a_chunk${ch}m.err | tail -12
done
for gb in 0.0011 0.009; do
  $B --gb $gb --steps 3 > gpurun_out/r2a_probe_$gb.json 2> gpurun_out/r2a_probe_$gb.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2a_probe_$gb.json").read().strip().splitlines()[-1])
p = d["phases_ms_per_step"]; print("probe $gb GB", d["chain"], {k: p[k] for k in ("c_gen", "c_qlt", "c_rec", "d_gen", "d_qlt", "d_rec")})
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct --clock-control none -k regex:'k_qlt_decode|k_decode|k_gen_model|k_rc_encode|k_qlt_|k_encode' --launch-skip 12 -c 14 --csv --log-file gpurun_out/r2a_ncu_10gb_metrics.csv python bench.py --gb 10 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log | cut -c1-300
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2a_ncu_10gb_metrics.csv")) if len(r) > 10]
for r in rows[:400]:
    if r[0].isdigit(): print(r[0], r[4][:40], r[-3], r[-2], r[-1])
PY
