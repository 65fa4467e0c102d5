# round 2, pass aa: single-pass ncu metrics of the DECODE launches of the benched 10 GB step (the full-step pass r2y ran
# out of time behind the one-wave compress kernels), merged with r2y's compress rows into profiles/r2_traffic_10gb.json
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none --cache-control none -k regex:'k_qlt_decode|k_decode|k_gen_exceptions|k_assemble|k_out_|k_pack|k_blob' --csv --log-file gpurun_out/r2aa_ncu_10gb_decode.csv python bench.py --gb 10 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/r2aa_ncu.log 2>&1
tail -2 gpurun_out/r2aa_ncu.log | cut -c1-300
grep -c "k_qlt_decode" gpurun_out/r2aa_ncu_10gb_decode.csv
