# round 2, pass b: source-level stall profile of the two decoder chains at low residency (124 chunks: one warp per SM)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_qlt_decode|k_decode' -c 3 -o gpurun_out/r2b_dec_lowres -f python bench.py --gb 0.13 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
