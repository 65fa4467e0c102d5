# round 2, pass u: source-level stall profile of the 4-lane quality decoder at low residency, compact header on / off
mkdir -p gpurun_out
for k in 1 0; do
  SFQ_QLPC=4 SFQ_QCH=$k timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_qlt_decode' -c 1 -o gpurun_out/r2u_qd4_ch$k -f python bench.py --gb 0.13 --steps 1 --warmup 0 --no-cpu --no-e2e --no-extras > gpurun_out/r2u_ncu_ch$k.log 2>&1
  tail -2 gpurun_out/r2u_ncu_ch$k.log | cut -c1-200
done
ls -la gpurun_out/r2u_*.ncu-rep
