# full ncu capture (source-level) of selected kernels on a small workload
mkdir -p gpurun_out
KERN=${KERN:-'k_gen_model|k_qlt_scatter|k_rc_encode|k_qlt_model'}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERN" -c ${COUNT:-5} -o gpurun_out/prof_${TAG:-x} -f python bench.py --gb ${GB:-0.5} --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
