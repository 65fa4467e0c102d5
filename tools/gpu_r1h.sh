# r1b evidence: GPU parity tests, full 10 GB bench, reference arm, launch list, one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_10gb.json 2> gpurun_out/bench_10gb.err; tail -c 5000 gpurun_out/bench_10gb.json; tail -3 gpurun_out/bench_10gb.err
GB=1 bash tools/gpu_launches.sh
KERN='k_gen_model|k_qlt_scatter|k_rc_encode|k_qlt_model|k_decode' COUNT=8 TAG=r1b GB=0.25 bash tools/gpu_ncu_full.sh 2>&1 | tail -3
