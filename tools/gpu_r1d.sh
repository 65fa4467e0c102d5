mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "illumina or edge or ont" > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
KERN='k_decode' COUNT=6 TAG=r1d GB=0.5 bash tools/gpu_ncu_full.sh 2>&1 | tail -3
GB=2 bash tools/gpu_launches.sh | grep -E "k_decode|k_gen_model|k_rc"
