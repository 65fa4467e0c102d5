# round-1 first GPU pass: parity tests, smoke, 10 GB bench, launch list, one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_10gb.json 2> gpurun_out/bench_10gb.err; tail -c 5000 gpurun_out/bench_10gb.json; tail -3 gpurun_out/bench_10gb.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --gb 1 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_decode' -s 6 -c 6 -o gpurun_out/prof_coders_r1 -f python bench.py --gb 0.25 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
