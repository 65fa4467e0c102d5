# lighter ncu capture (source counters + warp states) of selected kernels
mkdir -p gpurun_out
timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section InstructionStats --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"${KERN:-k_decode}" -s ${SKIP:-0} -c ${COUNT:-2} -o gpurun_out/prof_${TAG:-x} -f python bench.py --gb ${GB:-0.25} --chunk ${CHUNK:-1048576} --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_sec.log 2>&1
tail -2 gpurun_out/ncu_sec.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
