# round 2, pass w (two GPUs): the two-device tests, the driver's own 2-GPU launch of bench.py, and one file over two GPUs
mkdir -p gpurun_out
T=r2w
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_shapes.py tests/test_batch_driver.py -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
( time timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; tail -4 gpurun_out/${T}_bench_2gpu.err
( time timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 --single-file --gb 10 ) > gpurun_out/${T}_single_file_2gpu.json 2> gpurun_out/${T}_single_file_2gpu.err; tail -4 gpurun_out/${T}_single_file_2gpu.err
( time timeout 600 python bench.py --gpus 1 --steps 2 --warmup 1 --single-file --gb 10 ) > gpurun_out/${T}_single_file_1gpu.json 2> gpurun_out/${T}_single_file_1gpu.err; tail -4 gpurun_out/${T}_single_file_1gpu.err
python - <<PY
import json
for f in ("bench_2gpu", "single_file_2gpu", "single_file_1gpu"):
    try:
        d = json.loads([l for l in open("gpurun_out/${T}_%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "compress_GBps", "decompress_GBps", "ms_per_step", "n_gpus", "check")}, (d.get("e2e") or {}).get("value"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
