# round 2, pass s: compact-header quality decoder (2 dirty sectors per quality instead of 5) against the lane-owned layout
TAG=r2s TESTS="tests/test_gpu_shapes.py tests/test_gpu_parity.py" TAILN=0 KNOBS="A=1 SFQ_QCH=0 SFQ_DEC_SCHED=1" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
