# round 2, pass x: quality keys+scan ahead of the base path (SFQ_ENC_ORDER=1) with the header coder held back or not; shard+parts test
TAG=r2x TESTS="tests/test_gpu_shapes.py" TAILN=12 KNOBS="SFQ_TRACE=1,SFQ_ENC_ORDER=1 SFQ_TRACE=1,SFQ_ENC_ORDER=1,SFQ_ENC_SCHED=4 SFQ_TRACE=1,SFQ_ENC_ORDER=1,SFQ_ENC_SCHED=1" ARGS="--steps 2 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
