# round 2, pass ab: header coder's scratch in global memory instead of shared memory (SFQ_REC_GLOBAL=1)
TAG=r2ab TESTS="tests/test_gpu_parity.py" TAILN=12 KNOBS="SFQ_TRACE=1,SFQ_REC_GLOBAL=1 A=1" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
