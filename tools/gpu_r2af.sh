# round 2, pass af: base coder chain in 4-warp CTAs sharing one reciprocal table (frees 24 KB of shared memory per SM for its neighbours)
TAG=r2af TESTS="tests/test_gpu_parity.py" TAILN=12 KNOBS="SFQ_TRACE=1 SFQ_TRACE=1,SFQ_RC_WARPS=1" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
