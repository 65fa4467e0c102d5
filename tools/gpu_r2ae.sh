# round 2, pass ae: quality coder chain with four lanes per chunk-stream (k_rc_encode_q4) against the one-lane form
TAG=r2ae TESTS="tests/test_gpu_parity.py tests/test_gpu_shapes.py" TAILN=12 KNOBS="SFQ_TRACE=1 SFQ_TRACE=1,SFQ_RC_Q4=0" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
