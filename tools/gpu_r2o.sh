# round 2, pass o: one-wave compress (partition lists and quality steps share memory), head part of the pipelined sfq_compress,
# decoder CTAs at most one per SM - against the previous behaviour; host marks of both entry points
TAG=r2o TESTS="tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_plane_hooks.py" TAILN=60 \
KNOBS="SFQ_MARKS=1 SFQ_ALIAS=0,SFQ_DEC_FIT=0,SFQ_PARTS=2 SFQ_DEC_FIT=0" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --gb 10" bash tools/gpu_ab2.sh
