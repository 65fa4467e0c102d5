# round 2, pass ak: quality decoder keeps its frequencies packed between steps (one add into a packed word instead of a 16-way select + repacking)
TAG=r2ak TESTS="tests/test_gpu_shapes.py tests/test_gpu_parity.py" TAILN=0 KNOBS="A=1 SFQ_SERIAL_ROLES=1" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
