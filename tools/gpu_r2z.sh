# round 2, pass z: level-4 e2e (the configs[2] leg printed 4.1 GB/s e2e against 12.3 device-resident), head-part size sweep
TAG=r2z TAILN=0 KNOBS="A=1 SFQ_HEAD_FRAC=0.05 SFQ_HEAD_FRAC=0.1 SFQ_HEAD_FRAC=0.25" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --gb 10" bash tools/gpu_ab2.sh
TAG=r2z_l4 TAILN=0 KNOBS="SFQ_MARKS=1" ARGS="--steps 2 --warmup 1 --no-cpu --no-extras --gb 10 --level 4" bash tools/gpu_ab2.sh
grep "sfq host" gpurun_out/r2z_l4_SFQ_MARKS=1.err | tail -40
