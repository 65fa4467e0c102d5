# round 2, pass r: header decoder after the base decoder (SFQ_DEC_SCHED=1), same-symbol model prefetch in the quality decoder (SFQ_QSPEC=1)
TAG=r2r TAILN=0 KNOBS="A=1 SFQ_DEC_SCHED=1 SFQ_QSPEC=1 SFQ_DEC_SCHED=1,SFQ_QSPEC=1 SFQ_DEC_SCHED=1,SFQ_GEN_AHEAD2=0" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
