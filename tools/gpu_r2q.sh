# round 2, pass q: when the header encoder may start (SFQ_ENC_SCHED), one-wave compress
TAG=r2q TAILN=13 KNOBS="SFQ_TRACE=1,SFQ_ENC_SCHED=1 SFQ_TRACE=1,SFQ_ENC_SCHED=2 SFQ_TRACE=1,SFQ_ENC_SCHED=4" ARGS="--steps 2 --warmup 1 --no-cpu --no-extras --no-e2e --gb 10" bash tools/gpu_ab2.sh
