# round 2, pass p: traces of the one-wave and two-wave compress, decoder CTA fitting A/B
TAG=r2p TAILN=14 KNOBS="SFQ_TRACE=1 SFQ_TRACE=1,SFQ_ALIAS=0 SFQ_DEC_FIT=0" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --gb 10" bash tools/gpu_ab2.sh
