# round 2, pass ad: k_assemble with 8 records per warp, ranks bound to the GPU's NUMA node
nvidia-smi topo -m 2>/dev/null | head -8; lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8
cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head -5
TAG=r2ad TESTS="tests/test_gpu_parity.py tests/test_plane_hooks.py tests/test_reference_format.py" TAILN=0 KNOBS="A=1" ARGS="--steps 3 --warmup 1 --no-cpu --no-extras --gb 10" bash tools/gpu_ab2.sh
python - <<PY
import json
d = json.loads(open("gpurun_out/r2ad_A=1.json").read().strip().splitlines()[-1])
print(d["config"].get("host_numa_node"), d["phases_ms_per_step"]["d_pack"], d["e2e"]["per_step_ms[c_total,c_code,d_total,d_gen,d_qlt,d_rec]"])
PY
