# PD leaf phase: correctness on the GPU, then A/B against the old leaf phase and parameter variants
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 65536 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
print('$RTK_B200_PD $*', '| Mrays/s %.1f trace_ms %.2f nodes %.2f tris %.2f exact %s/%s'%(d['value'], d['kernels_ms']['k_trace'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests'], p.get('bit_exact'), p.get('gpu_bruteforce_bit_exact')))
"; }
run
RTK_B200_PD=0 run
for v in min2 min3 min6 min8 per2 per5 per8 asg1 asg5; do run --lib rtk_b200/librtk_b200_$v.so; done
run --workload C2 --rays 16588800
run --workload C4
run --cull 0
