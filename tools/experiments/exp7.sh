run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 0 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', '| Mrays/s %.1f trace_ms %.2f nodes %.2f tris %.2f'%(d['value'], d['kernels_ms']['k_trace'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests']))
"; }
run --presort
run --workload C2 --rays 16588800
run --presort --cull 0
