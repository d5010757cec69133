run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 262144 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('LANES=$RTK_B200_LANES $*', '| Mrays/s %.1f trace_ms %.2f nodes %.2f tris %.2f parity %s'%(d['value'], d['kernels_ms']['k_trace'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests'], d['parity']))
"; }
for L in 8 4 2; do export RTK_B200_LANES=$L; run; done
export RTK_B200_LANES=4; run --cull 0
export RTK_B200_LANES=2; run --cull 0
