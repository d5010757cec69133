run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 65536 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('LANES=$RTK_B200_LANES $*', '| Mrays/s %.1f trace_ms %.2f nodes %.2f leaves %.2f tris %.2f ok %s'%(d['value'], d['kernels_ms']['k_trace'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['leaf_visits'], d['roofline']['per_ray']['triangle_tests'], d['parity']['gpu_bruteforce_bit_exact'] and d['parity']['bit_exact']))
"; }
export RTK_B200_LANES=2
run --lib rtk_b200/librtk_b200_p2.so
run --lib rtk_b200/librtk_b200_p3.so
run --lib rtk_b200/librtk_b200_p4.so
export RTK_B200_LANES=4
run --lib rtk_b200/librtk_b200_p2.so
run --lib rtk_b200/librtk_b200_p3.so
