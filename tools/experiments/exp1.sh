run() { python bench.py --steps 10 --warmup 3 --build-mode sah --no-cpu-baseline --e2e-steps 1 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', '| Mrays/s %.1f trace_ms %.2f nodes %.2f tris %.2f parity %s'%(d['value'], d['kernels_ms']['k_trace'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests'], d['parity']))
"; }
run --parity-rays 1024
run --lib rtk_b200/librtk_b200_m4.so --parity-rays 0
run --lib rtk_b200/librtk_b200_m5.so --parity-rays 0
run --lib rtk_b200/librtk_b200_m4p.so --parity-rays 0
run --lib rtk_b200/librtk_b200_m3p.so --parity-rays 0
run --cull 0 --parity-rays 1048576
run --cull 1 --parity-rays 1048576 --lib rtk_b200/librtk_b200_m4.so
