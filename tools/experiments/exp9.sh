# PD + distributed ray preparation: correctness, variants, then ncu --set full of k_trace
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 65536 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
print('$RTK_B200_PD $*', '| Mrays/s %.1f trace_ms %.2f nodes %.2f tris %.2f exact %s/%s'%(d['value'], d['kernels_ms']['k_trace'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests'], p.get('bit_exact'), p.get('gpu_bruteforce_bit_exact')))
"; }
run
for v in asg1 asg2 asg5 part st8 st12 per3; do run --lib rtk_b200/librtk_b200_$v.so; done
run --workload C2 --rays 16588800
run --workload C4
run --cull 0
python tools/prof_trace.py C3 4 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_r1c python tools/prof_trace.py C3 4 > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/prof_plain.log
