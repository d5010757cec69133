# build-kernel changes (bin ILP, one-pass partition, 4 radix passes in SAH mode): tests, timings, launch lists
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_build.py C3 sah; python tools/prof_build.py C4 sah; python tools/prof_build.py C3 lbvh
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --parity-rays 65536 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
e=d.get('e2e') or {}
print('$*', '| Mrays/s %.1f trace_ms %.2f e2e %.1f (%.2f ms) exact %s/%s build %.2f ms sah %.3f nodes %.2f tris %.2f'%(d['value'], d.get('kernels_ms',{}).get('k_trace',0), e.get('value',0), e.get('ms_per_step',0), p.get('bit_exact'), p.get('gpu_bruteforce_bit_exact'), d['build']['device_ms'], d['build']['sah_cost'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests']))
"; }
run
run --workload C4
python tools/prof_build.py C3 sah > gpurun_out/pb_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_build_r1f.csv python tools/prof_build.py C3 sah > gpurun_out/pb_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bstep_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bstep_ncu.log 2>&1
echo done
