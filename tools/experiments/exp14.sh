# REDUX-based warp binning and small-subtree variants: correctness, then build timings
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_build.py C3 sah; python tools/prof_build.py C4 sah
for v in nocoop w128 t192; do echo $v; RTK_LIB=rtk_b200/librtk_b200_$v.so python tools/prof_build.py C3 sah; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 65536 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
print('bench | Mrays/s %.1f exact %s/%s build %.2f ms sah %.3f nodes %.2f tris %.2f'%(d['value'], p.get('bit_exact'), p.get('gpu_bruteforce_bit_exact'), d['build']['device_ms'], d['build']['sah_cost'], d['roofline']['per_ray']['wide_node_visits'], d['roofline']['per_ray']['triangle_tests']))"
python tools/prof_build.py C3 sah > gpurun_out/pb_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_build_r1g.csv python tools/prof_build.py C3 sah > gpurun_out/pb_ncu.log 2>&1
echo done
