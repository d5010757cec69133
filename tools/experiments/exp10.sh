# node layout v2 (256-bit child loads), dense-row host pipeline, C5 wavefront; then ncu of k_trace
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --parity-rays 65536 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
e=d.get('e2e') or {}
print('L=$RTK_B200_LANES T=$RTK_B200_HOST_THREADS $*', '| Mrays/s %.1f trace_ms %.2f e2e %.1f (%.2f ms, rows ok %s) exact %s/%s build %.2f ms'%(d['value'], d.get('kernels_ms',{}).get('k_trace',0), e.get('value',0), e.get('ms_per_step',0), e.get('rows_equal_device_path'), p.get('bit_exact'), p.get('gpu_bruteforce_bit_exact'), d['build']['device_ms']))
"; }
run
RTK_B200_LANES=4 run
run --lib rtk_b200/librtk_b200_st12.so
run --lib rtk_b200/librtk_b200_st8.so
run --lib rtk_b200/librtk_b200_fp.so
run --presort
RTK_B200_HOST_THREADS=4 run
RTK_B200_HOST_THREADS=16 run
run --workload C2 --rays 16588800
run --workload C4
run --cull 0
python bench.py --workload C5 --steps 3 --warmup 1 --no-cpu-baseline 2>>gpurun_out/exp.err | tee gpurun_out/c5.json | cut -c1-400
python tools/prof_trace.py C3 4 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_r1d python tools/prof_trace.py C3 4 > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/prof_plain.log
