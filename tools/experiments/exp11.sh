# leaf slots + stack fast path (main) vs fast path only (fp); presort upside; e2e floor
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --parity-rays 65536 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
e=d.get('e2e') or {}
print('T=$RTK_B200_HOST_THREADS skip=$RTK_B200_SKIP_PLACE chunk=$RTK_B200_HOST_CHUNK_LOG2 $*', '| Mrays/s %.1f trace_ms %.2f e2e %.1f (%.2f ms, rows ok %s) exact %s/%s build %.2f ms'%(d['value'], d.get('kernels_ms',{}).get('k_trace',0), e.get('value',0), e.get('ms_per_step',0), e.get('rows_equal_device_path'), p.get('bit_exact'), p.get('gpu_bruteforce_bit_exact'), d['build']['device_ms']))
"; }
run
run --lib rtk_b200/librtk_b200_fp.so
run --presort
RTK_B200_SKIP_PLACE=1 run
RTK_B200_HOST_THREADS=16 run
RTK_B200_HOST_THREADS=16 RTK_B200_HOST_CHUNK_LOG2=19 run
RTK_B200_HOST_THREADS=16 RTK_B200_HOST_CHUNK_LOG2=21 run
run --workload C2 --rays 16588800
run --workload C4
run --build-mode lbvh
nproc; grep -m1 "model name" /proc/cpuinfo; free -g | head -2
