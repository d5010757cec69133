# upload stream ahead of the kernels: tests, official bench line, build variant timing
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_n1.json 2>gpurun_out/bench_r1_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step','e2e','kernels_ms','occlusion','parity','clocks')})
print(d['build'])
PY
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 5 --parity-rays 0 "$@" 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
e=d.get('e2e') or {}
print('T=$RTK_B200_HOST_THREADS skip=$RTK_B200_SKIP_PLACE $*', '| Mrays/s %.1f e2e %.1f (%.2f ms, rows ok %s)'%(d['value'], e.get('value',0), e.get('ms_per_step',0), e.get('rows_equal_device_path')))
"; }
RTK_B200_SKIP_PLACE=1 run
RTK_B200_HOST_THREADS=6 run
RTK_B200_HOST_THREADS=16 run
run --workload C4
python tools/prof_build.py C3 sah; RTK_LIB=rtk_b200/librtk_b200_hyb.so python tools/prof_build.py C3 sah; RTK_LIB=rtk_b200/librtk_b200_hyb.so python tools/prof_build.py C4 sah
