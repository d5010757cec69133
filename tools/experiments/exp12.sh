# round-1 closing run: tests, the official bench lines, C5, launch lists, ncu --set full of k_trace
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>gpurun_out/bench_r1_ref.err; cut -c1-300 gpurun_out/bench_r1_ref.json
python bench.py > gpurun_out/bench_r1_n1.json 2>gpurun_out/bench_r1_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_n1.json').read())
print({k:d[k] for k in ('value','ms_per_step','e2e','kernels_ms','occlusion','parity','clocks','cpu_baseline')})
print(d['roofline']); print(d['build'])
PY
python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/exp.err > gpurun_out/bench_r1_c5.json; cut -c1-200 gpurun_out/bench_r1_c5.json
python bench.py --workload C4 --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/exp.err > gpurun_out/bench_r1_c4.json; cut -c1-200 gpurun_out/bench_r1_c4.json
python tools/prof_build.py C3 sah > gpurun_out/pb_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_build_r1e.csv python tools/prof_build.py C3 sah > gpurun_out/pb_ncu.log 2>&1
cat gpurun_out/pb_plain.log
python tools/prof_trace.py C3 4 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_r1e python tools/prof_trace.py C3 4 > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/prof_plain.log
