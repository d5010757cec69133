# two ranks on one box: the weak-scaling bench line and the C5 wavefront, as the driver launches them
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_n2.json').read())
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','kernels_ms','parity','clocks')})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload C5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r1_c5_n2.json 2>> gpurun_out/bench_r1_n2.err
cut -c1-260 gpurun_out/bench_r1_c5_n2.json
tail -3 gpurun_out/bench_r1_n2.err
