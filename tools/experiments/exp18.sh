# 2 ranks: does keeping SMs free for NCCL's kernels pay?
for r in 0 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$r bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 0 --reserve-sms $r 2>>gpurun_out/exp18.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('reserve $r | Mrays/s %.1f ms/step %.3f k_trace %.3f'%(d['value'], d['ms_per_step'], d['kernels_ms']['k_trace']))"
done
