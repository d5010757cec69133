# Round 2 (2 GPUs, then 8): NCCL gather against the peer-memory gather (copy-engine pushes).
#   gpurun --gpus 2 --timeout 600 -- 'sh tools/experiments/exp20.sh 2'      (then the same with 8)
# gather_check.equal must be true in both modes.
N=${1:-2}
mkdir -p gpurun_out
for g in nccl p2p; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 0 --gather $g 2>>gpurun_out/exp20.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$N GPUs gather $g | Mrays/s %.1f ms/step %.3f k_trace %.3f | check %s'%(d['value'], d['ms_per_step'], d['kernels_ms']['k_trace'], d.get('gather_check')))"
done
