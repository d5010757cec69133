python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>>gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
e=d.get('e2e') or {}
print('final | Mrays/s %.1f e2e %.1f exact %s build %.2f ms l2 %.0f reserved %s'%(d['value'], e.get('value',0), p.get('bit_exact'), d['build']['device_ms'], d['roofline']['l2']['peak'], d['config'].get('reserved_sms')))"
