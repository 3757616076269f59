python -m pytest tests/test_gpu_ttt.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
python bench.py --workload ttt4 --steps 1000 --warmup 20 --no-cpu 2>&1 | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print('ttt4 ms/step %.5f value %.3e frac %.3f e2e %.3e'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as e: print(t[-1500:])
"
