python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 1500 gpurun_out/r02_bench_a.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r02_bench_a.json').read().strip().splitlines()[-1])
    print('value',l['value'],'ms',l['ms_per_step'],'frac',l['roofline']['frac'],'kernel_ms',l['roofline']['kernel_ms'],'e2e',l['e2e']['value'], l['e2e']['ms_per_step'])
    print('host_rows',l['e2e_host_rows'])
    for k,v in l['extra'].items():
        if 'error' in v: print(k,v); continue
        print(k,'value',v['value'],'ms',v['ms_per_step'],'frac',v['roofline']['frac'],'kernel_ms',v['roofline']['kernel_ms'],'e2e',v['e2e']['value'],'cpu',v['cpu_baseline']['value'])
except Exception as e: print('ERR',e)
PY
