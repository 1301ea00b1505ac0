python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 200 gpurun_out/r2f_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
def walk(o,p=''):
    if isinstance(o,dict):
        if o.get('stale'): print('STALE', p)
        for k,v in o.items(): walk(v,p+'/'+k)
walk(d)
"
