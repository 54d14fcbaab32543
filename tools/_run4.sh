python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_err.log; tail -5 gpurun_out/bench_err.log; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d.get('cpu_baseline',{}).get('value'))
for k,v in d['extra'].get('per_config',{}).items(): print(k, v['kernel'], round(v['evals_per_sec'],1), round(v['frac'],3), 'cpu', round(v['cpu_baseline']['value'],1), 'x', round(v['speedup_vs_cpu_all_threads'],1))
print(d['extra'].get('single_eval_latency'))
print(d['extra'].get('risk_neutral_sample_sharded'))
"
