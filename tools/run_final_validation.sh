#!/bin/bash
# usage: bash tools/run_final_validation.sh <tag>  -- what the driver runs at round end: the whole GPU suite, smoke(), the default bench line, the reference arm
T=$1
timeout 1700 python -m pytest tests -x -q -m gpu --durations=25 > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?; tail -n 32 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$?; tail -n 2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo ref rc=$?; cut -c1-300 gpurun_out/${T}_bench_reference.json
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','ms_per_step','gpu_launches','clocks')}); print('e2e', d['e2e']); print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','traffic','per_command_bound_ms','frac_of_per_command_bound')}); print('cpu_baseline', d['cpu_baseline'])
for k,v in d.get('variants',{}).items(): print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','unit')} if isinstance(v,dict) else v)
"
