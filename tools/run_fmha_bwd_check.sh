timeout 200 python tools/debug_fmha_bwd.py > gpurun_out/$1_fmha_bwd_debug.log 2>&1; echo debug rc=$?; cut -c1-150 gpurun_out/$1_fmha_bwd_debug.log; timeout 400 python -m pytest tests/test_parity_sdpa.py -q -m gpu -k "backward" > gpurun_out/$1_sdpa.log 2>&1; echo sdpa rc=$?; tail -n 3 gpurun_out/$1_sdpa.log; timeout 300 python bench.py --workload sdpa_cfg5 > gpurun_out/$1_sdpa_cfg5.json 2> gpurun_out/$1_sdpa_cfg5.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/$1_sdpa_cfg5.json'))['sdpa']
for k in ('full','causal'): print(k, d[k]['ms'], d[k]['backward']['ms'], d[k]['backward']['tflops'])
"
