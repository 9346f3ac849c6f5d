#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_c.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline']['frac'], 'step_frac', l['roofline']['step_frac'], 'e2e', l['e2e']['value'], 'rt', l['roundtrip_rel_err'])
for n,v in l['per_n'].items(): print(n, v['ms_pair'], v['frac_of_peak'], l['vkfft_cuda_ref']['per_n'].get(n,{}).get('ms_pair'))
print('ref sweep ms', l['vkfft_cuda_ref'].get('ms_sweep'), 'cpu', l['cpu_baseline']['value'])
PY
python tools/ktune.py 1 > gpurun_out/ktune_f64.log 2>&1; cut -c1-130 gpurun_out/ktune_f64.log | grep kind
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu3.log
