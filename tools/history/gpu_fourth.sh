#!/bin/bash
mkdir -p gpurun_out
python tools/ktune.py 0 > gpurun_out/ktune_f32_b.log 2>&1; grep -E "kind=[012] n= *(8|16|32|64|128|256|512|1024|2048|4096|8192|16384) " gpurun_out/ktune_f32_b.log | cut -c1-140
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline']['frac'], 'step_frac', l['roofline']['step_frac'])
for n,v in l['per_n'].items(): print(n, v['ms_pair'], v['frac_of_peak'], l['vkfft_cuda_ref']['per_n'].get(n,{}).get('ms_pair'))
print('ref sweep ms', l['vkfft_cuda_ref'].get('ms_sweep'))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c2c_1d_f32 or full_size" 2>&1 | tail -3
