#!/bin/bash
mkdir -p gpurun_out
for mode in two full tile; do
  echo "== TW mode $mode"
  B200FFT_TW_MODE=$mode python tools/ktune.py 0 16,32,64,128,256,512,1024,2048 2>&1 | grep -E "kind=2 .* v=0" | cut -c1-120
done
echo "== bench with full-table phases"
( B200FFT_TW_MODE=full timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-ref-gpu ) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'], 'rt_err', l['roundtrip_rel_err'])
print({n: v['ms_pair'] for n,v in l['per_n'].items()})
PY
