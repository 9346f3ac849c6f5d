#!/bin/bash
# end-of-round evidence: bench line, launch list of the same command, full ncu capture of the dominant kernel
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-ref-gpu --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_final.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 3 -c 2 -o gpurun_out/prof_n4096_final \
    python tools/run_one.py 4096 28 3 > gpurun_out/ncu_full_final.log 2>&1
echo "ncu full rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline']['frac'],'e2e',l['e2e']['value'],'ref ms', l['vkfft_cuda_ref'].get('ms_sweep'), 'cpu', l['cpu_baseline']['value'], l['clocks'])
PY
