#!/bin/bash
# end-of-round evidence (after the twiddle-chain change): tests, bench line, launch list, full ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err
echo "bench rc=$?"; tail -n 2 gpurun_out/bench_final3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_final3.csv \
    python bench.py --steps 2 --warmup 3 --no-ref-gpu --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_final3.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 3 -c 2 -o gpurun_out/prof_n4096_final3 \
    python tools/run_one.py 4096 28 3 > gpurun_out/ncu_full_final3.log 2>&1
echo "ncu full rc=$?"
python tools/bench_configs.py > gpurun_out/bench_configs_g.log 2>&1; tail -n 11 gpurun_out/bench_configs_g.log
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline']['frac'],'e2e',l['e2e']['value'],'ref ms', l['vkfft_cuda_ref'].get('ms_sweep'), 'cpu', l['cpu_baseline']['value'], l['clocks'], l['roundtrip_rel_err'])
print({k:(v["ms_pair"], v["frac_of_peak"]) for k,v in l["per_n"].items()} if "per_n" in l else list(l.keys()))
PY
