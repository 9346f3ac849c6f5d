#!/bin/bash
# round 2, first GPU call of the last session: one-launch Bluestein correctness + A/B timing, the bench line, launch list, ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_curated_lengths.py -m gpu -x -q -k "bluestein or non_pow2" 2>&1 | tail -n 4
timeout 600 python tools/bench_bluestein.py > gpurun_out/bench_bluestein.log 2>&1; echo "bluestein rc=$?"; cat gpurun_out/bench_bluestein.log
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench rc=$?"; tail -n 2 gpurun_out/bench_r2_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r2_final.csv \
    python bench.py --steps 2 --warmup 3 --no-ref-gpu --no-cpu --e2e-steps 1 --no-configs --no-sample0 > gpurun_out/bench_under_ncu_r2.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 6 -c 2 -o gpurun_out/prof_2p20_r2_final \
    python tools/run_one.py 1048576 28 3 > gpurun_out/ncu_full_2p20_r2.log 2>&1
echo "ncu full 2^20 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 3 -c 1 -o gpurun_out/prof_blue509_r2 \
    python tools/run_one.py 509 28 3 > gpurun_out/ncu_full_blue509_r2.log 2>&1
echo "ncu full bluestein 509 rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline'].get('frac'),l['roofline'].get('kernel'),'e2e',l['e2e']['value'],'ref ms', l['vkfft_cuda_ref'].get('ms_sweep'), 'cpu', l['cpu_baseline']['value'], l['clocks'], l['roundtrip_rel_err'])
print({k:(v["ms_pair"], v["frac_of_peak"]) for k,v in l["per_n"].items()} if "per_n" in l else list(l.keys()))
print(l.get('sample0',{}).get('b200fft',{}).get('score'), l.get('sample0',{}).get('reference_vkfft_cuda',{}).get('score'))
PY
