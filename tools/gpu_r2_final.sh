#!/bin/bash
# end-of-round evidence: the whole GPU suite, the bench line, the launch list, the half-precision benchmark program
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 | tee gpurun_out/pytest_gpu_final.log
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench rc=$?"; tail -n 2 gpurun_out/bench_r2_final.err
cd oracle/_ref
( timeout 300 ./VkFFT_TestSuite_b200 -vkfft 2 ) > ../../gpurun_out/sample2_half_b200.log 2>&1; echo "sample_2 b200 rc=$?"
cd ../..
grep -E "Benchmark score" gpurun_out/sample2_half_b200.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r2_final.csv \
    python bench.py --steps 2 --warmup 3 --no-ref-gpu --no-cpu --e2e-steps 1 --no-configs --no-sample0 > gpurun_out/bench_under_ncu_r2.log 2>&1
echo "ncu list rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline'].get('frac'),l['roofline'].get('kernel'),'e2e',l['e2e']['value'],'ref ms', l['vkfft_cuda_ref'].get('ms_sweep'), 'cpu', l['cpu_baseline']['value'], l['clocks'], l['roundtrip_rel_err'])
print({k:(v["ms_pair"], v["frac_of_peak"]) for k,v in l["per_n"].items()} if "per_n" in l else list(l.keys()))
print(l.get('sample0',{}).get('b200fft',{}).get('score'), l.get('sample0',{}).get('reference_vkfft_cuda',{}).get('score'))
for r in l.get('other_lengths', []): print(r)
for r in l.get('per_config', []): print(r)
PY
