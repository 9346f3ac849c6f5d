#!/bin/bash
# end-of-round evidence: tests, bench line, launch list of the same command, full ncu capture of a Four-Step pair, sanitizer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
echo "bench rc=$?"; tail -n 2 gpurun_out/bench_final2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_final2.csv \
    python bench.py --steps 2 --warmup 3 --no-ref-gpu --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_final2.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 4 -c 2 -o gpurun_out/prof_fourstep_2p20_final2 \
    python tools/run_one.py 1048576 28 3 > gpurun_out/ncu_full_final2.log 2>&1
echo "ncu full rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck.log
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms_step',l['ms_per_step'],'roofline',l['roofline']['frac'],'e2e',l['e2e']['value'],'ref ms', l['vkfft_cuda_ref'].get('ms_sweep'), 'cpu', l['cpu_baseline']['value'], l['clocks'])
PY
