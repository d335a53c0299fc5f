#!/bin/bash
# FP4 form: probe table + peaks, parity suite, wave-sync A/B on C3 for both tensor forms, ncu of the FP4 kernel.
set -u
mkdir -p gpurun_out
timeout 300 python tools/fp4_probe.py > gpurun_out/fp4_probe.json 2> gpurun_out/fp4_probe.err; echo "probe rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
for k in umma fp4; do for ws in 0 1; do
  timeout 300 python bench.py --kernel $k --wave-sync $ws --no-cpu-baseline > gpurun_out/bench_${k}_ws${ws}.json 2> gpurun_out/bench_${k}_ws${ws}.err; echo "bench $k ws$ws rc=$?"
done; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_umma -c 1 -o gpurun_out/fp4_c3_full -f \
    python tools/prof_driver.py fp4 200000 131072 1 > gpurun_out/ncu_fp4.log 2>&1
python - <<'P'
import json
d=json.load(open('gpurun_out/fp4_probe.json'))
print({k:d[k] for k in d if k!='inexact_cases'}, 'inexact:', d['inexact_cases'][:6])
for k in ('umma','fp4'):
    for ws in (0,1):
        try:
            b=json.loads(open(f'gpurun_out/bench_{k}_ws{ws}.json').read().strip().split('\n')[-1])
            print(k, ws, 'value %.4g e2e %.4g ms %.1f frac %.3f clocks %s verified %s' % (b['value'], b['e2e']['value'], b['ms_per_step'], b['roofline']['frac'], b['clocks'], b['verified']['match']))
        except Exception as e:
            print(k, ws, 'failed', e)
P
