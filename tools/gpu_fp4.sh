#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/fp4_probe.py > gpurun_out/fp4_probe.json 2> gpurun_out/fp4_probe.err; echo "probe rc=$?"
timeout 600 python tools/fp4_check.py > gpurun_out/fp4_check.jsonl 2> gpurun_out/fp4_check.err; echo "check rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/fp4_probe.json'))
print({k:d[k] for k in d if k!='inexact_cases'}, 'inexact:', d['inexact_cases'][:6])
P
tail -n 5 gpurun_out/fp4_probe.err; cat gpurun_out/fp4_check.jsonl | cut -c1-400; tail -n 5 gpurun_out/fp4_check.err
