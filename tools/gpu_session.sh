#!/bin/bash
# One GPU lease: `gpurun -- bash tools/gpu_session.sh <step> [<step> ...]`; every step writes under gpurun_out/ with the
# round prefix.  Steps: smoke pytest microbench bench bench_ref sanitize ncu_list ncu_full multi
R=${ROUND:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${R}_gpus.txt 2>&1
for step in "$@"; do
  echo "=== $step $(date +%T)"
  case $step in
    smoke)      timeout 600 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log ;;
    pytest)     timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${R}_pytest.log ;;
    pytest_new) timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "${PYTEST_K}" > gpurun_out/${R}_pytest_new.log 2>&1; echo "pytest_new rc=$?"; tail -25 gpurun_out/${R}_pytest_new.log ;;
    microbench) timeout 300 python tools/microbench_report.py > gpurun_out/${R}_microbench.json 2> gpurun_out/${R}_microbench.err; echo "microbench rc=$?"; head -c 1500 gpurun_out/${R}_microbench.json ;;
    bench)      timeout 1200 python bench.py ${BENCH_ARGS} > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${R}_bench_n1.json; tail -5 gpurun_out/${R}_bench_n1.err ;;
    bench_ref)  timeout 900 python bench.py --impl reference > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "bench_ref rc=$?"; tail -c 800 gpurun_out/${R}_bench_reference_arm.json ;;
    sanitize)   for tool in memcheck racecheck; do
                  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/${R}_sanitizer_${tool}.log python tools/sanitize_smoke.py > gpurun_out/${R}_sanitizer_${tool}.out 2>&1
                  echo "sanitize $tool rc=$?"; tail -3 gpurun_out/${R}_sanitizer_${tool}.out; tail -4 gpurun_out/${R}_sanitizer_${tool}.log
                done ;;
    ncu_list)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_bench_c3_launches.csv \
                  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --verify-pairs 0 > gpurun_out/${R}_ncu_list.log 2>&1; echo "ncu_list rc=$?" ;;
    ncu_full)   timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dense_umma_kernel -s 1 -c 1 -o gpurun_out/${R}_fp4_c3 -f \
                  python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --verify-pairs 0 > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu_full rc=$?" ;;
    *)          echo "running: $step"; timeout 1800 bash -c "$step" ;;
  esac
done
echo "=== done $(date +%T)"
