#!/bin/bash
# Multi-GPU pass (run with gpurun --gpus N): bench.py under torchrun exactly as the driver launches it.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_multi.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
for n in $(seq 1 $N); do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err; echo "n=1 rc=$?"
  elif [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err; echo "n=$n rc=$?"
    tail -n 3 gpurun_out/bench_scale_n$n.err
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_scale_n*.json')):
    try:
        b = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'n', b['n_gpus'], 'value %.4g e2e %.4g ms %.2f frac %.3f verified %s clocks %s' % (b['value'], b['e2e']['value'], b['ms_per_step'], b['roofline']['frac'], b['verified']['match'], b['clocks']))
    except Exception as e:
        print(f, 'failed', e)
P
