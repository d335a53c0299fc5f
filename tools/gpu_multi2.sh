#!/bin/bash
# Scaling pass (run with gpurun --gpus N): bench.py under torchrun exactly as the driver launches it,
# largest rank count first (it is the one that has never run), then the smaller ones on the same box.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_multi2.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
for n in 8 4 2; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err; echo "n=$n rc=$?"
  tail -n 3 gpurun_out/bench_scale_n$n.err | cut -c1-300
done
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err; echo "n=1 rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_scale_n*.json')):
    try:
        b = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'n', b['n_gpus'], 'value %.4g e2e %.4g ms %.2f frac %.3f verified %s clocks %s' % (b['value'], b['e2e']['value'], b['ms_per_step'], b['roofline']['frac'], b['verified']['match'], b['clocks']))
    except Exception as e:
        print(f, 'failed', e)
P
