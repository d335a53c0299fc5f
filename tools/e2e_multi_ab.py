#!/usr/bin/env python
"""Multi-GPU host-matrix query, same-box A/B of the exchange schedule (run under torchrun, N >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/e2e_multi_ab.py [rows bits]

Variants: `serial` (slice upload, one all-gather, shard kernel) and `pipelined` with {bands} x {reserved SMs}.
Per variant: best-of-3 wall time of distributed.pairw_total_from_host between barriers (max over ranks), wp/s,
and the device-resident shard time beside it.  JSON lines from rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import stormbitmaps_b200 as sb
from stormbitmaps_b200 import distributed as D

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 131_072
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
W = (bits + 63) // 64
rows_t, _ = sb.alloc_rows(rows, bits, device=dev)
sb.synth_geno_device(rows_t, bits, 20260117)
host = torch.empty((rows, W), dtype=torch.int64, pin_memory=True)
host.copy_(rows_t[:, :W])
torch.cuda.synchronize()
wp = rows * (rows - 1) / 2 * W
total = torch.zeros(1, dtype=torch.int64, device=dev)


def barrier():
    dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    best, val = 1e30, None
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        val = fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best, val


def resident():
    total.zero_()
    sb.pairw_device(rows_t, n_words=W, shard=rank, n_shards=world, total=total)
    dist.all_reduce(total)
    return int(total.item())


t_res, want = timed(resident)
if rank == 0:
    print(json.dumps({"variant": "device-resident", "world": world, "rows": rows, "bits": bits, "seconds": t_res,
                      "wp_per_s": wp / t_res, "total": want}), flush=True)
arena_s = D.alloc_gather_arena(rows, W, world, dev)
arena_p = D.alloc_stream_arena(rows, W, world, dev)
variants = [("serial", dict(pipelined=False, arena=arena_s))]
for bands in (4, 8, 16):
    for rsv in (0, 2, 4, 8):
        variants.append((f"pipelined bands={bands} reserved_sms={rsv}", dict(pipelined=True, arena=arena_p, bands=bands, reserved_sms=rsv)))
for name, kw in variants:
    t, got = timed(lambda: D.pairw_total_from_host(host, total=total, **kw))
    if rank == 0:
        print(json.dumps({"variant": name, "seconds": t, "wp_per_s": wp / t, "overhead_vs_resident_ms": (t - t_res) * 1e3,
                          "match": got == want}), flush=True)
dist.destroy_process_group()
