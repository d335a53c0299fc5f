#!/usr/bin/env python
"""A/B the code variants of the dense kernels on the GPU box (results -> stdout as JSON lines).

    python tools/umma_variants.py [rows] [bits]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 131072


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        fn(); ev[k + 1].record()
    torch.cuda.synchronize()
    return min(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))


# correctness shape: ragged rows and words, checked against the CUDA-core kernel
ts, Ws = sb.alloc_rows(1531, 8192 * 3 + 64)
sb.synth_geno_device(ts, 8192 * 3 + 64, 5)
ref_small = int(sb.pairw_device(ts, n_words=Ws, kernel="popc").item())

t, W = sb.alloc_rows(rows, bits)
sb.synth_geno_device(t, bits, 1)
total = torch.zeros(1, dtype=torch.int64, device="cuda")
wp = rows * (rows - 1) / 2 * W
ref = None
for cg in (2, 1):
    sb.set_umma_cta_group(cg)
    for var in (0, 1, 2, 3):
        sb.set_umma_variant(var)
        ok_small = int(sb.pairw_device(ts, n_words=Ws, kernel="umma").item()) == ref_small

        def run():
            total.zero_()
            sb.pairw_device(t, n_words=W, kernel="umma", total=total)
        ms = timed(run)
        got = int(total.item())
        if ref is None:
            ref = got
        print(json.dumps({"kernel": "umma", "cg": cg, "variant": var, "rows": rows, "bits": bits, "ms": ms,
                          "wp_per_s": wp / (ms * 1e-3), "tops": wp * 128 / (ms * 1e-3) / 1e12,
                          "match_small": ok_small, "match_big": got == ref}), flush=True)
sb.set_umma_cta_group(2)

# CUDA-core kernels on a smaller matrix (they are ~12x slower)
r2 = min(rows, 12000)
t2, W2 = sb.alloc_rows(r2, bits)
sb.synth_geno_device(t2, bits, 1)
wp2 = r2 * (r2 - 1) / 2 * W2
vals = {}
for k in ("popc", "csa"):
    def run():
        total.zero_()
        sb.pairw_device(t2, n_words=W2, kernel=k, total=total)
    ms = timed(run)
    vals[k] = int(total.item())
    print(json.dumps({"kernel": k, "rows": r2, "bits": bits, "ms": ms, "wp_per_s": wp2 / (ms * 1e-3),
                      "wp_per_clk_per_sm_at_1965": wp2 / (ms * 1e-3) / 1.965e9 / 148,
                      "match": vals[k] == vals["popc"]}), flush=True)
