#!/usr/bin/env python
"""Small, fast UMMA bring-up check (run on the GPU box): compares a few shapes with the POPC kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stormbitmaps_b200 as sb
sb.load()
cg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sb.set_umma_cta_group(cg)
ok = True
for (N, M, draws) in [(128, 128, 40), (256, 256, 100), (300, 4096, 1500), (600, 8192, 4000), (1000, 65536, 30000), (257, 192, 60)]:
    rows, W = sb.alloc_rows(N, M)
    sb.synth_uniform_device(rows, M, draws, 7)
    torch.cuda.synchronize()
    want, _ = sb.pairw_rect_device(rows, 0, N, 0, N, n_words=W, kernel="popc")
    try:
        got, tot = sb.pairw_rect_device(rows, 0, N, 0, N, n_words=W, kernel="umma")
        torch.cuda.synchronize()
    except Exception as e:
        print("FAIL launch", N, M, e); ok = False; break
    bad = (got != want).nonzero()
    print(f"cg={cg} N={N} M={M}: mismatches={bad.shape[0]} total_umma={int(tot.item())} total_popc={int(want.sum().item())}")
    if bad.shape[0]:
        ok = False
        i, j = bad[0].tolist()
        print("  first mismatch at", (i, j), "got", int(got[i, j]), "want", int(want[i, j]))
        print("  got[0,:8]", got[0, :8].tolist(), "want[0,:8]", want[0, :8].tolist())
        print("  got[1,:8]", got[1, :8].tolist(), "want[1,:8]", want[1, :8].tolist())
    t = sb.pairw_device(rows, n_words=W, kernel="umma")
    t2 = sb.pairw_device(rows, n_words=W, kernel="popc")
    torch.cuda.synchronize()
    print("   triangle totals umma/popc:", int(t.item()), int(t2.item()))
    ok &= int(t.item()) == int(t2.item())
print("UMMA_OK" if ok else "UMMA_BAD")
