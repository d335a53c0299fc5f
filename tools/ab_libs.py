#!/usr/bin/env python
"""Same-box A/B of several builds of libstorm_b200.so (tools/build_variant.sh): the kernel is power-capped, so
boxes differ by up to 10 % and only timings taken on one device in interleaved rounds compare.

    python tools/ab_libs.py name1,name2,... [rows:bits ...]       (names under stormbitmaps_b200/_variants/)

Uses only entry points every build has (synth, pairw_device); JSON lines."""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1].split(",")
shapes = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(30000, 131072), (65536, 4096), (10000, 65536), (16384, 16384)]
libs = {}
for n in names:
    L = C.CDLL(os.path.join(ROOT, "stormbitmaps_b200", "_variants", f"lib_{n}.so"))
    L.STORM_b200_pairw_device.restype = C.c_int
    L.STORM_b200_pairw_device.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.STORM_b200_synth_geno_device.restype = C.c_int
    L.STORM_b200_synth_geno_device.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
    L.STORM_b200_last_error.restype = C.c_char_p
    libs[n] = L
torch.cuda.set_device(0)
for (n_rows, M) in shapes:
    W = (M + 63) // 64
    stride = (W + 15) // 16 * 16
    rows = torch.zeros((n_rows, stride), dtype=torch.int64, device="cuda")
    L0 = libs[names[0]]
    assert L0.STORM_b200_synth_geno_device(rows.data_ptr(), n_rows, W, stride, M, 1, 0, None) == 0
    torch.cuda.synchronize()
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    best = {n: 1e30 for n in names}
    tot = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rnd in range(7):                       # round 0 = warm-up (FP4 self-test, prefix upload)
        for n in names[rnd % len(names):] + names[:rnd % len(names)]:      # rotate: the part is power-capped and heats up
            L = libs[n]
            total.zero_()
            ev[0].record()
            rc = L.STORM_b200_pairw_device(rows.data_ptr(), n_rows, W, stride, 0, 1, 0, total.data_ptr(), None)
            ev[1].record()
            torch.cuda.synchronize()
            assert rc == 0, L.STORM_b200_last_error()
            tot[n] = int(total.item())
            if rnd:
                best[n] = min(best[n], ev[0].elapsed_time(ev[1]))
    wp = n_rows * (n_rows - 1) / 2 * W
    print(json.dumps({"rows": n_rows, "bits": M, "ms": {n: round(best[n], 5) for n in names},
                      "wp_per_s": {n: wp / best[n] * 1e3 for n in names}, "totals_agree": len(set(tot.values())) == 1}), flush=True)
    del rows
