#!/usr/bin/env python3
"""Host-side nibble packing rate (bb_pack_nibbles) vs thread count, pinned and pageable buffers: python tools/pack_bench.py"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np, torch
    import barbell_b200 as bb
    n = 1 << 30
    src = torch.randint(65, 85, (n,), dtype=torch.uint8)
    for kind in ("pageable", "pinned"):
        s = src.pin_memory() if kind == "pinned" else src
        d = torch.empty(n // 2 + 64, dtype=torch.uint8)
        d = d.pin_memory() if kind == "pinned" else d
        L = bb.lib()
        L.bb_pack_nibbles(s.data_ptr(), n, d.data_ptr())
        t0 = time.perf_counter()
        for _ in range(5):
            L.bb_pack_nibbles(s.data_ptr(), n, d.data_ptr())
        dt = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter(); d2 = s.clone(); dc = time.perf_counter() - t0
        print(f"threads={os.environ.get('BB_PACK_THREADS')} {kind}: pack {n / dt / 1e9:.1f} GB/s of input; single-thread clone {n / dc / 1e9:.1f} GB/s", flush=True)
else:
    print("cpus:", os.cpu_count())
    os.system("lscpu | egrep 'Model name|Socket|Thread|Core|NUMA node\\(s\\)|Flags' | cut -c1-200 | sed 's/Flags.*avx512[a-z_]*.*/Flags: has avx512/'")
    for t in (1, 2, 4, 8, 16, 32):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, BB_PACK_THREADS=str(t)))
