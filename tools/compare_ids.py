#!/usr/bin/env python3
"""Compare greedy ids saved from an engine run (bench.py with VOX_BENCH_SAVE_IDS=1, tools/dump_ids.py) with a reference fixture.
   python tools/compare_ids.py tests/golden/synth_s600_oneshot.npz profiles/r02_ids_600s_default.npy [more .npy ...]"""
import sys

import numpy as np

g = np.load(sys.argv[1])
ref = g["tokens"]
margin = g["top_val"][:, 0] - g["top_val"][:, 1]
runner_up = g["top_idx"][:, 1]
print(f"reference: {len(ref)} ids; smallest top-2 margins: {np.sort(margin)[:5]}")
for fn in sys.argv[2:]:
    ids = np.load(fn)
    n = min(len(ids), len(ref))
    bad = np.nonzero(ids[:n] != ref[:n])[0]
    print(f"{fn}: {len(ids)} ids, {n - bad.size} of {n} equal")
    for i in bad[:20]:
        print(f"   step {i}: reference {ref[i]} (top-2 margin {margin[i]:.3e}, runner-up {runner_up[i]}), engine {ids[i]}")
