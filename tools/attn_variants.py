#!/usr/bin/env python3
"""One GPU call: where should P live between the softmax warps and P V?  For each variant of the tcgen05 attention
(VOX_CUDA_ATTN_P unset = shared memory, the default; tmem = tensor memory, 8 softmax warps; tmem16 = tensor memory, 16 softmax
warps) run the attention parity tests and time the encoder pass of the 60 s clip; then run the WHOLE GPU suite under the fastest
variant that passed and time the 10-minute clip with it.  Writes gpurun_out/<tag>_variants.json; a variant only becomes the default
(FA_P_DEFAULT_TMEM in vb_attn_tc.cu) if everything here is green for it."""
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02i"
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)


def run(cmd, env_extra, log, timeout):
    env = {**os.environ, **env_extra}
    t0 = time.time()
    with open(os.path.join(out_dir, log), "w") as f:
        try:
            rc = subprocess.call(cmd, env=env, stdout=f, stderr=subprocess.STDOUT, timeout=timeout, cwd=ROOT)
        except subprocess.TimeoutExpired:
            rc = -9
    return rc, time.time() - t0


def encoder_ms(log):
    txt = open(os.path.join(out_dir, log)).read()
    m = re.findall(r"pass (\d+): encoder ([0-9.]+) ms for (\d+) positions .* ids md5 (\w+)", txt)
    return [(int(p), float(ms), int(n), h) for p, ms, n, h in m]


res = {}
variants = [("smem", {}), ("tmem", {"VOX_CUDA_ATTN_P": "tmem"}), ("tmem16", {"VOX_CUDA_ATTN_P": "tmem16"})]
for name, env in variants:
    r = {}
    if name != "smem":
        rc, dt = run([sys.executable, "-m", "pytest", "tests/test_gpu_ops_parity.py", "-q", "-m", "gpu", "-k", "attention", "-x"], env,
                     f"{tag}_{name}_ops.log", 240)
        r["attention_tests_rc"], r["attention_tests_s"] = rc, round(dt, 1)
        if rc != 0:
            res[name] = r
            continue
    rc, dt = run([sys.executable, "tools/encoder_ab.py", "60", "3", "0"], env, f"{tag}_{name}_ab60.log", 240)
    r["ab60_rc"] = rc
    r["ab60"] = encoder_ms(f"{tag}_{name}_ab60.log")
    res[name] = r
    print(name, r, flush=True)

ok = {n: min(ms for p, ms, _, _ in r["ab60"] if p >= 1) for n, r in res.items() if r.get("ab60_rc") == 0 and len(r.get("ab60", [])) >= 2}
md5 = {n: {h for _, _, _, h in res[n]["ab60"]} for n in ok}
best = min((n for n in ok if n != "smem" and md5[n] == md5.get("smem", md5[n])), key=lambda n: ok[n], default=None)
res["encoder_ms_60s"] = ok
res["best_candidate"] = best
print("encoder ms (60 s clip):", ok, "best candidate:", best, flush=True)
if best is not None and ok[best] < 0.97 * ok.get("smem", 1e9):
    env = dict(variants)[best]
    rc, dt = run([sys.executable, "-m", "pytest", "tests", "-q", "-m", "gpu", "-rxs"], env, f"{tag}_{best}_full.log", 420)
    tail = open(os.path.join(out_dir, f"{tag}_{best}_full.log")).read().strip().splitlines()[-3:]
    res["full_suite"] = {"variant": best, "rc": rc, "seconds": round(dt, 1), "tail": tail}
    print("full suite under", best, "rc", rc, tail, flush=True)
    rc, dt = run([sys.executable, "tools/encoder_ab.py", "600", "2", "0"], env, f"{tag}_{best}_ab600.log", 240)
    res["ab600"] = {"variant": best, "rc": rc, "passes": encoder_ms(f"{tag}_{best}_ab600.log")}
    print("10-minute clip:", res["ab600"], flush=True)
json.dump(res, open(os.path.join(out_dir, f"{tag}_variants.json"), "w"), indent=1)
