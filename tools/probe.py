"""Quick device probe: per-kernel device times of one v2 (and v1) mastering pass over a random batch."""
import argparse
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "audio-mastering-web_b200"))

import numpy as np
import torch

from mm_b200 import _lib, pipeline as P
from mm_b200.engine import get_engine, style_struct


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=8)
    ap.add_argument("--sec", type=float, default=180.0)
    ap.add_argument("--sr", type=int, default=44100)
    ap.add_argument("--chain", default="v2")
    ap.add_argument("--style", default="standard")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--compressor", default="soft_knee", help="soft_knee | envelope")
    a = ap.parse_args()
    eng = get_engine()
    n = int(a.sec * a.sr)
    b = eng.empty(a.tracks, 2, n, a.sr)
    with torch.cuda.stream(eng.stream):
        b.t.normal_(0.0, 0.1)
        b.t.add_(0.001)
    out = eng.like(b)
    sts = [style_struct(P.STYLE_CONFIGS[a.style], P.STYLE_CONFIGS[a.style]["lufs"]) for _ in range(a.tracks)]
    chain = _lib.CHAIN_V1 if a.chain == "v1" else _lib.CHAIN_V2
    for rep in range(a.reps):
        eng.timing(True)
        t0 = time.time()
        eng.master(b, chain, sts, out=out, want_int16=True, seed=1,
                   flags=_lib.FLAG_ENVELOPE_COMPRESSOR if a.compressor == "envelope" else 0)
        eng.sync()
        wall = time.time() - t0
        kt = eng.kernel_times()
        eng.timing(False)
        tot = sum(v[0] for v in kt.values())
        print(f"rep {rep}: wall {wall*1e3:.1f} ms, kernels {tot:.1f} ms, audio-s/s {a.tracks*a.sec/wall:.0f}")
    frames = a.tracks * n
    for k, (ms, cnt, _smp) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:32s} {ms:9.3f} ms  x{cnt:<3d}  {ms/tot*100:5.1f}%   {frames*2*4/ (ms/cnt*1e-3)/1e9:8.1f} GB/s per 1R-stream-equivalent")
    print("workspace GB", eng.lib.mm_ctx_workspace_bytes(eng.ctx) / 1e9, "launches", eng.launch_count())


if __name__ == "__main__":
    main()
