"""Sinkhorn kernel time against the iteration count (fixed cost vs per-iteration slope) at the headline block size.

    python tools/sinkhorn_timing.py [out.json]

Launches otgan_sinkhorn_ex_f32 directly (pre-allocated outputs, 16 back-to-back launches between one CUDA-event pair) on the
six cost blocks of the bench's synthetic embeddings (N = 256, D = 32768, lambda = 500) for T in {1, 2, 10, 50, 100, 200, 500},
with and without the plan output, and reports the slow-path step counts."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otgan_b200 import _lib                      # noqa: E402
from otgan_b200.utils import matching as M       # noqa: E402
from oracle import matching_oracle as mo          # noqa: E402  (input generator only)


def main():
    dev = torch.device("cuda", 0)
    N, D, lam = 256, 32768, 500.0
    h = N // 2
    sets = []
    for s in range(4):
        a = torch.from_numpy(mo.synth_embeddings(N, D, 100 + 2 * s, "clustered", sigma=1.0)).to(dev)
        b = torch.from_numpy(mo.synth_embeddings(N, D, 101 + 2 * s, "clustered", sigma=1.0)).to(dev)
        sets.append(M.cost_blocks([a[:h], b[h:], a[:h], a[:h], a[h:], a[h:]], [a[h:], b[:h], b[:h], b[h:], b[:h], b[h:]], lam))
    if "--dump" in sys.argv:                     # raw L0 of the first input set for tools/sinkhorn_phases.cu
        sets[0].cpu().numpy().tofile(sys.argv[sys.argv.index("--dump") + 1])
        return
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    P = torch.empty((6, h, h), device=dev)
    ent = torch.empty((6,), device=dev)
    pc = torch.empty((6,), device=dev)
    slow = torch.zeros((6,), device=dev, dtype=torch.int32)
    out = {}
    for want_p in (True, False):
        for T in (1, 2, 10, 50, 100, 200, 500):
            def run(i):
                rc = lib.otgan_sinkhorn_ex_f32(6, h, h, T, lam, sets[i % 4].data_ptr(), P.data_ptr() if want_p else None, ent.data_ptr(),
                                               pc.data_ptr(), slow.data_ptr(), 0, stream.cuda_stream)
                assert rc == 0
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            REP = 16
            e0.record(stream)
            for i in range(REP):
                run(i)
            e1.record(stream)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / REP * 1e3
            out["T%d_%s" % (T, "P" if want_p else "noP")] = {"us": us, "slow_steps": slow.tolist(), "iters_per_sec": T / (us * 1e-6)}
    t1, t5 = out["T100_P"]["us"], out["T500_P"]["us"]
    out["slope_us_per_iter"] = (t5 - t1) / 400.0
    out["fixed_us"] = t1 - 100.0 * out["slope_us_per_iter"]
    txt = json.dumps(out, indent=1)
    print(txt)
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        open(sys.argv[1], "w").write(txt)


if __name__ == "__main__":
    main()
