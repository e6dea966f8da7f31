"""Sinkhorn on blocks larger than one SM (128 < side <= 512): the 8-CTA cluster kernel against the streaming rung.

    python tools/sinkhorn_large_timing.py [out.json]

For (nblk, h) in {(6, 256) = BASELINE config 5, (3, 256) = single-batch N = 256, (6, 512) = the weak-scaling config at 8 GPUs} and
T in {10, 100}: replays a CUDA graph of one otgan_sinkhorn_ex_f32 call (as the training step does) with impl = AUTO (cluster kernel)
and impl = SIMT (streaming kernels: 2T + 1 launches), reports microseconds per call and the largest plan difference between the two."""
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
    lib = _lib.load()
    lam, D = 500.0, 4096
    out = {}
    for nblk, h in ((6, 256), (3, 256), (6, 512)):
        xs = [torch.from_numpy(mo.synth_embeddings(h, D, 300 + k, "clustered", sigma=1.0)).to(dev) for k in range(nblk)]
        ys = [torch.from_numpy(mo.synth_embeddings(h, D, 400 + k, "clustered", sigma=1.0)).to(dev) for k in range(nblk)]
        L0 = M.cost_blocks(xs, ys, lam)
        P = [torch.empty((nblk, h, h), device=dev) for _ in range(2)]
        ent, pc = torch.empty((nblk,), device=dev), torch.empty((nblk,), device=dev)
        for T in (10, 100):
            row = {}
            for k, (name, impl) in enumerate((("cluster", _lib.IMPL_AUTO), ("stream", _lib.IMPL_SIMT))):
                side = torch.cuda.Stream()
                g = torch.cuda.CUDAGraph()

                def call(stream):
                    rc = lib.otgan_sinkhorn_ex_f32(nblk, h, h, T, lam, L0.data_ptr(), P[k].data_ptr(), ent.data_ptr(), pc.data_ptr(),
                                                   None, impl, stream.cuda_stream)
                    _lib.check(rc, "otgan_sinkhorn_ex_f32")
                call(torch.cuda.current_stream())
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=side):
                    call(torch.cuda.current_stream())
                for _ in range(3):
                    g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                REP = 10
                e0.record()
                for _ in range(REP):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                row[name + "_us"] = e0.elapsed_time(e1) / REP * 1e3
            row["max_abs_dP_over_max_P"] = float((P[0] - P[1]).abs().max() / P[1].abs().max())
            out["nblk%d_h%d_T%d" % (nblk, h, T)] = row
    txt = json.dumps(out, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt)


if __name__ == "__main__":
    main()
