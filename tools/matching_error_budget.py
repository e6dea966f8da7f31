"""Error budget of the CUDA matching path against the fp64 oracle, stage by stage (diagnostic; prints a table).

    python tools/matching_error_budget.py [golden.npz | N D G T sigma]
For each stage the GPU kernel is fed the ORACLE's input for that stage (rounded to fp32), so its own error is isolated:
  cost        GPU L(A, B)               vs -lam * C64
  sinkhorn    GPU P(L32 of the oracle)  vs P64 (the oracle run on the same fp32 L)
  plan-apply  GPU f(P32 of the oracle)  vs P32 @ F in fp64      (tcgen05 and SIMT implementations)
  end to end  GPU get_matched_features  vs the oracle
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import matching_oracle as mo  # noqa: E402
from otgan_b200 import _lib  # noqa: E402
from otgan_b200.utils import matching as M  # noqa: E402


def rel(a, ref):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    return float(np.abs(a - ref).max() / np.abs(ref).max())


def main():
    if len(sys.argv) == 2:
        g = np.load(sys.argv[1])
        A, B, lam, T, G = g["A"], g["B"], float(g["lam"]), int(g["T"]), int(g["G"])
    else:
        N, D, G, T = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 256, 8, 100)
        sigma = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
        lam = 500.0
        A, B = mo.synth_embeddings(N, D, 1, "clustered", sigma=sigma), mo.synth_embeddings(N, D, 2, "clustered", sigma=sigma)
    N, D = A.shape
    h = N // 2
    fa, fb = list(np.split(A, G)), list(np.split(B, G))
    res, plans, dists = mo.get_matched_features(fa, fb, lam, T, np.float64, True)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    Ad, Bd = dev(A), dev(B)
    a1, a2, b1, b2 = Ad[:h], Ad[h:], Bd[:h], Bd[h:]
    X, Y = [a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2]
    print("N=%d h=%d D=%d G=%d T=%d lam=%g" % (N, h, D, G, T, lam))
    for name, impl in (("tcgen05", _lib.IMPL_TCGEN05), ("simt", _lib.IMPL_SIMT)):
        try:
            L = M.cost_blocks(X, Y, lam, 0, None, impl)
            print("cost (%s): max |dC| / max C = %.3e" % (name, max(rel(L[k] / -lam, dists[k]) for k in range(6))))
        except Exception as e:
            print("cost (%s): %s" % (name, e))
    L32 = np.stack([(-lam * d).astype(np.float32) for d in dists])
    P64 = [mo.sinkhorn(L32[k].astype(np.float64) / -lam, lam, T, np.float64)[0] for k in range(6)]
    auto_name = "scaling-form, one CTA per block" if h <= 128 else "log-domain, one 8-CTA cluster per block" if h <= 512 else "log-domain, streaming"
    simt_name = "log-domain, one CTA per block" if h <= 128 else "log-domain, streaming"
    for name, impl in ((auto_name, 0), (simt_name, 1)):
        P, ent, pc = M.sinkhorn(dev(L32), lam, T, True, impl)
        print("sinkhorn (%s): max |dP| / max P = %.3e" % (name, max(rel(P[k], P64[k]) for k in range(6))))
    P32 = np.stack([p.astype(np.float32) for p in plans])
    Pd = dev(P32)
    ref = mo._combine_two_batch([p.astype(np.float64) for p in P32], A[:h].astype(np.float64), A[h:].astype(np.float64),
                                B[:h].astype(np.float64), B[h:].astype(np.float64))
    lib = _lib.load()
    ws, ws_bytes = M._plan_ws(Ad.device, h)
    for name, impl in (("tcgen05", 2), ("simt", 1)):
        outs = [torch.empty(N, D, device="cuda") for _ in range(4)]
        rc = lib.otgan_matched_two_batch_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, outs[0].data_ptr(), outs[1].data_ptr(),
                                             outs[2].data_ptr(), outs[3].data_ptr(), D, ws.data_ptr(), ws_bytes, impl,
                                             torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            print("plan-apply (%s): rc=%d" % (name, rc))
            continue
        print("plan-apply (%s): " % name + "  ".join("%.3e" % rel(o, r) for o, r in zip(outs, ref)))
    ta, tb = list(torch.chunk(Ad, G, 0)), list(torch.chunk(Bd, G, 0))
    for name, impl in (("auto", 0), ("simt", 1)):
        got = M.get_matched_features(ta, tb, lam, T, impl=impl)
        print("end to end (%s): " % name + "  ".join("%.3e" % rel(torch.cat(got[i]), np.concatenate(res[i])) for i in range(4)))
    r32 = mo.get_matched_features(fa, fb, lam, T, np.float32)
    print("fp32 numpy oracle: " + "  ".join("%.3e" % rel(np.concatenate(r32[i]), np.concatenate(res[i])) for i in range(4)))


if __name__ == "__main__":
    main()
