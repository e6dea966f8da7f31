"""One small launch of every tcgen05 / mbarrier kernel of libotgan.so, for `compute-sanitizer --tool racecheck` (and memcheck):

    compute-sanitizer --tool racecheck python tools/racecheck.py [--cluster-only]
`--cluster-only`: just the 8-CTA cluster Sinkhorn kernel (distributed-shared-memory pushes + mbarrier exchange).  Shapes are the smallest that still take the tensor-core paths (the sanitizer slows kernels ~100x)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otgan_b200 import _lib  # noqa: E402
from otgan_b200.utils import matching as M  # noqa: E402


def cluster_sinkhorn():
    L3 = (torch.rand(2, 256, 200, device="cuda") * -600.0).contiguous()    # sinkhorn_cluster_kernel: DSMEM pushes + mbarrier exchange
    M.sinkhorn(L3, 500.0, 6)
    M.sinkhorn((torch.rand(1, 300, 512, device="cuda") * -600.0).contiguous(), 500.0, 4)


def main():
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(0)
    if "--cluster-only" in sys.argv:
        cluster_sinkhorn()
        torch.cuda.synchronize()
        print("racecheck driver: cluster launches completed")
        return
    # cost_tc_kernel (forced) + cost_finalize, sinkhorn_fast_kernel, plan_apply_tc_kernel + plan_prep
    A = torch.nn.functional.normalize(torch.rand(64, 256, device="cuda"), dim=1)
    B = torch.nn.functional.normalize(torch.rand(64, 256, device="cuda"), dim=1)
    a1, a2, b1, b2 = A[:32], A[32:], B[:32], B[32:]
    L = M.cost_blocks([a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2], 500.0, 0, None, _lib.IMPL_TCGEN05)
    P, ent, pc = M.sinkhorn(L, 500.0, 10)
    L2 = (torch.rand(2, 128, 128, device="cuda") * -600.0).contiguous()    # full-size blocks with a wide cost range: slow steps of both kinds
    M.sinkhorn(L2, 500.0, 30)
    cluster_sinkhorn()
    ws, wsb = M._plan_ws(A.device, 32)
    Ga, Gb = torch.empty_like(A), torch.empty_like(B)
    _lib.check(lib.otgan_grad_features_f32(32, 256, P.data_ptr(), A.data_ptr(), B.data_ptr(), 256, Ga.data_ptr(), Gb.data_ptr(), 256,
                                           ws.data_ptr(), wsb, _lib.IMPL_TCGEN05, st), "grad_features")
    # conv_gemm_tc_kernel<256> / <128>, conv_wgrad_tc_kernel, generic <32> and <256>
    Bn, H, W, Ci, Co = 2, 8, 8, 128, 256
    x = torch.randn(Bn, H, W, Ci, device="cuda")
    w = torch.randn(Co, 25 * Ci, device="cuda") * 0.05
    y = torch.empty(Bn, H, W, Co, device="cuda")
    _lib.check(lib.otgan_conv2d_fprop_tf32(Bn, H, W, Ci, Co, 5, 5, 1, 2, 2, x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, 0, st), "fprop")
    zc = torch.empty(Bn, H, W, 2 * Co, device="cuda")                      # the fused CReLU epilogue of the plain fprop kernel
    _lib.check(lib.otgan_conv2d_fprop_crelu_tf32(Bn, H, W, Ci, Co, 5, 5, 1, 2, 2, x.data_ptr(), w.data_ptr(), None, zc.data_ptr(), st), "fprop_crelu")
    wt = torch.empty(Ci, 25 * Co, device="cuda")
    _lib.check(lib.otgan_ohwi_to_ihwo_f32(Co, 25, Ci, w.data_ptr(), wt.data_ptr(), st), "ihwo")
    dx = torch.empty_like(x)
    _lib.check(lib.otgan_conv2d_dgrad_tf32(Bn, H, W, Ci, Co, 5, 5, 1, 2, 2, y.data_ptr(), wt.data_ptr(), dx.data_ptr(), None, 0, st), "dgrad")
    wsz = lib.otgan_workspace_bytes_conv_wgrad(Bn, H, W, Ci, Co, 5, 5, 1)
    wsw = torch.empty(wsz // 4 + 64, device="cuda")
    dw = torch.empty_like(w)
    _lib.check(lib.otgan_conv2d_wgrad_tf32(Bn, H, W, Ci, Co, 5, 5, 1, 2, 2, y.data_ptr(), x.data_ptr(), dw.data_ptr(), wsw.data_ptr(), wsw.numel() * 4, st), "wgrad")
    xg = torch.randn(3, 8, 8, 40, device="cuda")
    wg = torch.randn(16, 9 * 40, device="cuda")
    z = torch.empty(3, 8, 8, 32, device="cuda")
    _lib.check(lib.otgan_conv2d_fprop_ex_tf32(3, 8, 8, 40, 40, 16, 32, 3, 3, 1, 1, 1, xg.data_ptr(), wg.data_ptr(), None, z.data_ptr(), 1, st), "fprop_ex")
    wg2 = torch.randn(144, 9 * 40, device="cuda")
    y2 = torch.empty(3, 4, 4, 144, device="cuda")
    _lib.check(lib.otgan_conv2d_fprop_ex_tf32(3, 8, 8, 40, 40, 144, 144, 3, 3, 2, 0, 0, xg.data_ptr(), wg2.data_ptr(), None, y2.data_ptr(), 0, st), "fprop_ex s2")
    torch.cuda.synchronize()
    print("racecheck driver: all launches completed")


if __name__ == "__main__":
    main()
