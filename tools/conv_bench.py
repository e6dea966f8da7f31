"""Per-layer timing of the tcgen05 convolution kernels on the full-size DCGAN layer shapes (N = 256: critic on 512 images,
generator on 256), next to cuDNN's TF32 kernels through torch on the same tensors.  CUDA events, warm-up, inputs far
larger than L2.  Writes gpurun_out/conv_bench.json.

    python tools/conv_bench.py [--iters 5] [--layers c1,c2,...]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otgan_b200 import _lib  # noqa: E402

LAYERS = {  # name: (B, H, W, Cin, Cout, k, stride)
    "c1": (512, 32, 32, 256, 256, 5, 2),
    "c2": (512, 16, 16, 512, 512, 5, 2),
    "c3": (512, 8, 8, 1024, 1024, 5, 2),
    "g1": (256, 8, 8, 1024, 1024, 5, 1),
    "g2": (256, 16, 16, 512, 512, 5, 1),
    "g3": (256, 32, 32, 256, 256, 5, 1),
}


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--layers", type=str, default=",".join(LAYERS))
    ap.add_argument("--no-cudnn", action="store_true")
    ap.add_argument("--batch-div", type=int, default=1, help="divide the batch (8 = the per-rank batch of an 8-GPU run)")
    args = ap.parse_args()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    torch.backends.cudnn.allow_tf32 = True
    rows = []
    for name in args.layers.split(","):
        B, H, W, Cin, Cout, k, s = LAYERS[name]
        B = max(B // args.batch_div, 8)
        Ho, Wo = H // s, W // s
        flops = 2.0 * B * Ho * Wo * Cout * k * k * Cin
        x = torch.randn(B, H, W, Cin, device="cuda")
        w = torch.randn(Cout, k * k * Cin, device="cuda") * 0.02
        b = torch.randn(Cout, device="cuda")
        dy = torch.randn(B, Ho, Wo, Cout, device="cuda")
        y = torch.empty(B, Ho, Wo, Cout, device="cuda")
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        wt = torch.empty(Cin, k * k * Cout, device="cuda")
        pt, pb = same_pad(H, k, s)
        pl, pr = same_pad(W, k, s)
        need = lib.otgan_workspace_bytes_conv_wgrad(B, H, W, Cin, Cout, k, k, s)
        ws = torch.empty(need // 4 + 64, device="cuda")
        gneed = max(lib.otgan_workspace_bytes_conv_gemm(B, Ho, Wo, Cout), lib.otgan_workspace_bytes_conv_gemm(B, H, W, Cin))
        gws = torch.empty(gneed // 4 + 64, device="cuda")

        def fprop():
            _lib.check(lib.otgan_conv2d_fprop_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), gws.data_ptr(), gws.numel() * 4, st), "fprop")

        def dgrad():
            _lib.check(lib.otgan_ohwi_to_ihwo_f32(Cout, k * k, Cin, w.data_ptr(), wt.data_ptr(), st), "transpose")
            _lib.check(lib.otgan_conv2d_dgrad_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), gws.data_ptr(), gws.numel() * 4, st), "dgrad")

        def wgrad():
            _lib.check(lib.otgan_conv2d_wgrad_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, st), "wgrad")

        row = {"layer": name, "shape": (B,) + LAYERS[name][1:], "gflop": flops / 1e9}
        for op, fn in (("fprop", fprop), ("dgrad", dgrad), ("wgrad", wgrad)):
            ms = timeit(fn, args.iters)
            row[op + "_ms"] = ms
            row[op + "_tflops"] = flops / ms / 1e9
        if not args.no_cudnn:
            xn = x.permute(0, 3, 1, 2)                               # NCHW view of channels-last memory
            if pt != pb:
                xn = F.pad(xn, (pl, pr, pt, pb)).contiguous(memory_format=torch.channels_last)
                padding = (0, 0)
            else:
                padding = (pt, pl)
            wn = w.view(Cout, k, k, Cin).permute(0, 3, 1, 2)
            dyn = dy.permute(0, 3, 1, 2)
            ms = timeit(lambda: F.conv2d(xn, wn, b, stride=s, padding=padding), args.iters)
            row["cudnn_fprop_ms"], row["cudnn_fprop_tflops"] = ms, flops / ms / 1e9
            for op, mask in (("dgrad", [True, False, False]), ("wgrad", [False, True, False])):
                ms = timeit(lambda: torch.ops.aten.convolution_backward(dyn, xn, wn, None, [s, s], list(padding), [1, 1], False, [0, 0], 1, mask), args.iters)
                row["cudnn_%s_ms" % op], row["cudnn_%s_tflops" % op] = ms, flops / ms / 1e9
        print(json.dumps(row))
        sys.stdout.flush()
        rows.append(row)
        del x, w, dy, y, dx, dw, wt, ws, gws
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "conv_bench%s.json" % ("" if args.batch_div == 1 else "_div%d" % args.batch_div)), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
