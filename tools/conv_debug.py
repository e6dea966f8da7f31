"""Diagnostic runner for the tcgen05 convolution kernels: every (op, shape) case runs in its OWN subprocess (a trapped or
hung kernel poisons the CUDA context) under a timeout, on integer inputs where the result must be exact, and reports the
structure of any mismatch (which pixels / channel blocks / taps are wrong).

    python tools/conv_debug.py                 # all cases -> stdout (and gpurun_out/conv_debug.txt)
    python tools/conv_debug.py --case fprop 0  # one case in this process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [
    (8, 8, 8, 128, 128, 5, 1),
    (8, 8, 8, 128, 128, 5, 2),
    (2, 32, 32, 128, 256, 5, 1),
    (2, 32, 32, 256, 256, 5, 2),
    (4, 16, 16, 256, 128, 5, 2),
    (8, 4, 4, 128, 128, 3, 1),
    (16, 8, 8, 1024, 1024, 5, 2),        # the DCGAN layers at batch 16: several N tiles / channel tiles per launch
    (16, 16, 16, 512, 512, 5, 2),
    (16, 32, 32, 256, 256, 5, 2),
    (16, 8, 8, 1024, 1024, 5, 1),
    (16, 16, 16, 512, 512, 5, 1),
    (16, 32, 32, 256, 256, 5, 1),
    (80, 16, 16, 128, 256, 5, 1),        # tail split: 160 tiles = one wave + 12
    (96, 32, 32, 128, 128, 5, 2),
]
OPS = ("fprop", "dgrad", "wgrad")


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def describe(name, got, ref, dims):
    import torch
    diff = (got.double() - ref).abs()
    bad = diff > 0
    nbad = int(bad.sum())
    print("  %s: %d / %d elements differ, max |diff| %g, max |ref| %g" % (name, nbad, diff.numel(), float(diff.max()), float(ref.abs().max())))
    if nbad == 0:
        return True
    # mismatch fraction along every axis
    for ax, label in enumerate(dims):
        other = [a for a in range(got.dim()) if a != ax]
        frac = bad.float().mean(dim=other)
        vals = ["%.2f" % v for v in frac.tolist()]
        if len(vals) > 64:                       # channel axes: summarise per block of 32
            blk = frac.view(-1, 32).mean(1)
            vals = ["%.2f" % v for v in blk.tolist()]
            label += " (per 32-block)"
        print("    bad fraction along %s: %s" % (label, " ".join(vals)))
    idx = bad.nonzero()[:6]
    for i in idx:
        t = tuple(i.tolist())
        print("    at %s got %g ref %g" % (t, float(got[t]), float(ref[t])))
    print("    got: any nonzero %s, nan %s" % (bool((got != 0).any()), bool(torch.isnan(got).any())))
    return False


def run_case(op, si):
    import torch
    import torch.nn.functional as F
    from otgan_b200 import _lib
    lib = _lib.load()
    B, H, W, Cin, Cout, k, s = SHAPES[si]
    g = torch.Generator(device="cuda").manual_seed(1 + si)
    x = torch.randint(-2, 3, (B, H, W, Cin), device="cuda", generator=g).float()
    w = torch.randint(-2, 3, (Cout, k, k, Cin), device="cuda", generator=g).float()
    b = torch.randint(-4, 5, (Cout,), device="cuda", generator=g).float()
    dy = torch.randint(-2, 3, (B, H // s, W // s, Cout), device="cuda", generator=g).float()
    pt, pb = same_pad(H, k, s)
    pl, pr = same_pad(W, k, s)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    y = F.conv2d(F.pad(xd.permute(0, 3, 1, 2), (pl, pr, pt, pb)), wd.permute(0, 3, 1, 2), bd, stride=s).permute(0, 2, 3, 1)
    dxr, dwr, dbr = torch.autograd.grad([y], [xd, wd, bd], [dy.double()])
    st = torch.cuda.current_stream().cuda_stream
    gneed = max(lib.otgan_workspace_bytes_conv_gemm(B, H // s, W // s, Cout), lib.otgan_workspace_bytes_conv_gemm(B, H, W, Cin))
    gws = torch.empty((gneed // 4 + 64,), device="cuda")
    print("case %s shape %s (pad %d/%d), split-K workspace %d bytes" % (op, SHAPES[si], pt, pb, gneed))
    ok = True
    if op == "fprop":
        out = torch.full((B, H // s, W // s, Cout), float("nan"), device="cuda")
        rc = lib.otgan_conv2d_fprop_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), gws.data_ptr(), gws.numel() * 4, st)
        _lib.check(rc, "fprop")
        torch.cuda.synchronize()
        ok = describe("y", out, y.detach(), ["n", "oh", "ow", "co"])
    elif op == "dgrad":
        wt = torch.empty((Cin, k * k * Cout), device="cuda")
        _lib.check(lib.otgan_ohwi_to_ihwo_f32(Cout, k * k, Cin, w.data_ptr(), wt.data_ptr(), st), "transpose")
        torch.cuda.synchronize()
        wt_ref = w.view(Cout, k * k, Cin).permute(2, 1, 0).reshape(Cin, -1)
        ok = describe("w_ihwo", wt, wt_ref.double(), ["ci", "tap*co"])
        out = torch.full((B, H, W, Cin), float("nan"), device="cuda")
        rc = lib.otgan_conv2d_dgrad_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, dy.data_ptr(), wt.data_ptr(), out.data_ptr(), gws.data_ptr(), gws.numel() * 4, st)
        _lib.check(rc, "dgrad")
        torch.cuda.synchronize()
        ok = describe("dx", out, dxr, ["n", "ih", "iw", "ci"]) and ok
    else:
        need = lib.otgan_workspace_bytes_conv_wgrad(B, H, W, Cin, Cout, k, k, s)
        ws = torch.empty((need // 4 + 64,), device="cuda")
        out = torch.full((Cout, k, k, Cin), float("nan"), device="cuda")
        rc = lib.otgan_conv2d_wgrad_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, dy.data_ptr(), x.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel() * 4, st)
        _lib.check(rc, "wgrad")
        torch.cuda.synchronize()
        ok = describe("dw", out, dwr, ["co", "kh", "kw", "ci"])
        P = B * (H // s) * (W // s)
        ws2 = torch.empty((lib.otgan_workspace_bytes_colsum(P, Cout) // 4 + 64,), device="cuda")
        db = torch.full((Cout,), float("nan"), device="cuda")
        _lib.check(lib.otgan_colsum_f32(P, Cout, dy.data_ptr(), db.data_ptr(), ws2.data_ptr(), ws2.numel() * 4, st), "colsum")
        torch.cuda.synchronize()
        ok = describe("db", db, dbr, ["co"]) and ok
    print("RESULT %s %d %s" % (op, si, "OK" if ok else "MISMATCH"))
    return ok


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--case":
        ok = run_case(sys.argv[2], int(sys.argv[3]))
        sys.exit(0 if ok else 1)
    if len(sys.argv) >= 3 and sys.argv[1] == "--op":          # all shapes of one op in this process
        ok = True
        for si in range(len(SHAPES)):
            try:
                ok = run_case(sys.argv[2], si) and ok
            except Exception as e:                            # a trapped kernel kills the context: report and stop
                print("RESULT %s %d EXCEPTION %r" % (sys.argv[2], si, e))
                ok = False
                if "CUDA" in repr(e) or "cuda" in repr(e):
                    break
            sys.stdout.flush()
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "conv_debug.txt"), "w")
    summary = []
    for op in OPS:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--op", op], capture_output=True, text=True, timeout=240)
            text = r.stdout + ("\n[stderr tail]\n" + r.stderr[-1500:] if r.returncode != 0 else "")
            status = "OK" if r.returncode == 0 else "FAIL(rc=%d)" % r.returncode
        except subprocess.TimeoutExpired as e:
            text = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            status = "TIMEOUT"
        summary.append("%s %s" % (op, status))
        for f in (sys.stdout, log):
            f.write(text + "\n")
            f.flush()
    for f in (sys.stdout, log):
        f.write("==== summary\n" + "\n".join(summary) + "\n")
    log.close()
    sys.exit(0 if all(s.endswith("OK") for s in summary) else 1)


if __name__ == "__main__":
    main()
