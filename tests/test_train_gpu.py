"""GPU tests (-m gpu) of the optimiser / critic-head kernels and a short run of the re-hosted train loop."""
import numpy as np
import pytest
import torch

from oracle import nn_oracle as no

pytestmark = pytest.mark.gpu


def test_adam_ema_kernel_vs_oracle():
    from otgan_b200.utils import nn
    rng = np.random.RandomState(0)
    n = 4096 + 8
    p0 = rng.randn(n).astype(np.float32)
    flat = torch.from_numpy(p0.copy()).cuda().requires_grad_(True)
    ema = nn.ExponentialMovingAverage(0.999)
    ema.shadow = flat.detach().clone()
    opt = nn.adam_updates(flat, lr=3e-4, mom1=0.5, mom2=0.999, ema=ema)
    p, v, mg, sh = p0.astype(np.float64), np.zeros(n), np.zeros(n), p0.astype(np.float64)
    for t in range(1, 6):
        g = rng.randn(n).astype(np.float32)
        opt.run(torch.from_numpy(g).cuda(), lr=-3e-4 if t % 2 else 3e-4)
        p, v, mg = no.adam_step(p, g.astype(np.float64), v, mg, t, -3e-4 if t % 2 else 3e-4, 0.5, 0.999)
        sh = sh - (1 - 0.999) * (sh - p)
    np.testing.assert_allclose(flat.detach().cpu().numpy(), p, rtol=0, atol=2e-6)
    np.testing.assert_allclose(ema.shadow.cpu().numpy(), sh, rtol=0, atol=2e-6)
    np.testing.assert_allclose(opt.state["mg"].cpu().numpy(), mg, rtol=1e-4, atol=1e-9)    # fp32 state vs fp64 oracle


@pytest.mark.parametrize("C", [96, 6])
def test_crelu_l2norm_kernel_fwd_bwd(C):
    """C = 96: float4 path; C = 6: scalar path of otgan_crelu_l2norm_{fwd,bwd}_f32."""
    from otgan_b200.utils import nn
    torch.manual_seed(0)
    x = torch.randn(5, 4, 4, C, device="cuda", requires_grad=True)
    y = nn.crelu_l2norm(x)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad([y], [x], [gy])
    xd = x.detach().double().requires_grad_(True)
    z = torch.cat([torch.relu(xd), torch.relu(-xd)], 3).reshape(5, -1)
    yr = z / torch.sqrt(torch.sum(z * z, 1, keepdim=True))
    (gr,) = torch.autograd.grad([yr], [xd], [gy.double()])
    assert float((y.double() - yr).abs().max()) < 1e-6 and float((gx.double() - gr).abs().max() / gr.abs().max()) < 1e-5
    ref = no.head(x.detach().cpu().double().numpy())
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref, atol=1e-6)


@pytest.mark.parametrize("shape", [(5, 5, 6, 40), (3, 3, 64, 16), (100, 200), (5, 5, 256, 96), (1, 1, 7, 3)])
def test_weightnorm_kernels_vs_torch(shape):
    """otgan_weightnorm_{fwd,bwd}_f32 against the literal op sequence of utils/nn.py:176-180 (float64 torch ops)."""
    from otgan_b200.utils import nn
    torch.manual_seed(len(shape) * 7 + shape[-1])
    V = (torch.randn(shape, device="cuda") * 0.05).requires_grad_(True)
    g = (1.0 + 0.3 * torch.rand(shape[-1], device="cuda")).requires_grad_(True)
    wt = nn._WeightNorm.apply(V, g)
    C, K = shape[-1], V.numel() // shape[-1]
    assert wt.shape == (C, K)
    gw = torch.randn_like(wt)
    gV, gg = torch.autograd.grad([wt], [V, g], [gw])
    Vd, gd = V.detach().double().requires_grad_(True), g.detach().double().requires_grad_(True)
    Wd = nn.l2_normalize(Vd, list(range(Vd.dim() - 1))) * gd
    wt_ref = Wd.reshape(K, C).t()
    rV, rg = torch.autograd.grad([wt_ref], [Vd, gd], [gw.double()])
    assert float((wt.double() - wt_ref).abs().max() / wt_ref.abs().max()) < 2e-6
    assert float((gV.double() - rV).abs().max() / rV.abs().max()) < 2e-5
    assert float((gg.double() - rg).abs().max() / rg.abs().max()) < 2e-5


def test_fused_activation_kernels_vs_torch():
    """otgan_crelu_pad_* and otgan_glu_up_* against the literal op sequences (float64 torch ops), forward and backward."""
    from otgan_b200.utils import nn
    torch.manual_seed(3)
    x = torch.randn(3, 6, 6, 8, device="cuda", requires_grad=True)
    for pads in ((1, 1, 2, 2), (2, 2, 2, 2), (0, 0, 1, 1), (0, 0, 0, 0)):
        z = nn._CreluPad.apply(x, pads)
        gz = torch.randn_like(z)
        (gx,) = torch.autograd.grad([z], [x], [gz])
        xd = x.detach().double().requires_grad_(True)
        zr = torch.nn.functional.pad(torch.relu(torch.cat([xd, -xd], 3)).permute(0, 3, 1, 2),
                                     (pads[1], pads[3], pads[0], pads[2])).permute(0, 2, 3, 1)
        (gr,) = torch.autograd.grad([zr], [xd], [gz.double()])
        assert torch.equal(z.double(), zr) and float((gx.double() - gr).abs().max()) == 0.0
    y = torch.randn(2, 5, 5, 16, device="cuda", requires_grad=True)
    for up, fusion in ((False, True), (True, True), (True, False)):
        nn.UPSAMPLE_FUSION = fusion                  # False: the GLU kernel writes the upsampled tensor itself (otgan_glu_up, up = 2)
        try:
            o = nn.glu(y, upsample=up)
        finally:
            nn.UPSAMPLE_FUSION = True
        if isinstance(o, nn.Upsampled2x):            # un-materialised handle for the fused upsample + conv path
            assert up and fusion and o.low.shape[1] == y.shape[1]
            o = o.materialize()
        go = torch.randn_like(o)
        (gy,) = torch.autograd.grad([o], [y], [go])
        yd = y.detach().double().requires_grad_(True)
        a, l = torch.chunk(yd, 2, 3)
        orf = a * torch.sigmoid(l)
        if up:
            orf = orf.repeat_interleave(2, 1).repeat_interleave(2, 2)
        (gr,) = torch.autograd.grad([orf], [yd], [go.double()])
        assert float((o.double() - orf).abs().max()) < 1e-6 and float((gy.double() - gr).abs().max()) < 1e-5


@pytest.mark.parametrize("backend,batch,tol", [("cudnn", 2, 2e-5), ("tcgen05", 8, 2e-3)])
def test_dcgan_forward_on_gpu_matches_oracle(backend, batch, tol):
    """Critic forward against the numpy float64 restatement of models/dcgan.py + utils/nn.py.  The library rung in strict
    fp32 pins the layer wiring to 2e-5; the tcgen05 convolution kernels (TF32 operands, batch 8 so that every layer tiles)
    agree to the TF32 operand rounding, 2e-3 of the largest feature."""
    from otgan_b200.models import dcgan
    from otgan_b200.utils import nn
    dcgan.discriminator.reset(); dcgan.generator.reset()
    torch.manual_seed(0)
    x = torch.rand(batch, 32, 32, 3, device="cuda") * 2 - 1
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, nn.CONV_BACKEND)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    nn.CONV_BACKEND = backend
    try:
        dcgan.discriminator(x, init=True)
        f = dcgan.discriminator(x).detach().cpu().double().numpy()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, nn.CONV_BACKEND = old
    table = {n: p.detach().cpu().double().numpy() for n, p in dcgan.discriminator.named_parameters()}
    ref = no.dcgan_discriminator(x.cpu().double().numpy(), table)
    assert np.abs(f - ref).max() / np.abs(ref).max() < tol
    dcgan.discriminator.reset(); dcgan.generator.reset()


@pytest.mark.parametrize("extra", [[], ["--single_batch"], ["--train_disc_against_ema"]])
def test_train_loop_smoke(extra):
    """20 steps of the re-hosted loop on synthetic data: finite distance, entropy in (0, ln h], D/G alternation, EMA."""
    from otgan_b200 import train as T
    args = T.build_parser().parse_args(["--synthetic", "--nr_gpu", "2", "--batch_size", "32", "--nr_sinkhorn_iter", "50",
                                        "--nr_gen_per_disc", "2"] + extra)
    tr = T.Trainer(args, torch.device("cuda", 0))
    assert tr.num_features == 32768
    g0, d0 = tr.generator.flat.detach().clone(), tr.discriminator.flat.detach().clone()
    kinds = []
    for s in range(9):
        x = torch.rand(64, 32, 32, 3, device="cuda") * 2 - 1
        kind, stats = tr.step(x)
        kinds.append(kind)
        dist_v, ent = stats.tolist()
        assert np.isfinite(dist_v) and np.isfinite(ent)
        h = 64 if "--single_batch" in extra else 32
        assert 0.0 <= ent <= np.log(h) + 1e-4
    assert kinds == ["disc", "gen", "gen"] * 3
    assert not torch.equal(tr.generator.flat.detach(), g0) and not torch.equal(tr.discriminator.flat.detach(), d0)
    assert not torch.equal(tr.ema.shadow, g0[: tr.ema.shadow.numel()])
    assert float((tr.ema.shadow - tr.generator.flat.detach()).abs().max()) > 0


def test_multi_gpu_parity_two_ranks():
    """G-rank step == 1-rank step on identical data (tests/mgpu_parity.py under torchrun); needs >= 2 visible GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(__file__), "mgpu_parity.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", script],
                         capture_output=True, text=True, timeout=280)
    assert out.returncode == 0 and "MGPU PARITY OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_cuda_graph_replay_matches_eager_steps():
    """enable_cuda_graphs(): six replayed steps (1 critic + 5 generator) leave the same parameters, EMA shadow and
    [distance, entropy] as six eagerly launched steps from the same state (same images, same latents)."""
    from otgan_b200 import train as T
    args = T.build_parser().parse_args(["--synthetic", "--nr_gpu", "2", "--batch_size", "8", "--nr_sinkhorn_iter", "20"])
    g = torch.Generator(device="cuda").manual_seed(11)
    xs = [torch.rand(16, 32, 32, 3, device="cuda", generator=g) * 2 - 1 for _ in range(7)]
    us = [torch.rand(16, 100, device="cuda", generator=g) * 2 - 1 for _ in range(7)]
    out = {}
    for mode in ("eager", "graph"):
        tr = T.Trainer(args, torch.device("cuda", 0))
        if mode == "graph":
            tr.enable_cuda_graphs()
            assert tr.step_counter == 0 and tr.gen_optimizer.state["t"] == 1
        stats = []
        for x, u in zip(xs, us):
            kind, s = tr.step(x, u=u)
            stats.append(s.clone())
        torch.cuda.synchronize()
        out[mode] = (tr.generator.flat.detach().clone(), tr.discriminator.flat.detach().clone(), tr.ema.shadow.clone(),
                     torch.stack(stats))
        if mode == "graph":
            assert tr.replayed_launches > 7 * 20 and tr.step_counter == 7
    errs = {name: float((a - c).abs().max()) / float(c.abs().max())
            for name, a, c in zip(("generator", "critic", "ema", "stats"), out["eager"], out["graph"])}
    print(errs)
    # same kernels in the same order; the only run-to-run freedom is cuDNN's algorithm choice for the two 3-channel layers
    assert all(e <= 1e-3 for e in errs.values()), errs


@pytest.mark.parametrize("graphs", [False, True])
def test_stage_and_fetch_pipeline_matches_plain_steps(graphs):
    """Trainer.stage (upload of the NEXT step's pinned host images on a copy stream) + Trainer.fetch_async (read-back of
    [distance, entropy] one step behind the GPU) -- the loop of train.main -- give exactly the statistics and parameters of
    plain step(device_tensor) calls, eagerly and under CUDA graphs."""
    from otgan_b200 import train as T
    args = T.build_parser().parse_args(["--synthetic", "--nr_gpu", "2", "--batch_size", "8", "--nr_sinkhorn_iter", "20"])
    gcpu = torch.Generator().manual_seed(5)
    xs_host = [(torch.rand(16, 32, 32, 3, generator=gcpu) * 2 - 1).pin_memory() for _ in range(7)]
    us = [(torch.rand(16, 100, generator=gcpu) * 2 - 1).cuda() for _ in range(7)]
    out = {}
    for mode in ("plain", "pipelined"):
        tr = T.Trainer(args, torch.device("cuda", 0))
        if graphs:
            tr.enable_cuda_graphs()
        stats = []
        if mode == "plain":
            for x, u in zip(xs_host, us):
                stats.append(tr.step(x.cuda(), u=u)[1].tolist())
        else:
            pending, x = None, tr.stage(xs_host[0])
            for i, u in enumerate(us):
                _, s = tr.step(x, u=u)
                if i + 1 < len(us):
                    x = tr.stage(xs_host[i + 1])
                h = tr.fetch_async(s)
                if pending is not None:
                    stats.append(pending.result())
                pending = h
            stats.append(pending.result())
        torch.cuda.synchronize()
        out[mode] = (torch.tensor(stats), tr.generator.flat.detach().clone(), tr.discriminator.flat.detach().clone())
    for a, b in zip(out["plain"], out["pipelined"]):
        assert torch.equal(a, b)


def test_critic_weight_cache_matches_recomputed_weights():
    """Generator steps reuse the critic's W = g V/||V|| (and its IHWO copy) computed once after the critic update: the
    features / input gradient through the cached weights equal those through freshly normalised weights, and an
    optimiser update invalidates the cache."""
    from otgan_b200.models.dcgan import discriminator, generator
    from otgan_b200.utils import nn
    dev = torch.device("cuda", 0)
    discriminator.reset(); generator.reset()
    torch.manual_seed(7)
    with torch.no_grad():
        discriminator(torch.zeros(8, 32, 32, 3, device=dev), init=True, device=dev)
    st = discriminator.store
    x = (torch.rand(8, 32, 32, 3, device=dev) * 2 - 1).requires_grad_(True)
    gy = torch.randn(8, 32768, device=dev)
    assert st.cache_version != st.version
    with nn.frozen_params():
        f0 = discriminator(x)
    (g0,) = torch.autograd.grad([f0], [x], [gy])
    st.refresh_weight_cache()
    assert st.cache_version == st.version and st.cached_weight("discriminator/conv2d_1") is None      # grad mode: not used
    with nn.frozen_params():
        assert st.cached_weight("discriminator/conv2d_1") is not None
        f1 = discriminator(x)
    (g1,) = torch.autograd.grad([f1], [x], [gy])
    assert torch.equal(f0, f1) and torch.equal(g0, g1)
    opt = nn.adam_updates(discriminator, lr=1e-3, mom1=0.5, mom2=0.999)
    opt.run(torch.randn_like(discriminator.flat))
    with nn.frozen_params():
        assert st.cached_weight("discriminator/conv2d_1") is None                                     # stale after the update
        f2 = discriminator(x)
    assert not torch.equal(f1, f2)
    discriminator.reset(); generator.reset()


def test_densenet_train_steps_on_gpu():
    """--model densenet (BASELINE config 4's model family): one critic and one generator step of the re-hosted loop run
    and give a finite distance and an entropy in (0, ln h]; D = 7296.  (Its 16-filter convolutions run on the library
    rung; the matching, head, weight-norm and optimiser kernels are this library's.)"""
    from otgan_b200 import train as T
    args = T.build_parser().parse_args(["--synthetic", "--model", "densenet", "--nr_gpu", "2", "--batch_size", "8",
                                        "--nr_sinkhorn_iter", "20"])
    tr = T.Trainer(args, torch.device("cuda", 0))
    assert tr.num_features == 7296
    p0 = tr.discriminator.flat.detach().clone()
    for expect in ("disc", "gen"):
        kind, stats = tr.step(torch.rand(16, 32, 32, 3, device="cuda") * 2 - 1)
        d, e = stats.tolist()
        assert kind == expect and np.isfinite(d) and 0.0 < e <= np.log(8) + 1e-4
    assert not torch.equal(tr.discriminator.flat.detach(), p0)


def test_dcgan_64x64_extension_trains_on_gpu():
    """--image_size 64 (BASELINE config 5's image shape; an extension, the reference is hard-wired to 32 x 32): D = 131072
    features, one critic and two generator steps with finite distance / entropy.  The stride-2 critic layers and the three
    fused upsample + convolution layers run on the tcgen05 kernels at these extents; the two 3-channel layers (64 pixels wide)
    fall back to the library rung."""
    from otgan_b200 import train as T
    args = T.build_parser().parse_args(["--synthetic", "--image_size", "64", "--nr_gpu", "2", "--batch_size", "8",
                                        "--nr_sinkhorn_iter", "20"])
    tr = T.Trainer(args, torch.device("cuda", 0))
    assert tr.num_features == 131072
    for expect in ("disc", "gen", "gen"):
        kind, stats = tr.step(torch.rand(16, 64, 64, 3, device="cuda") * 2 - 1)
        d, e = stats.tolist()
        assert kind == expect and np.isfinite(d) and 0.0 < e <= np.log(8) + 1e-4


def test_loss_parity_tf32_convolutions_vs_strict_fp32():
    """North star "loss parity to the reference": the same training trajectory (fixed seeds, images, latents; N = 64, T = 100,
    lambda = 500, Adam 3e-4) on the tcgen05 convolution kernels (TF32 operands) and on the strict-fp32 library rung -- the
    precision class of the reference's TensorFlow-1.x convolutions.  OT-GAN training is chaotic: a CONTROL run of the fp32 rung
    with 1e-6 relative noise on its input images separates from the fp32 run just as fast (profiles/r02_loss_parity_dcgan.json:
    mean distance gap over steps 20-60 0.099 for TF32-vs-fp32 against 0.128 for fp32-vs-perturbed-fp32).  So the gate is (i) the
    first 20 steps, before chaos takes over, within a band measured at 3x margin (distance 0.03 = 5% of its scale, entropy 0.15),
    and (ii) the later TF32-vs-fp32 gap no larger than 3x the control's own gap (+0.05)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("loss_parity", os.path.join(os.path.dirname(__file__), "..", "tools", "loss_parity.py"))
    lp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lp)
    r = lp.compare(steps=60, n=64, t_iters=100, lam=500.0, model="dcgan")
    a, c = r["tf32_vs_fp32"], r["fp32_vs_fp32_perturbed_1e-6"]
    assert a["distance_gap_max_steps_0_20"] < 0.03 and a["entropy_gap_max_steps_0_20"] < 0.15, a
    assert a["distance_gap_mean_steps_20_60"] < 3.0 * c["distance_gap_mean_steps_20_60"] + 0.05, (a, c)
    assert a["entropy_gap_mean_steps_20_60"] < 3.0 * c["entropy_gap_mean_steps_20_60"] + 0.15, (a, c)
