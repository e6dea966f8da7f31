#!/usr/bin/env python
"""bench.py -- OT-GAN training hot path on B200: images/sec + Sinkhorn-iters/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|matching]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (same at every N): DCGAN on synthetic CIFAR-10-shaped batches, N = 256 real + 256 generated images
per step, T = 100 Sinkhorn iterations, lambda = 500, the reference's 1 critic : 5 generator step schedule.  One "step" =
one `sess.run` of train.py:214-226: generator forward, critic forward on real+fake, the matching hot path (six cosine-cost
blocks -> Sinkhorn -> feature gradients + distance/entropy), backward, summed tower gradients, Adam (+EMA).
images/sec = N / step time.  At N GPUs the 256 images are split over the ranks (strong scaling): one all-gather of the
[2*256/N, D] embedding slab, replicated cost+Sinkhorn, each rank back-propagates its own rows, gradient all-reduce(sum).

`value` feeds the step from device-resident images; `e2e` goes through the public API the way otgan_b200.train.main drives
it: Trainer.stage (H2D of pinned HOST images on a copy stream) -> Trainer.step -> Trainer.fetch_async (D2H of [distance,
entropy]) -> host read of every step's result, software-pipelined (upload of step k+1 and read of step k-1 while step k
computes); `e2e.blocking_ms_per_step` is the same with sess.run semantics (upload, step, read, block -- every step).
`matching` is the matching hot path alone (this library's kernels only): kernel times and the 6-launch step are timed
through CUDA-graph replays (the per-call host work is as long as the kernels), `matching.kernels.sinkhorn.large_blocks` times
six 256 x 256 / 512 x 512 blocks on the cluster kernel beside the streaming rung; `roofline` is the step's dominant kernel
against MEASURED_PEAKS.json, with a cuBLAS TF32 GEMM measured in the same run as the honest ceiling of a TF32 kernel.
`configs` carries the other BASELINE.json configurations that fit the launch (cfg2 / cfg3 at 1 GPU, cfg4 = DenseNet at
4 GPUs, cfg5 = 64 x 64 / N = 512 at 8 GPUs) and, for N > 1, a weak-scaling line (128 images per rank); `mgpu_parity` is
the multi-rank parity check (otgan_b200.train.parity_check).
`cpu_baseline` / `--impl reference`: the reference algorithm on the host cores at the SAME N = 256 -- torch-CPU conv
stacks + oracle/torch_oracle.py (one torch-CPU op per TensorFlow op of utils/matching.py), the C+OpenMP oracle's matching
phase reported beside it (TensorFlow 1.x is not installable here: kind = "port").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TOTAL, T_ITERS, LAMBDA, D_FEAT = 256, 100, 500.0, 32768
N_INPUT_SETS = 4      # matching sub-benchmark: rotate 4 x 64 MiB embedding sets (> 126 MB L2)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def synth(N, D, seed):
    from oracle import matching_oracle as mo  # input generator only (seeded synthetic embeddings)
    return mo.synth_embeddings(N, D, seed, "clustered", sigma=1.0)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable or no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def workload_config(world, workload):
    cfg = {"workload": "OT-GAN DCGAN training step (generator + critic fwd/bwd, 6 cosine-cost blocks + %d Sinkhorn iters + "
                       "feature gradients + distance/entropy, summed tower gradients, Adam+EMA), CIFAR-10-shaped synthetic "
                       "images: N=%d real + %d generated per step, h=%d, D=%d, lambda=%g, schedule 1 critic : 5 generator"
                       % (T_ITERS, N_TOTAL, N_TOTAL, N_TOTAL // 2, D_FEAT, LAMBDA),
           "N": N_TOTAL, "h": N_TOTAL // 2, "D": D_FEAT, "T": T_ITERS, "lambda": LAMBDA, "ranks": world,
           "images_per_rank": N_TOTAL // world, "towers": 2 * world if world > 1 else 2,
           "conv_backend": "libotgan tcgen05 implicit-GEMM kernels (fprop / dgrad / wgrad, TF32 operands, fp32 accumulation) for every "
                           "convolution of the step; the generator's resize_nearest_neighbor + 5x5 convolution pairs run as the fused "
                           "sub-pixel form (4 parity classes x 3x3 pre-summed sub-filters on the low-resolution input: same algebra, "
                           "9 instead of 25 taps); the 100-wide dense layer runs on them as a 1x1 convolution",
           "precision": "fp32 storage everywhere; matching kernels are fp32-exact (3xTF32 operands + fp32 register accumulation); "
                        "convolutions read the fp32 tensors as TF32 on the tensor cores (the math class of cuDNN's default fp32 convolution)",
           "l2_policy": "per-step working set (activations + 72 M parameters + Adam state, > 1 GB) exceeds the 126 MB L2; the "
                        "matching sub-benchmark rotates %d input sets (%.0f MiB)" % (N_INPUT_SETS, N_INPUT_SETS * 2 * N_TOTAL * D_FEAT * 4 / 2**20)}
    if workload == "matching":
        cfg["workload"] = "matching hot path only (see `matching`): N=%d, h=%d, D=%d, T=%d" % (N_TOTAL, N_TOTAL // 2, D_FEAT, T_ITERS)
    return cfg


# ----------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_matching_time(N, D, T, lam, reps, warmup):
    """Matching phase on the host cores (C+OpenMP restatement, all threads): (times, best phases, cores)."""
    from oracle import c_oracle as co
    A, B = synth(N, D, 1), synth(N, D, 2)
    cores = co.max_threads()
    times, phases = [], None
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        r = co.two_batch(A, B, lam, T, want_plans=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            if phases is None or dt <= min(times):
                phases = r["phase_ms"]
    return times, phases, cores


class CpuTrainer:
    """The reference training step on the CPU at the contract size: torch-CPU conv stacks (the re-hosted layer library on CPU
    tensors = the literal torch op sequence of utils/nn.py), oracle/torch_oracle.py matching (utils/matching.py restated with
    one torch-CPU op per TensorFlow op), Adam as in utils/nn.py:50-73.  n_total real + n_total generated images per step."""

    def __init__(self, n_total, T, lam):
        import torch
        from otgan_b200.models import dcgan
        self.torch, self.n, self.T, self.lam = torch, n_total, T, lam
        torch.manual_seed(1)
        dcgan.generator.reset(); dcgan.discriminator.reset()
        self.gen, self.disc = dcgan.generator, dcgan.discriminator
        with torch.no_grad():
            self.disc(torch.zeros(2, 32, 32, 3) + 0.1, init=True, device="cpu")
            self.gen(2, init=True, device="cpu")
        self.state = {t.name: {"t": 1, "v": torch.zeros_like(t.flat), "mg": torch.zeros_like(t.flat)} for t in (self.gen, self.disc)}
        self.step_counter = 0
        self.phase_s = {"matching": [], "cost": [], "sinkhorn": [], "matched": []}

    def _adam(self, tpl, grad, lr, mom1=0.5, mom2=0.999):
        torch, st = self.torch, self.state[tpl.name]
        with torch.no_grad():
            st["v"].mul_(mom1).add_(grad, alpha=1 - mom1)
            st["mg"].mul_(mom2).addcmul_(grad, grad, value=1 - mom2)
            v_hat = st["v"] / (1 - mom1 ** st["t"])
            mg_hat = st["mg"] / (1 - mom2 ** st["t"])
            tpl.flat.sub_(lr * v_hat / torch.sqrt(mg_hat + 1e-8))
            st["t"] += 1

    def step(self, x_real):
        torch = self.torch
        from oracle import torch_oracle as to
        n = self.n
        train_disc = self.step_counter % 6 == 0
        if train_disc:
            with torch.no_grad():
                x_gen = self.gen(n)
            feats = self.disc(torch.cat([x_gen, x_real], 0))
            f_gen, f_dat = feats[:n], feats[n:]
        else:
            x_gen = self.gen(n)
            with torch.no_grad():
                f_dat = self.disc(x_real)
            f_gen = self.disc(x_gen)
        t0 = time.perf_counter()
        ph = []
        with torch.no_grad():                                  # two towers, like the repo arm at one rank
            fa, fb = list(torch.chunk(f_gen.detach(), 2, 0)), list(torch.chunk(f_dat.detach(), 2, 0))
            m = to.get_matched_features(fa, fb, self.lam, self.T, phases=ph)
            d = to.calc_distance(fa, fb, m)
            ga = torch.cat([x - y for x, y in zip(m[0], m[2])], 0)                               # train.py:111
            gb = torch.cat([x - y for x, y in zip(m[1], m[3])], 0)                               # train.py:126
        self.phase_s["matching"].append(time.perf_counter() - t0)
        for k, v in zip(("cost", "sinkhorn", "matched"), ph):
            self.phase_s[k].append(v)
        if train_disc:
            (g,) = torch.autograd.grad([feats], [self.disc.flat], grad_outputs=[torch.cat([ga, gb], 0)])
            self._adam(self.disc, g, -3e-4)
        else:
            (g,) = torch.autograd.grad([f_gen], [self.gen.flat], grad_outputs=[ga])
            self._adam(self.gen, g, 3e-4)
        self.step_counter += 1
        return float(d), float(m[4])


def cpu_train_time(n_total, steps, warmup, budget_s=None):
    """Times `steps` CPU training steps at n_total images (after `warmup`).  With budget_s the counts are cut (whole 1:5
    cycles where possible) so that the run ends within the budget; returns (times, threads, trainer, warmup_done)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    tr = CpuTrainer(n_total, T_ITERS, LAMBDA)
    x = torch.rand(n_total, 32, 32, 3) * 2 - 1
    times, done_warm = [], 0
    t_first = None
    i = 0
    while True:
        if done_warm >= warmup and len(times) >= steps:
            break
        t0 = time.perf_counter()
        tr.step(x)
        dt = time.perf_counter() - t0
        if t_first is None:
            t_first = dt
            if budget_s is not None:                           # bound the run: keep the warm-up short and whole cycles of timed steps
                warmup = min(warmup, max(1, int(0.2 * budget_s / dt))) if warmup > 0 else 0
                fit = int((budget_s - warmup * dt) / dt)
                if fit < steps:
                    steps = max(min(steps, 6), (fit // 6) * 6) if fit >= 6 else max(1, min(steps, fit))
        if done_warm < warmup:
            done_warm += 1
        else:
            times.append(dt)
        i += 1
    return times, torch.get_num_threads(), tr, done_warm


def cpu_info():
    model = "?"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"cpu_count": os.cpu_count(), "cpu_model": model}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    mt, ph, mcores = cpu_matching_time(N_TOTAL, D_FEAT, T_ITERS, LAMBDA, 3, 1)      # C+OpenMP oracle, before torch spins up its own pool
    times, cores, tr, warm = cpu_train_time(N_TOTAL, args.steps, args.warmup, budget_s=270.0)
    mean = float(np.mean(times))
    val = N_TOTAL / mean
    skip = warm                                                                        # phase times of the timed steps only
    line = {
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, args.workload),
        "sinkhorn_iters_per_sec": T_ITERS / float(np.mean(tr.phase_s["sinkhorn"][skip:])),
        "cpu_baseline": {"value": val, "unit": "images/sec", "cores": cores, "kind": "port",
                         "sample": "%d training steps (schedule 1 critic : 5 generator, starting with the critic step) of the SAME "
                                   "DCGAN step at the contract size N=%d real + %d generated images, after %d warm-up steps; torch-CPU "
                                   "conv stacks + oracle/torch_oracle.py matching (one torch-CPU op per TensorFlow op of "
                                   "utils/matching.py); TensorFlow 1.x not installable offline; counts cut to fit ~270 s when the "
                                   "host is slow (steps_requested / warmup_requested keep what was asked)"
                                   % (len(times), N_TOTAL, N_TOTAL, warm),
                         "host": cpu_info(), "torch": torch.__version__,
                         "matching_phase_torch_cpu": {"ms": float(np.mean(tr.phase_s["matching"][skip:])) * 1e3,
                                                      "phase_ms": {k: float(np.mean(tr.phase_s[k][skip:])) * 1e3 for k in ("cost", "sinkhorn", "matched")}},
                         "matching_phase_c_oracle": {"ms": float(np.min(mt)) * 1e3, "images_per_sec": N_TOTAL / float(np.min(mt)),
                                                     "cores": mcores,
                                                     "phase_ms": {"cost": ph[0], "sinkhorn": ph[1], "matched_distance": ph[2]}}},
        "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
def matching_benchmark(torch, devv, steps, warmup):
    """Matching hot path alone (device-resident rotating inputs): step time, per-kernel times, launches."""
    from otgan_b200 import _lib
    from otgan_b200.utils import matching as M
    N, D, T, lam, h = N_TOTAL, D_FEAT, T_ITERS, LAMBDA, N_TOTAL // 2
    dev_sets = [(torch.from_numpy(synth(N, D, 100 + 2 * s)).to(devv), torch.from_numpy(synth(N, D, 101 + 2 * s)).to(devv))
                for s in range(N_INPUT_SETS)]
    stream = torch.cuda.current_stream()

    def step(A, B):
        return M.matching_step(list(torch.chunk(A, 2, 0)), list(torch.chunk(B, 2, 0)), lam, T)

    for i in range(max(warmup, 3)):
        step(*dev_sets[i % N_INPUT_SETS])
    torch.cuda.synchronize()
    lib = _lib.load()
    # The cost entry includes its split-K finalize launch, the grad entry its plan-preparation launch -- they are part of the
    # kernel's work.
    REP = 16
    Ls = [M.cost_blocks([a[:h], b[h:], a[:h], a[:h], a[h:], a[h:]], [a[h:], b[:h], b[:h], b[h:], b[:h], b[h:]], lam) for a, b in dev_sets]
    Ps = [M.sinkhorn(L, lam, T)[0] for L in Ls]
    Ga, Gb = torch.empty_like(dev_sets[0][0]), torch.empty_like(dev_sets[0][1])
    ws, ws_bytes = M._plan_ws(devv)

    def k_cost(i, impl=_lib.IMPL_AUTO):
        a, b = dev_sets[i % N_INPUT_SETS]
        M.cost_blocks([a[:h], b[h:], a[:h], a[:h], a[h:], a[h:]], [a[h:], b[:h], b[:h], b[h:], b[:h], b[h:]], lam, impl=impl)

    def k_sinkhorn(i):
        M.sinkhorn(Ls[i % N_INPUT_SETS], lam, T)

    def k_grad(i, impl=_lib.IMPL_AUTO):
        a, b = dev_sets[i % N_INPUT_SETS]
        rc = lib.otgan_grad_features_f32(h, D, Ps[i % N_INPUT_SETS].data_ptr(), a.data_ptr(), b.data_ptr(), D, Ga.data_ptr(),
                                         Gb.data_ptr(), D, ws.data_ptr(), ws_bytes, impl, torch.cuda.current_stream().cuda_stream)
        assert rc == 0

    phases = {}
    # Kernel-only times: the REP launches are captured into ONE CUDA graph and the graph is replayed between a CUDA-event pair, so no
    # host work (argument checks, tensor-map encoding, ~60 us per call -- as long as the kernels themselves) sits inside the timed
    # region.
    side = torch.cuda.Stream()
    for name, fn in (("cost", k_cost), ("sinkhorn", k_sinkhorn), ("grad", k_grad)):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(REP):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        phases[name] = [e0.elapsed_time(e1) / (3 * REP)]
        del g
    kernel_ms = {k: float(np.mean(v)) for k, v in phases.items()}
    # the whole matching step (cost -> Sinkhorn -> feature gradients -> distance / entropy: 6 launches), N_INPUT_SETS calls on
    # rotating inputs captured into one CUDA graph -- the form the train loop runs it in (its steps are graph replays)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(*dev_sets[0])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [step(*dev_sets[i]) for i in range(N_INPUT_SETS)]
    launches = _lib.launch_count() / N_INPUT_SETS * steps
    g.replay()
    torch.cuda.synchronize()
    reps = max(1, steps // N_INPUT_SETS)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * N_INPUT_SETS)
    del g, outs
    pk = peaks()
    alg_bytes = {"cost": 4.0 * 2 * N * D + 24.0 * h * h, "sinkhorn": 48.0 * h * h, "grad": 16.0 * N * D}
    kernels = {}
    for k in kernel_ms:
        gbs = alg_bytes[k] / (kernel_ms[k] * 1e-3) / 1e9
        kernels[k] = {"ms": kernel_ms[k], "alg_bytes": alg_bytes[k], "achieved_gbs": gbs, "hbm_frac": gbs / pk["hbm_gbs"]}
    kernels["cost"]["alg_flops"] = 12.0 * h * h * D
    kernels["grad"]["alg_flops"] = 24.0 * h * h * D
    kernels["sinkhorn"]["exp_per_sec_log_domain_equiv"] = 12.0 * h * h * T / (kernel_ms["sinkhorn"] * 1e-3)
    kernels["sinkhorn"]["streaming_form_bytes"] = 96.0 * h * h * T
    # Roofline object: the dominant GEMM-shaped kernel (cost or grad).  Under 3xTF32 both are tensor-bound (SURVEY 8d:
    # t_tensor = 3 * alg_flops / (bf16_peak / 2) exceeds t_hbm), so `achieved` is ALGORITHMIC TFLOP/s (the 3x split is not
    # counted) against the measured dense bf16 peak; the HBM view and the 3xTF32 ceiling are given beside it.  The Sinkhorn
    # kernel keeps its block on chip (HBM bytes = 0.8 MB) and is latency-bound by construction: see kernels.sinkhorn.
    dom = "grad" if kernel_ms["grad"] >= kernel_ms["cost"] else "cost"
    for k in ("cost", "grad"):
        tfs = kernels[k]["alg_flops"] / (kernel_ms[k] * 1e-3) / 1e12
        kernels[k]["achieved_tflops"] = tfs
        kernels[k]["tensor_frac_of_bf16_peak"] = tfs / pk["bf16_tflops"]
        kernels[k]["frac_of_3xtf32_ceiling"] = 3.0 * tfs / (pk["bf16_tflops"] / 2.0)
    kernels["sinkhorn"]["bound"] = "latency/SFU (block resident in registers+smem for all T iterations; 6 of 148 SMs)"
    # the three fractions of the north star's "cost + Sinkhorn" target, each against its own roofline time
    t_cost_roof = max(alg_bytes["cost"] / (pk["hbm_gbs"] * 1e9), 3.0 * kernels["cost"]["alg_flops"] / (pk["bf16_tflops"] / 2.0 * 1e12))
    kernels["cost"]["roofline_time_us"] = t_cost_roof * 1e6
    kernels["cost"]["frac_of_roofline_time"] = t_cost_roof / (kernel_ms["cost"] * 1e-3)
    t_grad_roof = max(alg_bytes["grad"] / (pk["hbm_gbs"] * 1e9), 3.0 * kernels["grad"]["alg_flops"] / (pk["bf16_tflops"] / 2.0 * 1e12))
    kernels["grad"]["roofline_time_us"] = t_grad_roof * 1e6
    kernels["grad"]["frac_of_roofline_time"] = t_grad_roof / (kernel_ms["grad"] * 1e-3)
    # dram__bytes_read+write per launch: NOT measured live -- constants copied from the committed ncu --set full capture
    # (profiles/r01_e_matching_ncu_full.txt); `traffic_source` says so
    ncu_traffic = {"cost": 67.34e6 + 3.77e6, "grad": 94.43e6 + 34.05e6}
    roof = {"bound": "tensor", "kernel": dom + (" (plan_apply_tc_kernel)" if dom == "grad" else " (cost_tc_kernel)"),
            "achieved": kernels[dom]["achieved_tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": kernels[dom]["tensor_frac_of_bf16_peak"], "traffic": ncu_traffic[dom],
            "frac_of_3xtf32_ceiling": kernels[dom]["frac_of_3xtf32_ceiling"], "hbm_frac": kernels[dom]["hbm_frac"],
            "peak_source": pk["source"], "traffic_source": "constant from profiles/r01_e_matching_ncu_full.txt (ncu --set full), not live",
            "timing": "kernel-only: 16 launches on rotating inputs captured into one CUDA graph, replayed between a CUDA-event pair",
            "note": "fp32-exact 3xTF32: three tcgen05 passes per algorithmic flop, operands at TF32 rate (= bf16/2)"}
    kernels["cost"]["kernel"] = "cost_tc_kernel + cost_finalize_kernel (3xTF32)"
    kernels["grad"]["kernel"] = "plan_prep_kernel + plan_apply_tc_kernel (3xTF32)"
    # Blocks larger than one SM (cfg5: h = 256; weak scaling at 8 GPUs: h = 512): the 8-CTA cluster kernel (AUTO) beside the
    # one-launch-per-half-step rung (SIMT), each as ONE graph replay of 4 calls on synthetic cost blocks.
    large = {}
    for hl in (256, 512):
        Ll = (torch.rand((6, hl, hl), device=devv) * -600.0).contiguous()
        for name, impl in (("cluster_us", _lib.IMPL_AUTO), ("stream_us", _lib.IMPL_SIMT)):
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                M.sinkhorn(Ll, lam, T, True, impl)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(4):
                    M.sinkhorn(Ll, lam, T, True, impl)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            large.setdefault("6x%dx%d_T%d" % (hl, hl, T), {})[name] = e0.elapsed_time(e1) / 4 * 1e3
            del g
    kernels["sinkhorn"]["large_blocks"] = large
    res = {"ms_per_step": ms, "images_per_sec": N / (ms * 1e-3), "sinkhorn_iters_per_sec": T / (kernel_ms["sinkhorn"] * 1e-3),
           "kernels": kernels, "gpu_launches_per_step": launches / steps}
    del dev_sets
    torch.cuda.empty_cache()
    return res, roof


CONV_LAYERS = {  # DCGAN layers on the tcgen05 kernels at the N=256 step.  Critic: (B, H, W, Cin, Cout, k, stride).
    "critic conv2d_1": (2 * N_TOTAL, 32, 32, 256, 256, 5, 2), "critic conv2d_2": (2 * N_TOTAL, 16, 16, 512, 512, 5, 2),
    "critic conv2d_3": (2 * N_TOTAL, 8, 8, 1024, 1024, 5, 2),
    # generator: resize_nearest_neighbor(2x) -> conv 5x5, run as the fused sub-pixel kernels on the LOW-resolution input
    # (B, Hlow, Wlow, Cin, Cout, k, "up2"): 4 parity classes x 9 pre-summed taps instead of 25 taps at the high resolution
    "generator conv2d_0": (N_TOTAL, 4, 4, 1024, 1024, 5, "up2"), "generator conv2d_1": (N_TOTAL, 8, 8, 512, 512, 5, "up2"),
    "generator conv2d_2": (N_TOTAL, 16, 16, 256, 256, 5, "up2")}
# The roofline object quotes the launch that was captured with ncu --set full: generator conv2d_0 fprop (fused upsample form,
# conv_gemm_tc_kernel<256>; profiles/r01_l_conv_up2_fprop_ncu_full.txt): dram__bytes_read.sum + dram__bytes_write.sum =
# 182.5 + 44.6 MB per launch, against 235 MB algorithmic (x_low 16.8 MB + the four sub-filters 151 MB + y 67 MB).
CONV_ROOFLINE_LAUNCH, CONV_ROOFLINE_TRAFFIC = "generator conv2d_0", 182.51e6 + 44.59e6


def conv_benchmark(torch, iters=5):
    """Live per-launch times (CUDA events on the launching stream, 2 warm-up + `iters` launches, every layer's tensors are
    far larger than L2 in total) of this library's convolution kernels on every DCGAN layer of the N=256 step.
    `gflop` = FLOPs the launch EXECUTES (critic: 2*B*Ho*Wo*Cout*25*Cin; fused generator layers:
    2*B*Ho*Wo*Cout*9*Cin); `gflop_reference_form` = what the reference's resize + 25-tap convolution would spend."""
    from otgan_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, (B, H, W, Cin, Cout, k, s) in CONV_LAYERS.items():
        up2 = s == "up2"
        Ho, Wo = (2 * H, 2 * W) if up2 else (H // s, W // s)
        pad = (k - 1) // 2 if up2 else (max((Ho - 1) * s + k - H, 0)) // 2
        ref_flops = 2.0 * B * Ho * Wo * Cout * k * k * Cin
        x = torch.randn(B, H, W, Cin, device="cuda")
        w = torch.randn(Cout, k * k * Cin, device="cuda") * 0.02
        b = torch.randn(Cout, device="cuda")
        dy = torch.randn(B, Ho, Wo, Cout, device="cuda")
        y, dx, dw = torch.empty_like(dy), torch.empty_like(x), torch.empty_like(w)
        gws = torch.empty(max(lib.otgan_workspace_bytes_conv_gemm(B, Ho, Wo, Cout), lib.otgan_workspace_bytes_conv_gemm(B, H, W, Cin)) // 4 + 64, device="cuda")
        gwa = (gws.data_ptr(), gws.numel() * 4)          # split-K / tail-split workspace, as nn._ConvTC passes it
        if up2:
            n1 = lib.otgan_up2_subtaps(k, pad)
            slots = n1 * n1
            flops = 2.0 * B * Ho * Wo * Cout * slots * Cin
            w_sub = torch.empty(4, Cout, slots * Cin, device="cuda")
            w_sub_t = torch.empty(4, Cin, slots * Cout, device="cuda")
            dw_sub = torch.empty_like(w_sub)
            _lib.check(lib.otgan_up2_weight_presum_f32(Cout, k, k, Cin, pad, pad, w.data_ptr(), w_sub.data_ptr(), st), "presum")
            for c in range(4):
                _lib.check(lib.otgan_ohwi_to_ihwo_f32(Cout, slots, Cin, w_sub[c].data_ptr(), w_sub_t[c].data_ptr(), st), "ohwi_to_ihwo")
            ws = torch.empty(lib.otgan_workspace_bytes_conv_up2_wgrad(B, H, W, Cin, Cout, k, k, pad, pad) // 4 + 64, device="cuda")
            ops = {
                "fprop": lambda: lib.otgan_conv2d_up2_fprop_tf32(B, H, W, Cin, Cout, k, k, pad, pad, x.data_ptr(), w_sub.data_ptr(), b.data_ptr(), y.data_ptr(), gwa[0], gwa[1], st),
                "dgrad": lambda: lib.otgan_conv2d_up2_dgrad_tf32(B, H, W, Cin, Cout, k, k, pad, pad, dy.data_ptr(), w_sub_t.data_ptr(), dx.data_ptr(), gwa[0], gwa[1], st),
                "wgrad": lambda: lib.otgan_conv2d_up2_wgrad_tf32(B, H, W, Cin, Cout, k, k, pad, pad, dy.data_ptr(), x.data_ptr(), dw_sub.data_ptr(), ws.data_ptr(), ws.numel() * 4, st),
            }
        else:
            flops = ref_flops
            wt = torch.empty(Cin, k * k * Cout, device="cuda")
            ws = torch.empty(lib.otgan_workspace_bytes_conv_wgrad(B, H, W, Cin, Cout, k, k, s) // 4 + 64, device="cuda")
            _lib.check(lib.otgan_ohwi_to_ihwo_f32(Cout, k * k, Cin, w.data_ptr(), wt.data_ptr(), st), "ohwi_to_ihwo")
            ops = {
                "fprop": lambda: lib.otgan_conv2d_fprop_tf32(B, H, W, Cin, Cout, k, k, s, pad, pad, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), gwa[0], gwa[1], st),
                "dgrad": lambda: lib.otgan_conv2d_dgrad_tf32(B, H, W, Cin, Cout, k, k, s, pad, pad, dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), gwa[0], gwa[1], st),
                "wgrad": lambda: lib.otgan_conv2d_wgrad_tf32(B, H, W, Cin, Cout, k, k, s, pad, pad, dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, st),
            }
        row = {"shape": [B, H, W, Cin, Cout, k, s], "gflop": flops / 1e9, "gflop_reference_form": ref_flops / 1e9}
        for op, fn in ops.items():
            for _ in range(2):
                _lib.check(fn(), op)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            row[op] = {"ms": ms, "tflops": flops / ms / 1e9}
        out[name] = row
        del x, w, b, dy, y, dx, dw, ws, gws, ops
        torch.cuda.empty_cache()
    return out


def measure_tf32_peak(torch, n=8192, reps=8):
    """cuBLAS TF32 GEMM (fp32 tensors, allow_tf32) n^3, best of `reps`, CUDA events: the measured ceiling of a TF32 kernel on
    this box, taken in the same run (MEASURED_PEAKS.json has no TF32 entry)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        c = torch.empty(n, n, device="cuda")
        for _ in range(2):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


DENSE_BLOCKS = {  # DenseNet dense blocks (models/densenet.py): (H, W, base element channels); L = 16 layers of 16 filters each
    "critic block 1": (32, 32, [32]), "critic block 2": (16, 16, [144]), "critic block 3": (8, 8, [200]),
    "generator block 1": (8, 8, [16, 16]), "generator block 2": (16, 16, [144, 16]), "generator block 3": (32, 32, [208, 16])}


def densenet_benchmark(torch, batch_critic, batch_gen, iters=3):
    """Live times of the dense-block kernels (csrc/dense_block.cu) at the cfg4 shapes: forward (16 contribution-form launches) and
    backward (15 gather dgrads + base dgrads + ONE wgrad GEMM + bias column sums) per block.  gflop = algorithmic FLOPs of the 16
    3x3 convolutions (2 * pixels * 9 * 16 * sum_r cin_r) -- the same for forward, for the input gradients and for the filter
    gradients; the backward executes more (the batched wgrad also fills the causally-masked half of dW_all)."""
    import ctypes
    from otgan_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, (H, W, base) in DENSE_BLOCKS.items():
        B = batch_critic if name.startswith("critic") else batch_gen
        L, G = 16, 16
        geom = _lib.DenseGeom()
        geom.B, geom.H, geom.W, geom.n_base, geom.L, geom.growth = B, H, W, len(base), L, G
        for i, c in enumerate(base):
            geom.base_ch[i] = c
        c0 = sum(base)
        ctot = lib.otgan_dense_channels(ctypes.byref(geom))
        Z = torch.rand(B, H, W, ctot, device="cuda")
        S = torch.empty(B, H, W, G * L, device="cuda")
        wf = torch.randn(G * L, 9, ctot, device="cuda") * 0.02
        bias = torch.randn(G * L, device="cuda") * 0.1
        dZ = torch.randn(B, H, W, ctot, device="cuda")
        WB = torch.empty(lib.otgan_dense_wb_floats(ctypes.byref(geom)), device="cuda")
        dY = torch.empty(B, H, W, G * L, device="cuda")
        dbase = [torch.empty(B, H, W, c, device="cuda") for c in base]
        dW = torch.empty(G * L, 9, ctot, device="cuda")
        db = torch.empty(G * L, device="cuda")
        ws = torch.empty(lib.otgan_workspace_bytes_dense_bgrad(ctypes.byref(geom)) // 4 + 64, device="cuda")
        _lib.check(lib.otgan_dense_build_wb_f32(ctypes.byref(geom), wf.data_ptr(), WB.data_ptr(), st), "build_wb")
        ops = {"fprop": lambda: lib.otgan_dense_block_fprop_tf32(ctypes.byref(geom), wf.data_ptr(), bias.data_ptr(), Z.data_ptr(), S.data_ptr(), st),
               "bgrad": lambda: lib.otgan_dense_block_bgrad_tf32(ctypes.byref(geom), Z.data_ptr(), dZ.data_ptr(), WB.data_ptr(), dY.data_ptr(),
                                                                 _lib.ptr_array([t.data_ptr() for t in dbase]), dW.data_ptr(), db.data_ptr(),
                                                                 ws.data_ptr(), ws.numel() * 4, st)}
        flops = 2.0 * B * H * W * 9 * G * sum(2 * (c0 + G * r) for r in range(L))
        row = {"shape": [B, H, W, base, L], "channels": ctot, "gflop": flops / 1e9}
        for op, fn in ops.items():
            for _ in range(2):
                _lib.check(fn(), op)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            nf = 1.0 if op == "fprop" else 2.0           # backward = input gradients + filter gradients
            row[op] = {"ms": ms, "tflops_algorithmic": nf * flops / ms / 1e9}
        out[name] = row
        del Z, S, wf, dZ, WB, dY, dbase, dW, ws
        torch.cuda.empty_cache()
    return out


def conv_roofline(conv, tf32_peak=None):
    """Roofline object of the step's dominant kernel, conv_gemm_tc_kernel<256> (fprop / dgrad; 49% of the step in the ncu
    launch list profiles/r01_p_train_launches.txt), on the launch that was also captured with ncu --set full."""
    pk = peaks()
    full = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    r = conv[CONV_ROOFLINE_LAUNCH]["fprop"]
    gemm = [conv[n][op] for n in conv for op in ("fprop", "dgrad")]
    flops = [conv[n]["gflop"] for n in conv for _ in ("fprop", "dgrad")]
    mean_tf = sum(flops) / sum(g["ms"] for g in gemm)                        # GFLOP / ms = TFLOP/s
    return {"bound": "tensor", "kernel": "conv_gemm_tc_kernel<256> (%s fprop, %.1f GFLOP executed per launch)" % (CONV_ROOFLINE_LAUNCH, conv[CONV_ROOFLINE_LAUNCH]["gflop"]),
            "achieved": r["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": r["tflops"] / pk["bf16_tflops"],
            "traffic": CONV_ROOFLINE_TRAFFIC,
            "traffic_source": "constant from profiles/r01_l_conv_up2_fprop_ncu_full.txt (ncu --set full of this launch), not live",
            "tf32_tflops_cublas_measured_in_run": tf32_peak,
            "frac_of_measured_tf32": (r["tflops"] / tf32_peak) if tf32_peak else None,
            "frac_of_tf32_ceiling": r["tflops"] / (pk["bf16_tflops"] / 2.0),
            "achieved_all_fprop_dgrad_launches": mean_tf,
            "frac_of_tf32_ceiling_all_fprop_dgrad_launches": mean_tf / (pk["bf16_tflops"] / 2.0),
            "bf16_tflops_sustained": full.get("bf16_tflops_sustained"),
            "peak_source": pk["source"],
            "note": "operands are fp32 tensors consumed as TF32 (kind::tf32): the tensor pipe runs TF32 at half the bf16 rate, so "
                    "the ceiling of this kernel is peak/2; `frac` is against the measured dense bf16 burst peak as the contract "
                    "asks, frac_of_tf32_ceiling against half of it.  ncu: tensor pipe 84.5% active, DRAM 6% of peak (traffic field); "
                    "the kernel runs power-capped (sw_power_cap, ~1.7 GHz)."}


def time_config(torch, dist, T, devv, rank, world, n_total, t_iters, model="dcgan", image_size=32, steps=12, warmup=6, graphs=True):
    """One BASELINE.json configuration as a training-step timing (device-resident images, CUDA events, max over ranks):
    n_total real + n_total generated images per step split over the ranks, 2 towers per rank."""
    from otgan_b200 import _lib
    towers = 2 * world
    argv = ["--synthetic", "--nr_gpu", str(towers), "--batch_size", str(n_total // towers), "--nr_sinkhorn_iter", str(t_iters),
            "--sinkhorn_lambda", str(LAMBDA), "--model", model, "--image_size", str(image_size)]
    tr = T.Trainer(T.build_parser().parse_args(argv), devv, rank, world)
    graphs_on = False
    if graphs:
        try:
            tr.enable_cuda_graphs()
            graphs_on = True
        except Exception as e:
            sys.stderr.write("bench.py: CUDA-graph capture failed for %s (%r); eager launches\n" % (model, e))
            tr.graphs = None
        flag = torch.tensor([1.0 if graphs_on else 0.0], device=devv)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() < 1.0:
            tr.graphs, graphs_on = None, False
    bs = tr.bs_local
    gen = torch.Generator().manual_seed(1 + rank)
    imgs = [(torch.rand((bs, image_size, image_size, 3), generator=gen) * 2 - 1).to(devv) for _ in range(4)]
    stream = torch.cuda.current_stream()
    for i in range(max(warmup, 3)):
        tr.step(imgs[i % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.reset_launch_count()
    tr.replayed_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        kind, stats = tr.step(imgs[i % 4])
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=devv, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    d, e = stats.tolist()
    res = {"model": model, "N": n_total, "h": n_total // 2, "D": tr.num_features, "T": t_iters, "image_size": image_size, "ranks": world,
           "images_per_rank": n_total // world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
           "images_per_sec": n_total / (ms * 1e-3), "cuda_graphs": graphs_on, "gpu_launches": _lib.launch_count() + tr.replayed_launches,
           "last_distance": d, "last_entropy": e}
    tr.graphs = tr.g_stats = None
    del tr, imgs
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from otgan_b200 import _lib
    from otgan_b200 import train as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    devv = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=devv)
    _lib.load()
    assert N_TOTAL % (2 * world) == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.only_config:                           # profiling aid: one named configuration, e.g. densenet:256:100[:image_size]
        model, n_total, t_iters = args.only_config.split(":")[:3]
        isz = int(args.only_config.split(":")[3]) if args.only_config.count(":") >= 3 else 32
        res = time_config(torch, dist, T, devv, rank, world, int(n_total), int(t_iters), model=model, image_size=isz,
                          steps=args.steps, warmup=args.warmup, graphs=bool(args.cuda_graphs))
        if rank == 0:
            print(json.dumps(res))
        if world > 1:
            dist.barrier()
            os._exit(0)
        return
    if args.workload == "conv":                    # profiling aid: the per-layer convolution launches only (ncu -k regex:conv_)
        if rank == 0:
            res = conv_benchmark(torch)
            print(json.dumps({"conv_layers": res, "roofline": conv_roofline(res)}))
        return
    match_res, roof, conv_res, dense_res = (None, None, None, None)
    if rank == 0 and not args.step_only:
        match_res, match_roof = matching_benchmark(torch, devv, 100, 5)
        if args.n_total:                           # diagnostics run at another N: keep the matching roofline only
            roof = match_roof
        else:
            conv_res = conv_benchmark(torch)
            roof = conv_roofline(conv_res, measure_tf32_peak(torch))
            match_res["roofline_matching"] = match_roof
            # cfg4 per-rank shapes: the critic sees 2 * 256 / max(world, 1) images, the generator 256 / world
            dense_res = densenet_benchmark(torch, 2 * N_TOTAL // world, N_TOTAL // world)
    barrier()

    # ---- the training step (all ranks): towers = 2 per rank, N_TOTAL images in total
    towers = 2 * world
    targs = T.build_parser().parse_args(["--synthetic", "--nr_gpu", str(towers), "--batch_size", str(N_TOTAL // towers),
                                         "--nr_sinkhorn_iter", str(T_ITERS), "--sinkhorn_lambda", str(LAMBDA)])
    tr = T.Trainer(targs, devv, rank, world)
    graphs_on = False
    if args.cuda_graphs:
        try:
            tr.enable_cuda_graphs()
            graphs_on = True
        except Exception as e:                     # never lose the measurement to a capture problem: launch eagerly instead
            sys.stderr.write("bench.py: CUDA-graph capture failed (%r); falling back to eager launches\n" % (e,))
            tr.graphs = None
        flag = torch.tensor([1.0 if graphs_on else 0.0], device=devv)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)       # all ranks replay, or none does
        if flag.item() < 1.0:
            tr.graphs, graphs_on = None, False
    bs = tr.bs_local
    gen = torch.Generator().manual_seed(1 + rank)
    host_imgs = [(torch.rand((bs, 32, 32, 3), generator=gen) * 2 - 1).pin_memory() for _ in range(8)]
    dev_imgs = [h.to(devv) for h in host_imgs]
    stream = torch.cuda.current_stream()
    W = max(args.warmup, 3)
    if args.workload == "train":
        for i in range(W):
            tr.step(dev_imgs[i % 8])
        stats_host = torch.empty(2).pin_memory()
        x_dev = torch.empty((bs, 32, 32, 3), device=devv)

        def e2e_input(i):
            """The public API takes the step's images as a tensor: under CUDA graphs Trainer.step copies them straight from the
            pinned HOST buffer into the graph's input (one H2D copy); eagerly they are uploaded first (also one H2D copy)."""
            if tr.graphs is not None:
                return host_imgs[i % 8]
            x_dev.copy_(host_imgs[i % 8], non_blocking=True)
            return x_dev

        def measure():
            sampler = ClockSampler(local)
            sampler.start()
            barrier()
            _lib.reset_launch_count()
            tr.replayed_launches = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(args.steps):
                tr.step(dev_imgs[i % 8])
            e1.record(stream)
            barrier()
            n_launch = _lib.launch_count() + tr.replayed_launches      # eager launches + kernels inside replayed graphs
            t_dev = e0.elapsed_time(e1)
            # ---- end to end through the public API, the way otgan_b200.train.main drives it: pinned host images -> Trainer.stage
            # (H2D on a copy stream) -> Trainer.step -> Trainer.fetch_async (D2H of [distance, entropy]) -> host read, EVERY step.
            # Software-pipelined: the upload of step k+1 and the host's read of step k-1 happen while step k computes.
            def e2e_pipelined(nsteps):
                staged = tr.stage(host_imgs[0])
                pending = None
                for i in range(nsteps):
                    _, st = tr.step(staged)
                    staged = tr.stage(host_imgs[(i + 1) % 8])
                    hnd = tr.fetch_async(st)
                    if pending is not None:
                        pending.result()               # the host reads (distance, entropy) of the previous step
                    pending = hnd
                return pending.result()

            e2e_pipelined(2)
            barrier()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record(stream)
            e2e_pipelined(args.steps)
            e3.record(stream)
            barrier()
            t_e2e = e2.elapsed_time(e3)
            # ---- the same with sess.run semantics (upload, step, read back, block -- every step): what the pipelining buys
            for i in range(2):
                stats_host.copy_(tr.step(e2e_input(i))[1], non_blocking=True)
            barrier()
            e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e4.record(stream)
            for i in range(args.steps):
                stats_host.copy_(tr.step(e2e_input(i))[1], non_blocking=True)
                stream.synchronize()                   # the host reads (distance, entropy) every step, like sess.run
            e5.record(stream)
            barrier()
            measure.t_blocking = e4.elapsed_time(e5)
            sampler.stop_flag = True
            sampler.join()
            return t_dev, t_e2e, n_launch, sampler

        measure.t_blocking = 0.0
        ms_total, ms_e2e, launches, sampler = measure()
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        remeasured = False
        flag = torch.tensor([1.0 if bad & set(sampler.result().get("reasons", [])) else 0.0], device=devv)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if flag.item() > 0:                            # thermal / hardware slowdown seen: rejected, measured once more
            ms_total, ms_e2e, launches, sampler = measure()
            remeasured = True
        h2d, d2h = bs * 32 * 32 * 3 * 4 * world, 8
        ms_blocking = measure.t_blocking
    else:
        sampler = ClockSampler(local)
        sampler.start()
        res2, _ = matching_benchmark(torch, devv, args.steps, W) if rank == 0 else (None, None)
        sampler.stop_flag = True
        sampler.join()
        ms_total = (res2["ms_per_step"] * args.steps) if rank == 0 else 0.0
        ms_e2e, launches, h2d, d2h = ms_total, int(res2["gpu_launches_per_step"] * args.steps) if rank == 0 else 0, 0, 0
        ms_blocking = 0.0
        remeasured = False

    t = torch.tensor([ms_total, ms_e2e], device=devv, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    tr.graphs = tr.g_stats = None
    del tr
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations that fit this launch, the weak-scaling line, the multi-rank parity check
    configs, parity = {}, None
    if args.workload == "train" and not args.step_only and not args.n_total:
        def guarded(name, **kw):
            try:
                configs[name] = time_config(torch, dist, T, devv, rank, world, graphs=bool(args.cuda_graphs), **kw)
            except Exception as e:                 # an extra line must never cost the contract line
                configs[name] = {"error": repr(e)[:300]}
                sys.stderr.write("bench.py: config %s failed: %r\n" % (name, e))
        if world == 1:
            guarded("cfg2_dcgan_n128_T100", n_total=128, t_iters=100)
            guarded("cfg3_dcgan_n256_T500", n_total=256, t_iters=500)
            guarded("cfg4_densenet_n256_T100_1gpu", n_total=256, t_iters=100, model="densenet")
        if world == 4:
            guarded("cfg4_densenet_n256_T100", n_total=256, t_iters=100, model="densenet")
        if world == 8:
            guarded("cfg5_dcgan64_n512_T100", n_total=512, t_iters=100, image_size=64)
        if world > 1:
            guarded("weak_dcgan_n%d_T100" % (128 * world), n_total=128 * world, t_iters=100)
            try:
                parity = T.parity_check(world, rank, devv)
            except Exception as e:
                parity = {"ok": False, "error": repr(e)[:300]}

    if rank == 0:
        ms_step = ms_total / args.steps
        value = N_TOTAL / (ms_step * 1e-3)
        line = {
            "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32-conv/f32", "data": "synthetic", "config": workload_config(world, args.workload),
            "sinkhorn_iters_per_sec": match_res["sinkhorn_iters_per_sec"] if match_res else None,
            "roofline": roof, "conv_layers": conv_res, "densenet_blocks": dense_res, "matching": match_res,
            "e2e": {"value": N_TOTAL / (ms_e2e / args.steps * 1e-3), "unit": "images/sec", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "Trainer.stage (H2D of the next step's pinned host images on a copy stream) -> Trainer.step -> Trainer.fetch_async "
                           "(D2H of [distance, entropy]) -> host read of every step's result one step behind the GPU: the loop of otgan_b200.train.main",
                    "blocking_ms_per_step": ms_blocking / args.steps if ms_blocking else None},
            "gpu_launches": launches, "cuda_graphs": graphs_on,
            "clocks": dict(sampler.result(), remeasured_after_slowdown=remeasured),
            "configs": configs, "mgpu_parity": parity,
        }
        if world == 1 and not args.no_cpu and not args.step_only:
            mt, ph, cores = cpu_matching_time(N_TOTAL, D_FEAT, T_ITERS, LAMBDA, 3, 1)
            best = float(np.min(mt))
            tt, tcores, ctr, _ = cpu_train_time(N_TOTAL, 3, 0)     # 1 critic + 2 generator steps at the contract size (~25 s)
            t_cycle = tt[0] + 5.0 * float(np.mean(tt[1:]))         # one 1:5 cycle
            line["cpu_baseline"] = {"value": 6 * N_TOTAL / t_cycle, "unit": "images/sec", "cores": tcores, "kind": "port",
                                    "sample": "1 critic + 2 generator training steps of the SAME DCGAN step at the contract size "
                                              "N=%d real + %d generated images (~25 s of CPU work, no warm-up), combined as one "
                                              "1 critic : 5 generator cycle; torch-CPU conv stacks + oracle/torch_oracle.py matching (one "
                                              "torch-CPU op per TensorFlow op); TensorFlow 1.x not installable offline" % (N_TOTAL, N_TOTAL),
                                    "ms_per_step": t_cycle / 6 * 1e3, "ms_critic_step": tt[0] * 1e3, "ms_generator_step": float(np.mean(tt[1:])) * 1e3,
                                    "host": cpu_info(),
                                    "matching_phase_torch_cpu": {"ms": float(np.mean(ctr.phase_s["matching"])) * 1e3,
                                                                 "phase_ms": {k: float(np.mean(ctr.phase_s[k])) * 1e3 for k in ("cost", "sinkhorn", "matched")},
                                                                 "sinkhorn_iters_per_sec": T_ITERS / float(np.mean(ctr.phase_s["sinkhorn"]))},
                                    "matching_phase_c_oracle": {"ms": best * 1e3, "images_per_sec": N_TOTAL / best, "cores": cores,
                                                                "phase_ms": {"cost": ph[0], "sinkhorn": ph[1], "matched_distance": ph[2]},
                                                                "sinkhorn_iters_per_sec": T_ITERS / (ph[1] * 1e-3)}}
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # Tear-down: the graphs (which hold captured NCCL kernels) are gone and the device is drained; destroy the process
        # group cleanly, but under a watchdog -- in round 1 the collective destructor was seen to hang after the JSON line
        # had been printed, and a hung rank would cost the whole scaling run.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()

        th = threading.Thread(target=_destroy, daemon=True)
        th.start()
        if not done.wait(20.0):
            sys.stderr.write("bench.py: destroy_process_group did not return within 20 s on rank %d; leaving without it\n" % rank)
            sys.stderr.flush()
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "matching", "conv"],
                    help="train: the contract workload; matching / conv: only that sub-benchmark (profiling aid)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cuda-graphs", type=int, default=1, help="replay the training step from captured CUDA graphs (1) or launch eagerly (0)")
    ap.add_argument("--n-total", type=int, default=0, help="diagnostics only: override N (images per step); the contract workload is N=256")
    ap.add_argument("--only-config", default="", help="profiling aid: time one configuration model:N:T[:image_size] and print its line")
    ap.add_argument("--step-only", action="store_true", help="profiling aid: only the contract training step (no sub-benchmarks, extra configs or CPU leg)")
    args = ap.parse_args()
    if args.n_total:
        global N_TOTAL
        N_TOTAL = args.n_total
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
