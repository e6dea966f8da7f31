#!/usr/bin/env python
"""bench.py -- OT-GAN matching hot path on B200: images/sec + Sinkhorn-iters/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload headline|cfg2|cfg3|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the matching hot path over one batch of synthetic critic embeddings:
cost blocks (utils/matching.py:21-43) -> T Sinkhorn iterations (:46-61) -> feature gradients + distance + entropy
(:63-83, :139-153, train.py:111,125-126), i.e. everything `train.py` does between the critic forward and backward.
N real + N fake images are consumed per step, so images/sec = N / step time.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the public
Python API (otgan_b200.utils.matching.matching_step) with pinned HOST buffers, H2D and D2H inside the timed region.
`--impl reference` times the reference algorithm's CPU restatement (oracle/, C+OpenMP on all host cores; TensorFlow 1.x
is not installable here) on the same workload; it is the one place besides `cpu_baseline` that executes oracle/.
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

WORKLOADS = {
    # name: (N, D, T, lam)   h = N/2
    "headline": (256, 32768, 100, 500.0),     # BASELINE.json metric: B=256, 100 iters (DCGAN critic width)
    "cfg2": (128, 32768, 100, 500.0),
    "cfg3": (256, 32768, 500, 500.0),
    "cfg4": (256, 7296, 100, 500.0),          # DenseNet critic width
}
N_INPUT_SETS = 4      # rotate 4 x 64 MiB embedding sets (> 126 MB L2) so no step finds its inputs in L2


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


def cpu_reference_time(N, D, T, lam, reps, warmup):
    """The reference algorithm on the host cores (C+OpenMP restatement, all threads): returns (best seconds, phases, cores)."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, D, T, lam = WORKLOADS[args.workload]
    times, phases, cores = cpu_reference_time(N, D, T, lam, args.steps, args.warmup)
    mean = float(np.mean(times))
    val = N / mean
    line = {
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, 1),
        "sinkhorn_iters_per_sec": T / (phases[1] * 1e-3),
        "cpu_baseline": {"value": val, "unit": "images/sec", "cores": cores, "kind": "port",
                         "sample": "%d full matching steps (cost+Sinkhorn+matched+distance), C+OpenMP restatement of "
                                   "utils/matching.py; TensorFlow 1.x not installable offline" % args.steps,
                         "phase_ms": {"cost": phases[0], "sinkhorn": phases[1], "matched_distance": phases[2]}},
        "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(name, world):
    N, D, T, lam = WORKLOADS[name]
    return {"workload": "OT-GAN matching hot path (6 cosine-cost blocks + %d Sinkhorn iters + feature gradients + "
                        "distance/entropy), %s: N=%d real + %d fake embeddings per step, h=%d, D=%d, lambda=%g, T=%d"
                        % (T, name, N, N, N // 2, D, lam, T),
            "N": N, "h": N // 2, "D": D, "T": T, "lambda": lam, "towers": 2, "ranks": world,
            "l2_policy": "inputs rotate over %d sets of 2x[N,D] fp32 (%.0f MiB total) > 126 MB L2" %
                         (N_INPUT_SETS, N_INPUT_SETS * 2 * N * D * 4 / 2**20)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from otgan_b200 import _lib
    from otgan_b200.utils import matching as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    devv = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=devv)
    N, D, T, lam = WORKLOADS[args.workload]
    h = N // 2
    _lib.load()

    # ---- inputs resident in HBM (rotating sets) and pinned host copies for the e2e leg
    host_sets = []
    for s in range(N_INPUT_SETS):
        a = torch.from_numpy(synth(N, D, 100 + 2 * s + 1000 * rank)).pin_memory()
        b = torch.from_numpy(synth(N, D, 101 + 2 * s + 1000 * rank)).pin_memory()
        host_sets.append((a, b))
    dev_sets = [(a.to(devv), b.to(devv)) for a, b in host_sets]
    stream = torch.cuda.current_stream()

    def step(A, B):
        return M.matching_step(list(torch.chunk(A, 2, 0)), list(torch.chunk(B, 2, 0)), lam, T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(*dev_sets[i % N_INPUT_SETS])
    barrier()

    # ---- per-kernel breakdown (events around each ABI call), used for the roofline object
    lib = _lib.load()
    phases = {"cost": [], "sinkhorn": [], "grad": []}
    for i in range(12):
        A, B = dev_sets[i % N_INPUT_SETS]
        a1, a2, b1, b2 = A[:h], A[h:], B[:h], B[h:]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        L = M.cost_blocks([a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2], lam)
        ev[1].record(stream)
        P, ent, pc = M.sinkhorn(L, lam, T)
        ev[2].record(stream)
        Ga = torch.empty_like(A)
        Gb = torch.empty_like(B)
        ev[3] = torch.cuda.Event(enable_timing=True)
        e_start = torch.cuda.Event(enable_timing=True)
        e_start.record(stream)
        rc = lib.otgan_grad_features_f32(h, D, P.data_ptr(), A.data_ptr(), B.data_ptr(), D, Ga.data_ptr(), Gb.data_ptr(),
                                         D, 0, stream.cuda_stream)
        ev[3].record(stream)
        torch.cuda.synchronize()
        assert rc == 0
        if i >= 2:
            phases["cost"].append(ev[0].elapsed_time(ev[1]))
            phases["sinkhorn"].append(ev[1].elapsed_time(ev[2]))
            phases["grad"].append(e_start.elapsed_time(ev[3]))
    kernel_ms = {k: float(np.mean(v)) for k, v in phases.items()}

    # ---- timed region 1: device-resident inputs (value)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(*dev_sets[i % N_INPUT_SETS])
    e1.record(stream)
    barrier()
    launches = _lib.launch_count()
    ms_total = e0.elapsed_time(e1)

    # ---- timed region 2: end to end through the public API with pinned host buffers (H2D + D2H inside)
    dA, dB = torch.empty((N, D), device=devv), torch.empty((N, D), device=devv)
    stats_host = torch.empty(2).pin_memory()
    for i in range(3):
        dA.copy_(host_sets[i % N_INPUT_SETS][0], non_blocking=True)
        dB.copy_(host_sets[i % N_INPUT_SETS][1], non_blocking=True)
        stats_host.copy_(step(dA, dB)[2], non_blocking=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for i in range(args.steps):
        ha, hb = host_sets[i % N_INPUT_SETS]
        dA.copy_(ha, non_blocking=True)
        dB.copy_(hb, non_blocking=True)
        stats_host.copy_(step(dA, dB)[2], non_blocking=True)
        stream.synchronize()          # the host reads the step's (distance, entropy) every step, like sess.run
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    sampler.stop_flag = True
    sampler.join()

    t = torch.tensor([ms_total, ms_e2e], device=devv, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        ms_step = ms_total / args.steps
        images = N * world                       # every rank processes its own batch (independent matching problems)
        value = images / (ms_step * 1e-3)
        e2e_val = images / (ms_e2e / args.steps * 1e-3)
        alg_bytes = {"cost": 4.0 * 2 * N * D + 24.0 * h * h, "sinkhorn": 48.0 * h * h, "grad": 16.0 * N * D}
        dom = max(kernel_ms, key=kernel_ms.get)
        kernels = {}
        for k in kernel_ms:
            gbs = alg_bytes[k] / (kernel_ms[k] * 1e-3) / 1e9
            kernels[k] = {"ms": kernel_ms[k], "alg_bytes": alg_bytes[k], "achieved_gbs": gbs, "hbm_frac": gbs / pk["hbm_gbs"]}
        kernels["cost"]["alg_flops"] = 12.0 * h * h * D
        kernels["grad"]["alg_flops"] = 24.0 * h * h * D
        kernels["sinkhorn"]["exp_per_sec"] = 12.0 * h * h * T / (kernel_ms["sinkhorn"] * 1e-3)
        kernels["sinkhorn"]["streaming_form_bytes"] = 96.0 * h * h * T
        ach = kernels[dom]["achieved_gbs"]
        line = {
            "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, world),
            "sinkhorn_iters_per_sec": T / (kernel_ms["sinkhorn"] * 1e-3),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"]},
            "kernels": kernels,
            "e2e": {"value": e2e_val, "unit": "images/sec", "h2d_bytes_per_step": 2 * N * D * 4, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": sampler.result(),
        }
        if world == 1 and not args.no_cpu:
            times, ph, cores = cpu_reference_time(N, D, T, lam, 5, 1)
            best = float(np.min(times))
            line["cpu_baseline"] = {"value": N / best, "unit": "images/sec", "cores": cores, "kind": "port",
                                    "sample": "best of 5 full matching steps of the same workload (C+OpenMP restatement "
                                              "of utils/matching.py, all host threads)",
                                    "ms_per_step": best * 1e3,
                                    "phase_ms": {"cost": ph[0], "sinkhorn": ph[1], "matched_distance": ph[2]},
                                    "sinkhorn_iters_per_sec": T / (ph[1] * 1e-3)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
