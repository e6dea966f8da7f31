"""Re-hosted training loop of the reference's train.py (flags, step schedule, hand-built grad_ys, summed tower
gradients, Adam with the critic ascending, generator EMA, logging) on the B200 path.

    python -m otgan_b200.train --synthetic --nr_gpu 2 --batch_size 128 --nr_sinkhorn_iter 100 --max_steps 20
    torchrun --nproc-per-node 8 -m otgan_b200.train --synthetic --nr_gpu 8 --batch_size 64 ...

Reference correspondence (train.py line numbers):
    flags :14-33 (same names and defaults; additions: --synthetic, --max_steps, --log_every)
    init pass :52-54, parameter split :61-62, EMA :63-64
    towers :72-85  -> one process per GPU (torch.distributed, NCCL); each rank hosts nr_gpu / world_size towers
    matching :88-97, distances :101-105, grad_ys :108-130 -> matching.matching_step (fused) or the list API
    tower-gradient SUM :134-139 -> all_reduce(SUM) of the flat gradient buffer
    adam_updates with +lr (generator) / -lr (critic) :142-143, EMA on generator steps :223
    schedule: critic step when step_counter % (nr_gen_per_disc + 1) == 0 :214-226
The embedding exchange is ONE all-gather of the [2 * bs_local, D] feature slab per step (the reference's tf.concat,
utils/matching.py:16-19); cost + Sinkhorn are then replicated on every rank (cheap) and each rank back-propagates its
own rows of grad_ys.  Inception score / PNG tiles are out of scope.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from .data import cifar10_data
from .utils import matching, nn


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--seed', type=int, default=1)
    parser.add_argument('--batch_size', type=int, default=625)
    parser.add_argument('--learning_rate_disc', type=float, default=0.0003)
    parser.add_argument('--learning_rate_gen', type=float, default=0.0003)
    parser.add_argument('--data_dir', type=str, default='/home/tim/data')
    parser.add_argument('--save_dir', type=str, default='/local_home/tim/med_gan')
    parser.add_argument('--optimizer', type=str, default='adam')
    parser.add_argument('--nonlinearity', type=str, default='crelu')
    parser.add_argument('--nr_gpu', type=int, default=8, help='How many towers (per-GPU batches) to distribute the training across?')
    parser.add_argument('--nr_gen_per_disc', type=int, default=5, help='How many times to update the generator for each update of the discriminator?')
    parser.add_argument('--sinkhorn_lambda', type=float, default=500.)
    parser.add_argument('--nr_sinkhorn_iter', type=int, default=500)
    parser.add_argument('--single_batch', dest='single_batch', action='store_true', help='Use simplified batching using a single batch instead of 2')
    parser.add_argument('--train_disc_against_ema', dest='train_disc_against_ema', action='store_true', help='Should discriminator be trained against samples of EMA generator?')
    parser.add_argument('--model', type=str, default='dcgan')
    parser.add_argument('--load_params', dest='load_params', action='store_true')
    parser.add_argument('--model_name', type=str, default='med_gan_params-2399')
    parser.add_argument('--no_sinkhorn', dest='no_sinkhorn', action='store_true')
    # additions (not in the reference)
    parser.add_argument('--synthetic', action='store_true', help='CIFAR-10-shaped uniform noise instead of the dataset')
    parser.add_argument('--max_steps', type=int, default=0, help='stop after this many steps (0 = run like the reference)')
    parser.add_argument('--log_every', type=int, default=0, help='also print a line every N steps')
    parser.add_argument('--cuda_graphs', action='store_true', help='replay each step from a captured CUDA graph')
    parser.add_argument('--image_size', type=int, default=32, help='synthetic image side (extension: the reference is hard-wired to 32; dcgan only)')
    return parser


class _Fetched:
    """Handle of Trainer.fetch_async: .result() -> [distance, entropy] once the device-to-host copy has completed."""

    def __init__(self, buf, event):
        self.buf, self.event = buf, event

    def result(self):
        self.event.synchronize()
        return self.buf.tolist()


class Trainer:
    """One data-parallel rank of the OT-GAN step.  `step(x_real_local)` is one sess.run of train.py:214-226."""

    def __init__(self, args, device, rank=0, world=1):
        assert args.nr_gpu % 2 == 0                                            # train.py:34
        assert args.nr_gpu % world == 0, "nr_gpu (towers) must be a multiple of the number of ranks"
        self.args, self.device, self.rank, self.world = args, device, rank, world
        self.towers_local = args.nr_gpu // world
        self.bs_local = self.towers_local * args.batch_size
        if args.model == 'dcgan':
            from .models.dcgan import generator, discriminator
        elif args.model == 'densenet':
            from .models.densenet import generator, discriminator
        else:
            raise ValueError(args.model)
        generator.reset(); discriminator.reset()
        self.generator, self.discriminator = generator, discriminator
        self.model_opts = {'batch_size': self.bs_local, 'nonlinearity': args.nonlinearity}      # train.py:45
        self.image_size = getattr(args, 'image_size', 32)
        if self.image_size != 32:
            assert args.model == 'dcgan', "--image_size is an extension of the DCGAN models only"
            self.model_opts['image_size'] = self.image_size
        torch.manual_seed(args.seed)                                                            # :48-49
        # run once for (data dependent) initialization of parameters                              :52-54
        x_init = torch.zeros((args.batch_size, self.image_size, self.image_size, 3), device=device)
        with torch.no_grad():
            f = discriminator(x_init + 0.1, init=True, device=device, **self.model_opts)
            generator(init=True, device=device, **dict(self.model_opts, batch_size=args.batch_size))
        self.num_features = f.shape[-1]
        if world > 1:                                                                            # identical replicas
            dist.broadcast(discriminator.flat.data, 0)
            dist.broadcast(generator.flat.data, 0)
        torch.manual_seed(args.seed + 1000 * rank + 1)          # towers draw independent latents (tf.random_uniform per tower)
        self.ema = nn.ExponentialMovingAverage(decay=0.999).attach(generator)                   # :63-64
        opt = {'adam': nn.adam_updates, 'adamax': nn.adamax_updates, 'nesterov': nn.nesterov_updates}.get(args.optimizer)
        if opt is None:
            raise ValueError('unsupported optimizer')                                            # :151
        self.sync = {'disc': GradSync(discriminator, world), 'gen': GradSync(generator, world)}
        self.sharded_opt = args.optimizer == 'adam' and self.sync['gen'].sharded
        if not self.sharded_opt:
            for sy in self.sync.values():
                sy.sharded = False
        shard = (rank, world) if self.sharded_opt else None
        if args.optimizer == 'adam':
            self.gen_optimizer = opt(generator, lr=args.learning_rate_gen, mom1=0.5, mom2=0.999, ema=self.ema, shard=shard)
            self.disc_optimizer = opt(discriminator, lr=-args.learning_rate_disc, mom1=0.5, mom2=0.999, shard=shard)
        elif args.optimizer == 'adamax':
            self.gen_optimizer = opt(generator, lr=args.learning_rate_gen, mom1=0.5, mom2=0.999)
            self.disc_optimizer = opt(discriminator, lr=-args.learning_rate_disc, mom1=0.5, mom2=0.999)
        else:
            self.gen_optimizer = opt(generator, lr=args.learning_rate_gen, mom1=0.5)
            self.disc_optimizer = opt(discriminator, lr=-args.learning_rate_disc, mom1=0.5)
        self.match_all_rows = False                        # parity checks: compute grad_ys for every row on every rank
        self.step_counter = 0
        self.gather_buf = None
        self.graphs = None                                 # set by enable_cuda_graphs()
        self.replayed_launches = 0                         # libotgan kernels executed through graph replays
        self._stage = None                                 # stage(): double-buffered upload of the NEXT step's images
        self._fetch = None                                 # fetch_async(): double-buffered pinned read-back of [distance, entropy]

    def _gather_features(self, f_gen, f_dat):
        return gather_features(f_gen, f_dat, self.world)

    def _match(self, A, B, rows=None):
        """A (fake) / B (real): [N, D] features of ALL towers -> (Ga, Gb, [dist, entropy]); rows: the row range this rank needs."""
        a = self.args
        G = a.nr_gpu
        fa, fb = list(torch.chunk(A, G, 0)), list(torch.chunk(B, G, 0))
        if a.single_batch or a.no_sinkhorn:
            if a.single_batch:
                m = matching.get_matched_features_single_batch(fa, fb, a.sinkhorn_lambda, a.nr_sinkhorn_iter)
            else:
                m = matching.get_matched_features_random(fa, fb)
            d = matching.calc_distance(fa, fb, m)
            Ga = torch.cat([x - y for x, y in zip(m[0], m[2])], 0)                               # train.py:111
            Gb = torch.cat([x - y for x, y in zip(m[1], m[3])], 0)                               # train.py:126
            return Ga, Gb, torch.stack([d, m[4].to(d.dtype)])
        # partition "S" (SURVEY 8e): blocks larger than one 128-row tile shard their cost rows over the ranks; at h <= 128 a rank's
        # slab is a fraction of ONE tensor-core tile, so replicating the 66 us cost stage is cheaper than the extra collective
        shard = (self.rank, self.world) if (rows is not None and A.shape[0] // 2 > 128 and self.world % 2 == 0
                                            and os.environ.get("OTGAN_SHARD_COST", "1") == "1") else None
        ga, gb, stats = matching.matching_step(fa, fb, a.sinkhorn_lambda, a.nr_sinkhorn_iter, rows=rows, shard=shard)
        return matching._gather(ga), matching._gather(gb), stats      # the tower chunks of one [N, D] buffer: a view, no copy

    def step(self, x_real, u=None, apply_update=True):
        """x_real: this rank's [bs_local, 32, 32, 3] real images in [-1, 1].  Returns ('disc'|'gen', stats[2] tensor).
        `u` optionally fixes this rank's generator latents (parity tests); `apply_update=False` skips the optimiser and
        leaves the summed gradient in self.last_grad.  After enable_cuda_graphs() the step is one graph replay."""
        a = self.args
        train_disc = self.step_counter % (a.nr_gen_per_disc + 1) == 0                            # :214
        kind = 'disc' if train_disc else 'gen'
        slot = self._staged_slot(x_real)
        if slot is not None:                               # images uploaded by stage(): wait for that copy, not for the host
            torch.cuda.current_stream().wait_event(self._stage['ready'][slot])
        if self.graphs is not None and apply_update:
            stats = self._replay(kind, x_real, u)
        else:
            stats = self._step_body(kind, x_real, u, apply_update)
            self.step_counter += 1
        if slot is not None:                               # the staging buffer may be refilled once this step has consumed it
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._stage['consumed'][slot] = ev
        return kind, stats

    # ---- input / output pipelining around step(): the upload of step k+1 and the read-back of step k run while the GPU computes ----
    def stage(self, x_host):
        """Start uploading the images of a LATER step (pinned host tensor or numpy-backed CPU tensor, [bs_local, S, S, 3]) on a
        copy stream and return the device tensor to hand to step().  Called right after step(k) has been enqueued, the H2D copy
        of step k+1 overlaps step k's kernels; step() makes its own stream wait for the copy.  Two buffers alternate."""
        if self._stage is None:
            shape = (self.bs_local, self.image_size, self.image_size, 3)
            self._stage = {'bufs': [torch.empty(shape, device=self.device) for _ in range(2)], 'ready': [torch.cuda.Event(), torch.cuda.Event()],
                           'consumed': [None, None], 'stream': torch.cuda.Stream(device=self.device), 'next': 0}
        st = self._stage
        i = st['next']
        st['next'] ^= 1
        cs = st['stream']
        if st['consumed'][i] is not None:
            cs.wait_event(st['consumed'][i])               # the step that last read this buffer
        with torch.cuda.stream(cs):
            st['bufs'][i].copy_(x_host, non_blocking=True)
            st['ready'][i].record(cs)
        return st['bufs'][i]

    def _staged_slot(self, x):
        if self._stage is None or not isinstance(x, torch.Tensor):
            return None
        for i, b in enumerate(self._stage['bufs']):
            if x is b:
                return i
        return None

    def fetch_async(self, stats):
        """Queue the read-back of a step's [distance, entropy] into pinned host memory and return a handle whose .result() gives the
        two floats (blocking only until THAT copy has landed).  Reading step k's handle after step k+1 has been enqueued keeps the
        GPU busy across the host's read (the reference's sess.run blocks on every step)."""
        if self._fetch is None:
            self._fetch = {'bufs': [torch.empty(2).pin_memory() for _ in range(2)], 'next': 0}
        f = self._fetch
        i = f['next']
        f['next'] ^= 1
        f['bufs'][i].copy_(stats, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return _Fetched(f['bufs'][i], ev)

    def _step_body(self, kind, x_real, u, apply_update=True, hyper_dev=None):
        """One sess.run of train.py:214-226 as a pure stream of kernel launches (no host synchronisation): this is also
        exactly what enable_cuda_graphs() captures."""
        a = self.args
        train_disc = kind == 'disc'
        gen, disc = self.generator, self.discriminator
        bs = self.bs_local
        if train_disc:
            with torch.no_grad():
                if a.train_disc_against_ema:
                    self.sync_ema()
                x_gen = gen(ema=self.ema, u=u, **self.model_opts) if a.train_disc_against_ema else gen(u=u, **self.model_opts)
            feats = disc(torch.cat([x_gen, x_real], 0), **self.model_opts)                       # fake rows, then real rows
            f_gen, f_dat = feats[:bs], feats[bs:]
        else:
            x_gen = gen(u=u, **self.model_opts)
            with torch.no_grad():
                f_dat = disc(x_real, **self.model_opts)
            with nn.frozen_params():                      # tf.gradients(xs=gen_params): no critic filter gradients
                f_gen = disc(x_gen, **self.model_opts)
        A, B = self._gather_features(f_gen, f_dat)
        lo, hi = local_rows(self.rank, bs)
        Ga, Gb, stats = self._match(A.detach(), B.detach(), rows=(lo, hi) if (self.world > 1 and not self.match_all_rows) else None)
        ga, gb = Ga[lo:hi], Gb[lo:hi]                                                            # this rank's towers
        self.last_grad_ys = (Ga, Gb)
        if train_disc:
            grad = self.sync['disc'].backward([feats], [torch.cat([ga, gb], 0)])                 # :122-128 + :134-139 (sum, not mean)
            if apply_update:
                self.disc_optimizer.run(grad, lr=-a.learning_rate_disc, hyper_dev=hyper_dev)     # :143,215
                if self.sharded_opt:
                    self.sync['disc'].gather_params(self.rank)
                if a.optimizer == 'adam':
                    disc.store.refresh_weight_cache()      # W of the updated critic, reused by the generator steps that follow
        else:
            grad = self.sync['gen'].backward([f_gen], [ga])                                      # :111-112 + :134-139
            if apply_update:
                self.gen_optimizer.run(grad, lr=a.learning_rate_gen, hyper_dev=hyper_dev)        # :142,222 (+ EMA :223)
                if self.sharded_opt:
                    self.sync['gen'].gather_params(self.rank)
        if self.sharded_opt and not apply_update:          # parity checks read the FULL summed gradient
            full = torch.empty_like(self.sync[kind].flat_grad)
            dist.all_gather_into_tensor(full, grad)
            grad = full
        self.last_grad = grad
        return stats

    def sync_ema(self):
        """With the sharded optimiser every rank updates only its slice of the generator's EMA shadow (train.py:63-64, 223): gather
        the slices before the shadow is read (sampling from the EMA generator, --train_disc_against_ema, checkpoints)."""
        if self.sharded_opt:
            sh = self.ema.shadow
            n = sh.numel() // self.world
            dist.all_gather_into_tensor(sh, sh[self.rank * n:(self.rank + 1) * n].clone())

    # ---- CUDA graphs: one captured critic step and one captured generator step, replayed with refreshed inputs ----------
    def enable_cuda_graphs(self, warmup=3):
        """Capture the critic step and the generator step (forward, all-gather, matching, backward, all-reduce, Adam+EMA:
        ~150 kernel launches each) into two CUDA graphs.  A step then costs three small host-to-device copies (images,
        latents, [lr, d1, d2]) and one graph launch, which removes the Python / launch latency that dominates once the
        per-rank batch is small (8 GPUs: 32 images per rank).  Training state is snapshotted around the warm-up runs, so
        enabling graphs does not change the trajectory.  Only the adam optimiser has a replayable update."""
        if self.graphs is not None:
            return
        a = self.args
        assert a.optimizer == 'adam', "CUDA-graph replay needs the fused Adam kernel (device-resident step scalars)"
        dev = self.device
        bs = self.bs_local
        self.g_x = torch.zeros((bs, self.image_size, self.image_size, 3), device=dev)
        if a.model == 'densenet':                          # models/densenet.py:53-56: four noise inputs
            self.g_u = [torch.zeros((bs, 100), device=dev)] + [torch.zeros((bs, s, s, 16), device=dev) for s in (8, 16, 32)]
        else:
            self.g_u = torch.zeros((bs, 100), device=dev)
        self.g_hyper = {k: torch.zeros(3, device=dev) for k in ('disc', 'gen')}
        opts = {'disc': self.disc_optimizer, 'gen': self.gen_optimizer}
        # snapshot everything a training step mutates
        snap = [t.detach().clone() for t in self._mutable_state()]
        snap_t = {k: o.state["t"] for k, o in opts.items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in ('disc', 'gen'):
                lr = -a.learning_rate_disc if k == 'disc' else a.learning_rate_gen
                self.g_hyper[k].copy_(torch.tensor([lr, 0.5, 0.001]))
            for _ in range(warmup):                       # allocator / workspace / cudaFuncSetAttribute warm-up, eager
                for k in ('disc', 'gen'):
                    self._step_body(k, self.g_x, self.g_u, True, hyper_dev=self.g_hyper[k])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        if self.world > 1:
            dist.barrier()
        from . import _lib
        graphs, outs, self.g_launches = {}, {}, {}
        pool = None
        for k in ('disc', 'gen'):
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                outs[k] = self._step_body(k, self.g_x, self.g_u, True, hyper_dev=self.g_hyper[k])
            self.g_launches[k] = _lib.launch_count() - n0      # this library's kernels inside one replay of the graph
            pool = g.pool()
            graphs[k] = g
        torch.cuda.synchronize(dev)
        with torch.no_grad():                             # restore the pre-warm-up training state
            for t, s0 in zip(self._mutable_state(), snap):
                t.copy_(s0)
        for k, o in opts.items():
            o.state["t"] = snap_t[k]
        self.discriminator.store.refresh_weight_cache()   # ... including the cached critic weights (same buffers)
        self.graphs, self.g_stats = graphs, outs

    def _mutable_state(self):
        out = [self.generator.flat, self.discriminator.flat, self.ema.shadow]
        for o in (self.gen_optimizer, self.disc_optimizer):
            out += [o.state["mg"]] + ([o.state["v"]] if o.state.get("v") is not None else [])
        return out

    def _replay(self, kind, x_real, u):
        a = self.args
        opt = self.disc_optimizer if kind == 'disc' else self.gen_optimizer
        lr, d1, d2 = opt.hyper(-a.learning_rate_disc if kind == 'disc' else a.learning_rate_gen)
        # the step scalars travel as kernel ARGUMENTS of three fill launches (an async copy from a reused pinned buffer
        # would race with the next step's host write)
        hd = self.g_hyper[kind]
        hd[0:1].fill_(lr); hd[1:2].fill_(d1); hd[2:3].fill_(d2)
        self.g_x.copy_(x_real, non_blocking=True)
        bufs = self.g_u if isinstance(self.g_u, list) else [self.g_u]
        if u is None:
            for b in bufs:
                b.uniform_(-1.0, 1.0)                                                            # tf.random_uniform  dcgan.py:30
        else:
            for b, src in zip(bufs, u if isinstance(u, (list, tuple)) else [u]):
                b.copy_(src, non_blocking=True)
        self.graphs[kind].replay()
        self.replayed_launches += self.g_launches[kind]
        self.step_counter += 1
        return self.g_stats[kind]

    def save(self, path):
        torch.save({'discriminator': {n: p.detach().cpu() for n, p in self.discriminator.named_parameters()},
                    'generator': {n: p.detach().cpu() for n, p in self.generator.named_parameters()}}, path)

    def load(self, path):
        """saver.restore (train.py:190-193): V, g, b of both networks by TensorFlow variable name -- a Trainer.save() file,
        an .npz keyed by variable name, or (where TensorFlow is importable) a reference checkpoint prefix.  Like the
        reference, Adam moments, Adam's t and the EMA shadow restart."""
        from .utils import checkpoint
        checkpoint.assign((self.discriminator, self.generator), checkpoint.load_variables(path))
        self.ema.attach(self.generator)

    def sample_tiles(self, path, path_ema=None, n=100):
        """train.py:233-243: PNG tile of generated samples (and of the EMA generator's samples)."""
        from .utils import plotting
        self.sync_ema()
        with torch.no_grad():
            x = self.generator(**self.model_opts)
            plotting.save_tile_img(plotting.img_tile(x[:n].cpu().numpy(), aspect_ratio=1.0, border_color=1.0, stretch=False), path)
            if path_ema is not None:
                xe = self.generator(ema=self.ema, **self.model_opts)
                plotting.save_tile_img(plotting.img_tile(xe[:n].cpu().numpy(), aspect_ratio=1.0, border_color=1.0, stretch=False), path_ema)


class GradSync:
    """Backward pass of one network with the tower-gradient SUM (train.py:134-139) overlapped with it.

    The gradient is taken with respect to the per-variable views of the flat parameter buffer (no final concatenation); a hook
    on every view copies its gradient into a persistent flat gradient buffer, and as soon as all variables of a LAYER have
    arrived the layer's slice is all-reduced on a side stream while the rest of the backward pass keeps running on the main
    stream (layers finish last-to-first; each slice starts travelling as soon as its layer is done).  Works the same
    eagerly and under CUDA-graph capture (the side stream forks from and joins into the capturing stream)."""

    def __init__(self, template, world, overlap=None):
        self.tpl, self.world = template, world
        # OTGAN_GRAD_SYNC = sharded (default) | single | overlap: reduce-scatter + sharded Adam + all-gather of the parameters
        # (ZeRO-1), ONE all-reduce of the flat gradient after the backward pass, or per-layer all-reduces on a side stream while it
        # runs (A/B switches; measurements in DESIGN.md section 6)
        mode = os.environ.get("OTGAN_GRAD_SYNC", "sharded")
        self.overlap = (mode == "overlap") if overlap is None else bool(overlap)
        self.sharded = (mode == "sharded") and world > 1 and not self.overlap
        self.grad_shard = None
        st = template.store
        self.flat_grad = torch.zeros_like(st.flat.detach())
        self.side = torch.cuda.Stream(device=st.flat.device) if world > 1 else None
        self.layers = {}                                   # scope -> [lo, hi, n_vars]
        self.var_layer = []
        for name, shape, off, n in st.specs:
            scope = name.rsplit("/", 1)[0]
            ent = self.layers.setdefault(scope, [off, off + n, 0])
            ent[0], ent[1], ent[2] = min(ent[0], off), max(ent[1], off + n), ent[2] + 1
            self.var_layer.append(scope)

    def backward(self, outputs, grad_outputs):
        """Returns the flat gradient buffer holding d(outputs)/d(parameters) (summed over ranks when world > 1)."""
        st = self.tpl.store
        views = st.last_views
        assert views is not None and len(views) == len(st.specs), "GradSync.backward needs the views of the network's last call"
        pending = {k: v[2] for k, v in self.layers.items()}
        done_events = []
        main = torch.cuda.current_stream()
        handles = []

        def on_grad(i, g):
            _, _, off, n = st.specs[i]
            dst = self.flat_grad[off:off + n]
            if g.data_ptr() != dst.data_ptr():
                dst.copy_(g.reshape(-1))
            scope = self.var_layer[i]
            pending[scope] -= 1
            if pending[scope] == 0 and self.world > 1 and self.overlap:
                lo, hi, _ = self.layers[scope]
                self.side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.side):
                    dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM)
            return None

        for i, v in enumerate(views):
            handles.append(v.register_hook(lambda g, i=i: on_grad(i, g)))
        try:
            grads = torch.autograd.grad(outputs, list(views), grad_outputs=grad_outputs, allow_unused=True)
        finally:
            for h in handles:
                h.remove()
        missing = [k for k, c in pending.items() if c > 0]
        if missing:                                        # variables the outputs do not depend on: zero gradient, reduce now
            for i, g in enumerate(grads):
                if g is None:
                    _, _, off, n = st.specs[i]
                    self.flat_grad[off:off + n].zero_()
            if self.world > 1 and self.overlap:
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    for k in missing:
                        lo, hi, _ = self.layers[k]
                        dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM)
        if self.world > 1:
            if self.overlap:
                main.wait_stream(self.side)
            elif self.sharded:                             # ZeRO-1: every rank receives the SUM of its own 1/world slice only
                n = self.flat_grad.numel() // self.world
                if self.grad_shard is None:
                    self.grad_shard = torch.empty(n, device=self.flat_grad.device, dtype=torch.float32)
                dist.reduce_scatter_tensor(self.grad_shard, self.flat_grad, op=dist.ReduceOp.SUM)
                del done_events
                return self.grad_shard
            else:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        del done_events
        return self.flat_grad

    def gather_params(self, rank):
        """After the sharded optimiser step: all-gather the updated parameter slices (in place in the flat buffer)."""
        flat = self.tpl.store.flat.detach()
        n = flat.numel() // self.world
        dist.all_gather_into_tensor(flat, flat[rank * n:(rank + 1) * n].clone())


def gather_features(f_gen, f_dat, world):
    """The one collective of the matching path: all-gather of every rank's [2, bs_local, D] feature slab (the
    reference's tf.concat over towers, utils/matching.py:16-19).  Returns (A, B) = fake / real features of ALL towers,
    [world * bs_local, D] each, rank-major (rank r hosts towers r*towers_local .. (r+1)*towers_local - 1)."""
    if world == 1:
        return f_gen, f_dat
    local = torch.stack([f_gen.detach(), f_dat.detach()], 0).contiguous()
    bs, D = local.shape[1], local.shape[2]
    buf = torch.empty((world * 2, bs, D), device=local.device, dtype=local.dtype)      # concatenated along dim 0
    dist.all_gather_into_tensor(buf, local)
    buf = buf.view(world, 2, bs, D)
    return buf[:, 0].reshape(-1, D), buf[:, 1].reshape(-1, D)


def local_rows(rank, bs_local):
    """Row range of this rank's towers inside the gathered [N, D] feature / grad_ys matrices."""
    return rank * bs_local, (rank + 1) * bs_local


def parity_check(world, rank, device, n_total=64, backends=(("cudnn", 1e-3), ("tcgen05", 5e-3))):
    """Multi-rank parity of one critic step and one generator step (used by tests/mgpu_parity.py and `bench.py --gpus N`):
    the summed tower gradient, distance and entropy computed by `world` ranks (all-gather + own-row backward + all-reduce)
    against the same step computed by ONE rank on the full batch with identical images, latents and parameters, plus a
    BITWISE check that every rank derives identical grad_ys / distance / entropy from the gathered embeddings (the
    replicated cost + Sinkhorn partition is deterministic).  Not bitwise on the gradients: per-rank batch sizes change the
    tile / split-K configuration of the convolution kernels, i.e. the fp32 summation order, so the gate is relative
    (1e-3 strict-fp32 library rung, 5e-3 TF32 tensor-core rung) at lambda = 10 where Sinkhorn does not amplify that noise.
    Collective on every rank; returns {"ok": bool, "rows": [...], "matching_bitwise": bool}."""
    prev_tf32, prev_backend = torch.backends.cudnn.allow_tf32, nn.CONV_BACKEND
    torch.backends.cudnn.allow_tf32 = False
    towers = 2 * world
    argv = ["--synthetic", "--nr_gpu", str(towers), "--batch_size", str(n_total // towers), "--nr_sinkhorn_iter", "50",
            "--sinkhorn_lambda", "10"]
    g = torch.Generator().manual_seed(1234)
    x_all = (torch.rand((n_total, 32, 32, 3), generator=g) * 2 - 1).to(device)
    u_all = (torch.rand((n_total, 100), generator=g) * 2 - 1).to(device)
    ok, rows, bitwise = True, [], True
    try:
        for backend, gate in backends:
            nn.CONV_BACKEND = backend
            for step_kind in ("disc", "gen"):
                res = {}
                for mode in ("multi", "single"):
                    w, r = (world, rank) if mode == "multi" else (1, 0)
                    tr = Trainer(build_parser().parse_args(argv), device, r, w)                  # same seed -> identical parameters
                    tr.step_counter = 0 if step_kind == "disc" else 1
                    tr.match_all_rows = True
                    bs = tr.bs_local
                    lo = r * bs
                    kind, stats = tr.step(x_all[lo:lo + bs], u=u_all[lo:lo + bs], apply_update=False)
                    assert kind == step_kind
                    res[mode] = (tr.last_grad.clone(), stats.clone(), tr.last_grad_ys)
                gm, sm, ys = res["multi"]
                gs, ss, _ = res["single"]
                if world > 1:                                   # every rank must hold the same bits of grad_ys / stats
                    for t in (ys[0], ys[1], sm):
                        hi, lo_ = t.clone(), t.clone()
                        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
                        bitwise = bitwise and bool(torch.equal(hi, lo_))
                rel = float((gm - gs).abs().max() / gs.abs().max())
                dd, de = abs(float(sm[0] - ss[0])), abs(float(sm[1] - ss[1]))
                rows.append({"backend": backend, "step": step_kind, "grad_rel_err": rel, "d_distance": dd, "d_entropy": de, "gate": gate})
                ok = ok and rel < gate and dd < 1e-6 and de < 1e-5
    finally:
        nn.CONV_BACKEND, torch.backends.cudnn.allow_tf32 = prev_backend, prev_tf32
    # partition "S": row-sharded cost blocks + own-row feature gradients at h = 256 against the replicated computation
    sharded = None
    if world > 1 and world % 2 == 0 and 256 % (world // 2) == 0:
        h, D = 256, 2048
        g2 = torch.Generator().manual_seed(77)
        A = torch.nn.functional.normalize(torch.rand((2 * h, D), generator=g2), dim=1).to(device)
        B = torch.nn.functional.normalize(torch.rand((2 * h, D), generator=g2), dim=1).to(device)
        bs = 2 * h // world
        lo, hi = rank * bs, (rank + 1) * bs
        fa, fb = list(torch.chunk(A, 2 * world, 0)), list(torch.chunk(B, 2 * world, 0))
        # lambda = 10 like the gradient parity above: the slabs and the full blocks differ by their split-K summation order (1e-7),
        # which lambda amplifies inside Sinkhorn (measured 9.6e-6 on grad_ys at lambda = 100, 8 ranks); an indexing error is O(1)
        ga_s, gb_s, st_s = matching.matching_step(fa, fb, 10.0, 20, rows=(lo, hi), shard=(rank, world))
        ga_r, gb_r, st_r = matching.matching_step(fa, fb, 10.0, 20)
        Ga_s, Gb_s, Ga_r, Gb_r = torch.cat(ga_s, 0), torch.cat(gb_s, 0), torch.cat(ga_r, 0), torch.cat(gb_r, 0)
        scale = float(Ga_r.abs().max())
        err = max(float((Ga_s[lo:hi] - Ga_r[lo:hi]).abs().max()), float((Gb_s[lo:hi] - Gb_r[lo:hi]).abs().max())) / scale
        dstat = float((st_s - st_r).abs().max())
        hi_t, lo_t = st_s.clone(), st_s.clone()                 # the all-gathered blocks give every rank the same bits
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(hi_t, lo_t))
        sharded = {"h": h, "D": D, "own_rows_grad_rel_err_vs_replicated": err, "d_stats_vs_replicated": dstat, "stats_bitwise_across_ranks": same}
        ok = ok and err < 1e-5 and dstat < 1e-6 and same
    flag = torch.tensor([1.0 if (ok and bitwise) else 0.0], device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(flag.item() == 1.0), "world": world, "rows": rows, "matching_bitwise": bitwise, "sharded_cost_blocks": sharded}


def maybe_flip(x, rng):
    """train.py:163-170: per-image horizontal flip with probability 0.5 (vectorised)."""
    flip = rng.rand(x.shape[0]) < 0.5
    out = x.copy()
    out[flip] = x[flip][:, :, ::-1, :]
    return out


def main(argv=None):
    args = build_parser().parse_args(argv)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('otgan_b200.train needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    if rank == 0:
        print(args)
    trainer = Trainer(args, device, rank, world)
    if rank == 0:
        print('model has a hidden representation with %d features' % trainer.num_features)        # train.py:56
    rng = np.random.RandomState(args.seed + rank)
    if args.synthetic:
        trainx = cifar10_data.synthetic(max(args.nr_gpu * args.batch_size * 4, 2048), seed=args.seed, size=args.image_size)
    else:
        trainx, _ = cifar10_data.load(args.data_dir + '/cifar-10-python')
        trainx = np.transpose(trainx, (0, 2, 3, 1)) / 127.5 - 1.                                  # :158
        trainx = trainx.astype(np.float32)
    nr_batches_train_per_gpu = trainx.shape[0] // (args.nr_gpu * args.batch_size)                # :159
    current_epoch = 0
    if args.load_params:
        trainer.load(os.path.join(args.save_dir, args.model_name))
        ix = args.model_name.rfind('-')                                                          # :191-193
        current_epoch = int(args.model_name[ix + 1:]) if ix >= 0 and args.model_name[ix + 1:].isdigit() else 0
    if args.cuda_graphs:
        trainer.enable_cuda_graphs()
    os.makedirs(args.save_dir, exist_ok=True) if rank == 0 and not args.synthetic else None
    if rank == 0:
        print('starting training')
    start_time = time.time()
    for epoch in range(current_epoch, 1000000):                                                  # :196
        begin = time.time()
        inds = np.random.RandomState(args.seed + epoch).permutation(trainx.shape[0])             # :200 (same on every rank)
        trainx = trainx[inds]
        dist_gen, dist_disc, entropy = [], [], []
        def batch(t):                                                                            # :209-211
            xs = []
            for i in range(trainer.towers_local):
                tower = rank * trainer.towers_local + i
                td = t + tower * nr_batches_train_per_gpu
                xs.append(maybe_flip(trainx[td * args.batch_size:(td + 1) * args.batch_size], rng))
            return torch.from_numpy(np.concatenate(xs, 0)).pin_memory()

        def log(kind, s):                                                                        # the sess.run fetch
            (dist_disc if kind == 'disc' else dist_gen).append(s[0])
            entropy.append(s[1])

        # software pipeline: while step t computes, the images of step t+1 are uploaded and the statistics of step t-1 are read
        pending = None
        x = trainer.stage(batch(0)) if nr_batches_train_per_gpu > 0 else None
        for t in range(nr_batches_train_per_gpu):
            kind, stats = trainer.step(x)
            if t + 1 < nr_batches_train_per_gpu:
                x = trainer.stage(batch(t + 1))
            h = (kind, trainer.fetch_async(stats), trainer.step_counter)
            if pending is not None:
                log(pending[0], pending[1].result())
                if rank == 0 and args.log_every and pending[2] % args.log_every == 0:
                    print('step %d (%s): distance %.6f entropy %.6f' % (pending[2], pending[0], dist_disc[-1] if pending[0] == 'disc' else dist_gen[-1], entropy[-1]))
            pending = h
            if args.max_steps and trainer.step_counter >= args.max_steps:
                break
        if pending is not None:
            log(pending[0], pending[1].result())
            if rank == 0 and args.log_every and pending[2] % args.log_every == 0:
                print('step %d (%s): distance %.6f entropy %.6f' % (pending[2], pending[0], dist_disc[-1] if pending[0] == 'disc' else dist_gen[-1], entropy[-1]))
        if rank == 0:                                                                            # :231
            print("Iteration %d, time = %ds, train distance before gen = %.6f, train distance before disc = %.6f, avg matching entropy = %.6f"
                  % (epoch, time.time() - begin, np.mean(dist_gen) if dist_gen else float('nan'),
                     np.mean(dist_disc) if dist_disc else float('nan'), np.mean(entropy)))
            sys.stdout.flush()
        if rank == 0 and not args.synthetic:                                                     # :233-243
            trainer.sample_tiles(os.path.join(args.save_dir, 'sample%d.png' % epoch), os.path.join(args.save_dir, 'ema_sample%d.png' % epoch))
        if (epoch + 1) % 200 == 0 and epoch != current_epoch and rank == 0 and not args.synthetic:   # :275-277
            trainer.save(os.path.join(args.save_dir, 'med_gan_params-%d' % epoch))
        if args.max_steps and trainer.step_counter >= args.max_steps:
            break
    if rank == 0:
        print('total updates %d, elapsed %.3f s' % (trainer.step_counter, time.time() - start_time))
    if world > 1:
        if trainer.graphs is not None:                      # see bench.py: no collective destructor under captured NCCL graphs
            trainer.graphs = trainer.g_stats = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
