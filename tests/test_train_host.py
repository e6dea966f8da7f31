"""CPU tests of the train-loop host logic: flag surface identical to the reference, step schedule, and the N>1
sharding (all-gather ordering + own-row slicing) on a world_size-2 gloo group."""
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist

from otgan_b200 import train as T

REF_FLAGS = ['--seed', '--batch_size', '--learning_rate_disc', '--learning_rate_gen', '--data_dir', '--save_dir',
             '--optimizer', '--nonlinearity', '--nr_gpu', '--nr_gen_per_disc', '--sinkhorn_lambda', '--nr_sinkhorn_iter',
             '--single_batch', '--train_disc_against_ema', '--model', '--load_params', '--model_name', '--no_sinkhorn']


def test_flags_and_defaults_match_reference():
    """train.py:14-33"""
    p = T.build_parser()
    opts = {s for a in p._actions for s in a.option_strings}
    assert all(f in opts for f in REF_FLAGS)
    a = p.parse_args([])
    assert (a.seed, a.batch_size, a.learning_rate_disc, a.learning_rate_gen, a.optimizer, a.nonlinearity, a.nr_gpu,
            a.nr_gen_per_disc, a.sinkhorn_lambda, a.nr_sinkhorn_iter, a.model) == \
           (1, 625, 0.0003, 0.0003, 'adam', 'crelu', 8, 5, 500., 500, 'dcgan')
    assert not (a.single_batch or a.train_disc_against_ema or a.load_params or a.no_sinkhorn)


def test_step_schedule():
    """critic step when step_counter % (nr_gen_per_disc + 1) == 0 (train.py:214-226): D G G G G G D ..."""
    kinds = ['disc' if s % 6 == 0 else 'gen' for s in range(13)]
    assert kinds == ['disc'] + ['gen'] * 5 + ['disc'] + ['gen'] * 5 + ['disc']


def test_maybe_flip():
    x = np.arange(2 * 2 * 3 * 1, dtype=np.float32).reshape(2, 2, 3, 1)
    out = T.maybe_flip(x, np.random.RandomState(0))
    for i in range(2):
        assert np.array_equal(out[i], x[i]) or np.array_equal(out[i], x[i][:, ::-1, :])


def test_two_rank_gather_and_grad_sum_gloo():
    """world_size-2 gloo group on CPU: all-gather ordering (rank-major towers), own-row slicing, gradient SUM."""
    import subprocess
    import sys
    port = 29500 + os.getpid() % 2000
    worker = os.path.join(os.path.dirname(__file__), "_gloo_worker.py")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, worker], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "RANK 0 OK" in outs[0] and "RANK 1 OK" in outs[1]
