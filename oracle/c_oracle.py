"""ctypes binding of oracle/libotgan_oracle.so (C + OpenMP restatement) -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this (see
oracle/matching_oracle.c for the reference lines it follows).  PARITY UNPINNED (no reference tests exist).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libotgan_oracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle in place (called by __graft_entry__.build())."""
    src = os.path.join(_HERE, "matching_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.otgan_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().otgan_oracle_max_threads())


def set_threads(n):
    lib().otgan_oracle_set_threads(int(n))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def two_batch(A, B, lam, T, cost_kind=0, dtype=np.float32, want_plans=True):
    """Whole two-batch matching call on concatenated features A=[A1;A2], B=[B1;B2] ([N,D] each).

    Returns dict(f_aa, f_bb, f_ab, f_ba, P[6,h,h], entropy, dist, phase_ms[cost, sinkhorn, matched])."""
    real = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    fn = lib().otgan_oracle_two_batch_f32 if dtype == np.float32 else lib().otgan_oracle_two_batch_f64
    A = np.ascontiguousarray(A, dtype=dtype)
    B = np.ascontiguousarray(B, dtype=dtype)
    N, D = A.shape
    h = N // 2
    out = {k: np.empty((N, D), dtype=dtype) for k in ("f_aa", "f_bb", "f_ab", "f_ba")}
    P = np.empty((6, h, h), dtype=dtype) if want_plans else None
    ent, dist = real(0), real(0)
    phase = (ctypes.c_double * 3)()
    fn(_p(A), _p(B), ctypes.c_int(N), ctypes.c_int(D), ctypes.c_int(cost_kind), real(lam), ctypes.c_int(T),
       _p(out["f_aa"]), _p(out["f_bb"]), _p(out["f_ab"]), _p(out["f_ba"]), _p(P),
       ctypes.byref(ent), ctypes.byref(dist), phase)
    out.update(P=P, entropy=ent.value, dist=dist.value, phase_ms=list(phase))
    return out


def single_batch(A, B, lam, T, cost_kind=0, dtype=np.float32):
    real = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    fn = lib().otgan_oracle_single_batch_f32 if dtype == np.float32 else lib().otgan_oracle_single_batch_f64
    A = np.ascontiguousarray(A, dtype=dtype)
    B = np.ascontiguousarray(B, dtype=dtype)
    N, D = A.shape
    out = {k: np.empty((N, D), dtype=dtype) for k in ("f_aa", "f_bb", "f_ab", "f_ba")}
    P = np.empty((3, N, N), dtype=dtype)
    ent = real(0)
    fn(_p(A), _p(B), ctypes.c_int(N), ctypes.c_int(D), ctypes.c_int(cost_kind), real(lam), ctypes.c_int(T),
       _p(out["f_aa"]), _p(out["f_bb"]), _p(out["f_ab"]), _p(out["f_ba"]), _p(P), ctypes.byref(ent))
    out.update(P=P, entropy=ent.value)
    return out


def sinkhorn(C, lam, T, dtype=np.float32):
    """C: [nblk, m, n] cost blocks -> (P, entropy[nblk], pc[nblk])."""
    real = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    fn = lib().otgan_oracle_sinkhorn_f32 if dtype == np.float32 else lib().otgan_oracle_sinkhorn_f64
    C = np.ascontiguousarray(C, dtype=dtype)
    nblk, m, n = C.shape
    P = np.empty_like(C)
    ent = np.empty(nblk, dtype=dtype)
    pc = np.empty(nblk, dtype=dtype)
    fn(_p(C), ctypes.c_int(nblk), ctypes.c_int(m), ctypes.c_int(n), real(lam), ctypes.c_int(T), _p(P), _p(ent), _p(pc))
    return P, ent, pc
