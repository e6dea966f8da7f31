"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN FILES (imported unmodified from /root/reference)
over tests/golden/tf1_numpy_shim.py, a float64 numpy emulation of the TensorFlow-1.x primitives they call.

    python tests/golden/make_reference_golden.py        (needs /root/reference; the fixtures travel, this script's input does not)

What executes from the reference: utils/matching.py (get_matched_features, get_matched_features_single_batch,
get_matched_features_random, calc_distance), toy_example/matching_cpu.py (the same four on the tensor API with the
Euclidean cost), utils/nn.py (get_params, apply_pre_activation, conv2d, dense) and models/dcgan.py, models/densenet.py
(disc_spec, gen_spec through tf.make_template).  TensorFlow itself is not installable here (python 3.12, no wheel,
no network): the primitives (matmul, logsumexp, softmax, conv2d, ...) are the shim's, with the TF-1.x semantics documented
there.  The fixtures therefore pin the oracle / CUDA path to the reference's CODE -- block order, transposes, regrouping,
variable names and creation order, layer sequence, SAME padding, resize, CReLU list interleave -- with the primitive
arithmetic restated.

Model weights are not stored (34 M parameters): both this script and the tests derive every variable from its NAME and
shape (`seeded_variable`), so only inputs and outputs are committed.
"""
import importlib.util
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OTGAN_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import tf1_numpy_shim as shim  # noqa: E402


def seeded_variable(name, shape):
    """Deterministic value of a model variable from its TensorFlow name: V ~ N(0, 0.05), g ~ U(0.5, 1.5), b ~ N(0, 0.1)."""
    rng = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    leaf = name.rsplit("/", 1)[-1]
    if leaf == "V":
        return rng.normal(0.0, 0.05, size=shape).astype(np.float32)
    if leaf == "g":
        return rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
    return rng.normal(0.0, 0.1, size=shape).astype(np.float32)


def synth(n, d, seed, sigma=0.3):
    """Clustered, non-negative, row-normalised embeddings (same generator as oracle.synth_embeddings 'clustered')."""
    rng_c = np.random.RandomState(999)
    cent = rng_c.randn(10, d // 2)
    rng = np.random.RandomState(seed)
    x = cent[rng.randint(0, 10, size=n)] + sigma * rng.randn(n, d // 2)
    f = np.concatenate([np.maximum(x, 0), np.maximum(-x, 0)], axis=1)
    return (f / np.sqrt(np.sum(f * f, axis=1, keepdims=True))).astype(np.float32)


def load_reference():
    shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    import types
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = utils_pkg
    nn = load("utils.nn", "utils/nn.py")
    utils_pkg.nn = nn
    matching = load("utils.matching", "utils/matching.py")
    toy = load("ref_toy_matching_cpu", "toy_example/matching_cpu.py")
    dcgan = load("ref_models_dcgan", "models/dcgan.py")
    densenet = load("ref_models_densenet", "models/densenet.py")
    return matching, toy, dcgan, densenet


def towers(A, G):
    return [shim._t(a) for a in np.split(A.astype(np.float64), G)]


def cat(x):
    return np.concatenate([np.asarray(v) for v in x], axis=0)


def matching_cases(matching, toy):
    out = {}
    for name, (kind, N, D, G, lam, T, sigma) in {
        "ref_two_batch_g4": ("two", 32, 64, 4, 500.0, 50, 0.3),
        "ref_two_batch_g2_ragged": ("two", 12, 38, 2, 100.0, 7, 1.0),
        "ref_two_batch_g8": ("two", 64, 256, 8, 500.0, 100, 1.0),
        "ref_single_batch_g3": ("single", 24, 48, 3, 500.0, 15, 0.3),
        "ref_toy_two_batch": ("toy", 64, 16, 1, 50.0, 10, 0.0),
        "ref_toy_single_batch": ("toy_single", 32, 16, 1, 50.0, 10, 0.0),
    }.items():
        if kind.startswith("toy"):
            rng = np.random.RandomState(7)
            A = rng.randn(N, D).astype(np.float32)
            B = (rng.randn(N, D) * 0.5 + 1.0).astype(np.float32)
        else:
            A, B = synth(N, D, 1, sigma), synth(N, D, 2, sigma)
        rec = {"A": A, "B": B, "lam": lam, "T": T, "G": G, "kind": kind}
        if kind == "two":
            fa, fb = towers(A, G), towers(B, G)
            m = matching.get_matched_features(fa, fb, lam, T)
            d = matching.calc_distance(fa, fb, m)
            r = matching.get_matched_features_random(fa, fb)
            rec.update(rand_aa=cat(r[0]), rand_bb=cat(r[1]))
        elif kind == "single":
            fa, fb = towers(A, G), towers(B, G)
            m = matching.get_matched_features_single_batch(fa, fb, lam, T)
            d = matching.calc_distance(fa, fb, m)
        elif kind == "toy":
            a, b = shim._t(A), shim._t(B)
            m = toy.get_matched_features(a, b, lam, T)
            d = toy.calc_distance(a, b, m)
        else:
            a, b = shim._t(A), shim._t(B)
            m = toy.get_matched_features_single_batch([a], [b], lam, T, N)
            m = tuple(cat(x) if isinstance(x, list) else x for x in m)
            d = toy.calc_distance(a, b, m)
        c = (lambda x: np.asarray(x)) if kind.startswith("toy") else cat
        rec.update(f_aa=c(m[0]), f_bb=c(m[1]), f_ab=c(m[2]), f_ba=c(m[3]), entropy=float(m[4]), dist=float(d))
        out[name] = rec
    return out


def assign_seeded(prefix):
    for name, v in shim.variables():
        if name.startswith(prefix + "/"):
            shim.set_variable(name, seeded_variable(name, v.shape).astype(np.float64))


def model_cases(dcgan, densenet):
    out = {}
    rng = np.random.RandomState(2024)
    x = (rng.rand(2, 32, 32, 3) * 2 - 1).astype(np.float32)
    for mname, mod in (("dcgan", dcgan), ("densenet", densenet)):
        shim.set_seed(11)
        mod.discriminator(shim._t(np.zeros((2, 32, 32, 3)) + 0.1), init=True)          # train.py:52-54: creates the variables
        mod.generator(2, init=True)
        names = [(n, list(v.shape)) for n, v in shim.variables()]
        assign_seeded("discriminator")
        assign_seeded("generator")
        feats = np.asarray(mod.discriminator(shim._t(x)))
        # the generator draws its latents from tf.random_uniform: record the draws so the test can feed the same ones
        draws = []
        orig = shim.random_uniform

        def recording(shape, minval=0.0, maxval=1.0, dtype=None):
            u = (np.asarray(orig(shape, minval, maxval)).astype(np.float32)).astype(np.float64)   # fp32-representable latents
            draws.append(u)
            return shim._t(u)

        sys.modules["tensorflow"].random_uniform = recording
        try:
            img = np.asarray(mod.generator(2))
        finally:
            sys.modules["tensorflow"].random_uniform = orig
        rec = {"x": x, "features": feats, "image": img, "n_latents": len(draws),
               "var_names": np.array([n for n, _ in names]), "var_sizes": np.array([int(np.prod(s)) for _, s in names])}
        for i, u in enumerate(draws):
            rec["u%d" % i] = u.astype(np.float32)
        out["ref_model_" + mname] = rec
    return out


if __name__ == "__main__":
    matching, toy, dcgan, densenet = load_reference()
    cases = matching_cases(matching, toy)
    cases.update(model_cases(dcgan, densenet))
    for name, rec in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        extra = "dist=%.9g entropy=%.9g" % (rec["dist"], rec["entropy"]) if "dist" in rec else \
            "features %s |f|max %.4g, image %s" % (rec["features"].shape, np.abs(rec["features"]).max(), rec["image"].shape)
        print(name, extra)
