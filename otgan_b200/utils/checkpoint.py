"""Parameter import / export under the reference's TensorFlow variable names.

The reference saves `tf.train.Saver(all_params)` checkpoints (train.py:60,276: only V, g, b of every layer; Adam moments
and EMA shadows are NOT saved, train.py:190-193 restarts them) named `med_gan_params-<epoch>`.  The variables here carry the
same names and layouts (`discriminator/conv2d_1/V` is the HWIO kernel, `generator/dense_0/V` is [in, out]), so a
checkpoint maps one-to-one:

    load_variables(path) -> {name: np.ndarray}      the file FORMAT is sniffed, not the extension:
        numpy archive         keyed by variable name (what `np.savez(**{v.op.name: sess.run(v)})` writes on the TensorFlow
                              side -- the portable exchange format; a trailing ":0" on the keys is accepted)
        torch archive         Trainer.save() output (the reference's extensionless name `med_gan_params-<epoch>` included)
        <prefix>.index exists a TensorFlow checkpoint prefix, read with tf.train.load_checkpoint when TensorFlow is importable
                              (it is not in this image: a clear error is raised instead)
    assign(templates, variables)   copy into the flat parameter buffers (shape-checked), invalidating cached weights
    export_npz(templates, path)    the inverse, for loading these parameters back into the reference
"""
import numpy as np
import torch


def _sniff(path):
    """'npz' | 'torch' | 'tf' from the file contents (both numpy and torch archives are zip files)."""
    import os
    import zipfile
    if os.path.isfile(path) and zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            names = z.namelist()
        return "torch" if any(n.endswith("data.pkl") for n in names) else "npz"
    if os.path.isfile(path):
        return "torch"                              # legacy (non-zip) torch pickle
    for ext in (".pt", ".npz"):
        if os.path.isfile(path + ext):
            return _sniff(path + ext) + ":" + ext
    if os.path.isfile(path + ".index"):
        return "tf"
    raise FileNotFoundError("no checkpoint at %r (tried the path itself, +.pt, +.npz and the TensorFlow prefix +.index)" % path)


def load_variables(path):
    kind = _sniff(path)
    if ":" in kind:
        kind, ext = kind.split(":")
        path = path + ext
    if kind == "npz":
        with np.load(path) as z:
            return {k[:-2] if k.endswith(":0") else k: np.asarray(z[k]) for k in z.files}
    if kind == "torch":
        ck = torch.load(path, map_location="cpu")
        return {n: t.numpy() for group in ck.values() for n, t in group.items()}
    try:
        import tensorflow as tf                     # noqa: F401  (absent here; present where the reference runs)
    except ImportError as e:
        raise ImportError("reading the TensorFlow checkpoint %r needs tensorflow; export it on the TensorFlow side with "
                          "np.savez(path, **{v.op.name: sess.run(v) for v in tf.trainable_variables()}) and load the .npz"
                          % path) from e
    reader = tf.train.load_checkpoint(path)
    return {n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()}


def assign(templates, variables, strict=True):
    """templates: iterable of nn.Template (discriminator, generator).  Returns the list of variable names assigned."""
    done = []
    with torch.no_grad():
        for tpl in templates:
            for name, p in tpl.named_parameters():
                if name not in variables:
                    if strict:
                        raise KeyError("checkpoint has no variable %r" % name)
                    continue
                v = np.asarray(variables[name], dtype=np.float32)
                if tuple(v.shape) != tuple(p.shape):
                    raise ValueError("variable %r: checkpoint shape %s, model shape %s" % (name, v.shape, tuple(p.shape)))
                p.copy_(torch.from_numpy(v))
                done.append(name)
            tpl.store.version += 1                  # W = g V/||V|| caches are stale
    return done


def export_npz(templates, path):
    arrays = {}
    for tpl in templates:
        for name, p in tpl.named_parameters():
            arrays[name] = p.detach().cpu().numpy()
    np.savez(path, **arrays)
    return sorted(arrays)
