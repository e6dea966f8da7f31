"""Host mirror of the reference's ``utils/nn.py`` (layer library + optimisers) re-hosted on PyTorch for the B200 path.

Same public names and argument meaning as /root/reference/utils/nn.py:
    conv2d(x, num_filters, pre_activation='celu', filter_size=[3,3], stride=[1,1], pad='SAME', dilate=1, upsample=False,
           init_scale=1., counters={}, init=False, ema=None, weight_norm=True, use_b=True, use_g=True)      utils/nn.py:328-338
    dense(x, num_units, pre_activation='celu', init_scale=1., counters={}, init=False, ema=None, ...)      utils/nn.py:315-325
    get_params / apply_pre_activation / get_name                                                           utils/nn.py:95-206
    adam_updates / adamax_updates / nesterov_updates                                                       utils/nn.py:29-87
plus the two TensorFlow mechanisms the reference leans on, restated:
    make_template(name, spec)  -- tf.make_template: variables are created on the first call and shared afterwards
    arg_scope([...], **kw)     -- tf.contrib.framework.arg_scope: default keyword arguments for the listed layer functions
    ExponentialMovingAverage   -- tf.train.ExponentialMovingAverage (train.py:63-64)

Tensors are NHWC at this API exactly like the reference (`[B, H, W, C]`); kernels are stored HWIO (utils/nn.py:123) under
the TensorFlow variable names (`discriminator/conv2d_0/V`, `.../g`, `.../b`) so reference checkpoints map one-to-one.
TensorFlow semantics that differ from PyTorch defaults are restated explicitly: `SAME` padding (asymmetric for stride 2),
`tf.nn.l2_normalize` (x * rsqrt(max(sum x^2, 1e-12))), nearest-neighbour resize (src = dst // 2), CReLU over a LIST of
inputs interleaved per element ([x0, -x0, x1, -x1, ...], utils/nn.py:198-200).

On the GPU the layers run on this library's CUDA kernels through the C ABI (include/otgan.h): nn.conv2d on the tcgen05
implicit-GEMM convolutions (_ConvTC: fprop / dgrad / wgrad; _ConvUp2TC: the generator's resize + convolution pairs in the
fused sub-pixel form; _ConvNarrow: the two 3-channel layers), weight norm, CReLU, GLU, the critic head and Adam+EMA on
their fused kernels.  Shapes the convolution kernels do not tile (DenseNet's 16-filter layers, odd batch sizes) and the
100-wide dense layer go to cuDNN / cuBLAS through torch (CONV_BACKEND = "cudnn" forces that rung everywhere, for A/B
tests).  CPU tensors take the literal torch op sequence (host-side tests only; there is no CPU product path).
"""
import contextlib
import math
import threading

import torch
import torch.nn.functional as F

from .. import _lib

DATA_DEPENDENT_INIT = False   # the reference builds the init assigns but never runs them: default off

_tls = threading.local()


# ------------------------------------------------------------------------------------------------ arg_scope
def _scope_stack():
    if not hasattr(_tls, "arg_scopes"):
        _tls.arg_scopes = []
    return _tls.arg_scopes


@contextlib.contextmanager
def arg_scope(funcs, **kwargs):
    """tf.contrib.framework.arg_scope restated: default kwargs for the listed layer functions."""
    _scope_stack().append(({f.__name__ for f in funcs}, kwargs))
    try:
        yield
    finally:
        _scope_stack().pop()


def add_arg_scope(fn):
    def wrapped(*args, **kwargs):
        merged = {}
        for names, kw in _scope_stack():
            if fn.__name__ in names:
                merged.update(kw)
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


# ------------------------------------------------------------------------------------------------ variables / templates
class VariableStore:
    """All variables of one template, packed into ONE flat fp32 leaf so that gradients come out contiguous and the
    optimiser is a single fused kernel launch over the flat buffers (utils/nn.py:50-73 does ~10 ops per tensor)."""

    def __init__(self, name, device):
        self.name, self.device = name, device
        self.specs = []            # (var_name, shape, offset, numel) in creation order == tf.trainable_variables() order
        self.index = {}
        self.pending = {}          # values created during the init pass, before packing
        self.flat = None           # torch leaf [total], requires_grad
        self.frozen = False
        self.version = 0           # bumped by every optimiser update of this store's variables
        self.cache_version = -1    # version the weight cache below was computed from (-1: never)
        self.wcache = {}           # scope -> (wt [C,K], inv [C], wt_ihwo or None): W = g V/||V|| for gradient-free calls

    def create(self, var_name, shape, init_fn):
        if var_name in self.index:
            return
        if self.frozen:
            raise RuntimeError("variable %s requested after the template was built" % var_name)
        numel = int(math.prod(shape))
        offset = sum(s[3] for s in self.specs)
        self.specs.append((var_name, tuple(shape), offset, numel))
        self.index[var_name] = len(self.specs) - 1
        self.pending[var_name] = init_fn(shape).to(self.device, torch.float32)

    def pack(self):
        total = sum(s[3] for s in self.specs)
        pad = (-total) % 64                     # float4-sized for the fused optimiser kernel AND divisible into <= 16 equal float4 shards (sharded Adam)
        flat = torch.zeros(total + pad, device=self.device, dtype=torch.float32)
        for name, shape, off, n in self.specs:
            flat[off:off + n] = self.pending[name].reshape(-1)
        self.flat = flat.requires_grad_(True)
        self.pending = {}
        self.frozen = True

    def begin_call(self):
        """One autograd node for ALL variable views of a template call (split_with_sizes: its backward is a single
        concatenation into the flat gradient; per-variable slices would each materialise a full-size zero gradient).
        Under `frozen_params()` the views are detached: the call back-propagates to its INPUT only (the critic inside a
        generator step, train.py:111-112 -- tf.gradients(xs=gen_params) never builds the critic's filter gradients)."""
        if self.frozen:
            total = self.num_params()
            src = self.flat.detach() if getattr(_tls, "freeze_params", False) else self.flat
            self._views = torch.split_with_sizes(src[:total], [s[3] for s in self.specs])
            self._views_src = {}

    def get(self, var_name, source=None):
        """View of one variable; `source` substitutes another flat buffer (the EMA shadow, utils/nn.py:89-93)."""
        if not self.frozen:
            return self.pending[var_name]
        i = self.index[var_name]
        _, shape, off, n = self.specs[i]
        if source is None:
            if getattr(self, "_views", None) is None:          # outside a template call (inspection): plain slice
                return self.flat[off:off + n].view(shape)
            return self._views[i].view(shape)
        return source[off:off + n].view(shape)

    def refresh_weight_cache(self):
        """Recompute W = g * V / ||V|| of every weight-normalised layer (and the IHWO copy the dgrad kernel reads) into
        persistent buffers.  Calls that build no parameter gradients (torch.no_grad() / frozen_params(): the critic inside
        the five generator steps that follow a critic update, train.py:214-226) then reuse them instead of re-running the
        weight-norm and transpose kernels over all 34 M critic parameters twice per step.  The buffers keep their
        addresses, so a captured CUDA graph can both refresh and read them."""
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        with torch.no_grad():
            for name, shape, off, n in self.specs:
                if not name.endswith("/V"):
                    continue
                scope = name[:-2]
                if scope + "/g" not in self.index:
                    continue
                C = shape[-1]
                K = n // C
                V = self.flat[off:off + n]
                _, _, goff, gn = self.specs[self.index[scope + "/g"]]
                g = self.flat[goff:goff + gn]
                ent = self.wcache.get(scope)
                if ent is None:
                    wt = torch.empty((C, K), device=self.device, dtype=torch.float32)
                    inv = torch.empty((C,), device=self.device, dtype=torch.float32)
                    wt_t = None
                    if len(shape) == 4 and shape[2] % 32 == 0 and (C % 32 == 0):
                        wt_t = torch.empty((shape[2], shape[0] * shape[1] * C), device=self.device, dtype=torch.float32)
                    ent = self.wcache[scope] = (wt, inv, wt_t)
                wt, inv, wt_t = ent
                ws = _wn_workspace(self.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
                if wt_t is not None and K % 4 == 0 and C % 4 == 0:      # W and the IHWO dgrad operand from one pass over V
                    rc = lib.otgan_weightnorm_fwd2_f32(K, C, shape[0] * shape[1], V.data_ptr(), g.data_ptr(), wt.data_ptr(), wt_t.data_ptr(),
                                                       inv.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
                    _lib.check(rc, "otgan_weightnorm_fwd2_f32")
                else:
                    rc = lib.otgan_weightnorm_fwd_f32(K, C, V.data_ptr(), g.data_ptr(), wt.data_ptr(), inv.data_ptr(), ws.data_ptr(),
                                                      ws.numel() * 4, stream)
                    _lib.check(rc, "otgan_weightnorm_fwd_f32")
                    if wt_t is not None:
                        rc = lib.otgan_ohwi_to_ihwo_f32(C, shape[0] * shape[1], shape[2], wt.data_ptr(), wt_t.data_ptr(), stream)
                        _lib.check(rc, "otgan_ohwi_to_ihwo_f32")
        self.cache_version = self.version

    def cached_weight(self, scope):
        """(wt, wt_ihwo) if the cache is current and this call builds no parameter gradients, else None."""
        if self.cache_version != self.version or scope not in self.wcache:
            return None
        if torch.is_grad_enabled() and not getattr(_tls, "freeze_params", False):
            return None
        wt, _, wt_t = self.wcache[scope]
        return wt, wt_t

    def named_parameters(self):
        if self.frozen:
            return [(n, self.flat[off:off + k].view(shape)) for n, shape, off, k in self.specs]
        return [(n, self.pending[n]) for n, _, _, _ in self.specs]

    def num_params(self):
        return sum(s[3] for s in self.specs)


class Template:
    """tf.make_template(name, spec): first call creates the variables (init pass, train.py:52-54), later calls share them."""

    def __init__(self, name, spec):
        self.name, self.spec, self.store = name, spec, None

    def __call__(self, *args, **kwargs):
        device = kwargs.pop("device", None)
        if self.store is None:
            if device is None:
                device = next((a.device for a in args if isinstance(a, torch.Tensor)), None) or torch.device("cuda")
            self.store = VariableStore(self.name, torch.device(device))
        prev = getattr(_tls, "store", None)
        _tls.store = self.store
        self.store.begin_call()
        try:
            out = self.spec(*args, **kwargs)
        finally:
            _tls.store = prev
            self.store.last_views = getattr(self.store, "_views", None)     # kept for gradient hooks (train.GradSync)
            self.store._views = None
        if not self.store.frozen:
            self.store.pack()
        return out

    # convenience accessors used by the train loop / optimisers / checkpointing
    @property
    def flat(self):
        return self.store.flat

    def named_parameters(self):
        return self.store.named_parameters()

    def reset(self):
        self.store = None


def make_template(name, spec):
    return Template(name, spec)


@contextlib.contextmanager
def frozen_params():
    """Template calls inside this context treat their variables as constants (no parameter gradients are built)."""
    prev = getattr(_tls, "freeze_params", False)
    _tls.freeze_params = True
    try:
        yield
    finally:
        _tls.freeze_params = prev


class ExponentialMovingAverage:
    """tf.train.ExponentialMovingAverage(decay) over one template's flat variable buffer (train.py:63-64).
    `apply` is folded into adam_updates' fused kernel when the optimiser is given `ema=`; `average(store)` is what
    get_var_maybe_avg (utils/nn.py:89-93) reads."""

    def __init__(self, decay):
        self.decay = float(decay)
        self.shadow = None

    def attach(self, template):
        self.shadow = template.flat.detach().clone()      # TF initialises the shadow to the variable's value
        return self

    def apply(self, template):
        if self.shadow is None:
            self.attach(template)
        else:
            self.shadow.mul_(self.decay).add_(template.flat.detach(), alpha=1.0 - self.decay)


# ------------------------------------------------------------------------------------------------ helpers (utils/nn.py)
def int_shape(x):
    return list(x.shape)


def get_name(layer_name, counters):
    """utils/nn.py:95-100"""
    if layer_name not in counters:
        counters[layer_name] = 0
    name = layer_name + "_" + str(counters[layer_name])
    counters[layer_name] += 1
    return name


def l2_normalize(v, axes, eps=1e-12):
    """tf.nn.l2_normalize(v, axes): v * rsqrt(max(sum(v^2, axes), eps))"""
    sq = torch.sum(v * v, dim=axes, keepdim=True)
    return v * torch.rsqrt(torch.clamp(sq, min=eps))


def _as_list(x):
    if isinstance(x, tuple):
        return list(x)
    if not isinstance(x, list):
        return [x]
    return x


def apply_pre_activation(x, pre_activation, axis=3):
    """utils/nn.py:190-206.  For crelu/celu on a list: [x0, -x0, x1, -x1, ...] (interleaved per list element)."""
    x = _as_list(x)
    if pre_activation is None:
        return torch.cat(x, axis) if len(x) > 1 else x[0]
    if pre_activation == "celu":
        return F.elu(torch.cat([xs for xi in x for xs in (xi, -xi)], axis))
    if pre_activation == "crelu":
        return F.relu(torch.cat([xs for xi in x for xs in (xi, -xi)], axis))
    if pre_activation == "elu":
        return F.elu(torch.cat(x, axis))
    if pre_activation == "relu":
        return F.relu(torch.cat(x, axis))
    raise ValueError("unsupported pre-activation")


def same_padding(n, k, s):
    """TensorFlow 'SAME': out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total // 2, after = the rest."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def resize_nearest_neighbor(x, size):
    """tf.image.resize_nearest_neighbor(x, [H, W]) (align_corners=False): src = floor(dst * in / out).  NHWC."""
    B, H, W, C = x.shape
    oh, ow = int(size[0]), int(size[1])
    if oh == H and ow == W:
        return x
    if oh % H == 0 and ow % W == 0:
        return x.repeat_interleave(oh // H, dim=1).repeat_interleave(ow // W, dim=2)
    ih = torch.div(torch.arange(oh, device=x.device) * H, oh, rounding_mode="floor")
    iw = torch.div(torch.arange(ow, device=x.device) * W, ow, rounding_mode="floor")
    return x[:, ih][:, :, iw]


def _conv2d_nhwc(x, W, stride, pad, bias=None):
    """tf.nn.conv2d(x, W, [1,s,s,1], pad) with NHWC input and HWIO kernel (utils/nn.py:241)."""
    if isinstance(W, TransposedWeight):
        kh, kw = W.vshape[0], W.vshape[1]
        w_oihw = W.as_oihw()
    else:
        kh, kw = W.shape[0], W.shape[1]
        w_oihw = W.permute(3, 2, 0, 1)
    xn = x.permute(0, 3, 1, 2)                                 # NCHW view of channels-last memory: no copy
    if pad == "SAME":
        pt, pb = same_padding(x.shape[1], kh, stride[0])
        pl, pr = same_padding(x.shape[2], kw, stride[1])
        if pt == pb and pl == pr:
            y = F.conv2d(xn, w_oihw, bias, stride=tuple(stride), padding=(pt, pl))
        else:
            y = F.conv2d(F.pad(xn, (pl, pr, pt, pb)), w_oihw, bias, stride=tuple(stride))
    elif pad == "VALID":
        y = F.conv2d(xn, w_oihw, bias, stride=tuple(stride))
    else:
        raise ValueError(pad)
    return y.permute(0, 2, 3, 1)


CONV_BACKEND = "tcgen05"      # "tcgen05": this library's implicit-GEMM kernels where they tile the shape; "cudnn": library rung only
UPSAMPLE_FUSION = True        # nn.upsample2x / nn.glu(upsample=True) hand the following conv2d an un-materialised Upsampled2x
CONV_NARROW = True            # route the two 3-channel layers through _ConvNarrow (False: cuDNN, for A/B timing)
DENSE_ON_TCGEN05 = True       # nn.dense as a 1x1 convolution on the generic tcgen05 kernels (False: cuBLAS through F.linear)
WN_FUSION = True              # weight norm fused into the tcgen05 convolution nodes (_ConvTCWN / _ConvUp2TCWN: HWIO gradient pipeline)
DENSE_BLOCK_FUSION = True     # nn.dense_block runs DenseNet's blocks on the dense-block kernels (False: the literal list code)
CRELU_FUSION = True           # nn.conv2d(crelu_out=True) writes the CReLU from the convolution's epilogue (False: separate pass, for A/B tests)
CRELU_FUSION_MIN_TILES = 148  # ... for launches with at least this many output tiles (the fused epilogue is never split over the taps)
_conv_ws = {}
_wn_ws = {}
_retired_ws = []              # outgrown scratch buffers stay alive: a captured CUDA graph may still hold their addresses


def _workspace(device, nbytes):
    """Grow-only per-device scratch buffer for the convolution kernels (split-K partials, bias-gradient partials)."""
    ws = _conv_ws.get(device.index)
    if ws is None or ws.numel() * 4 < nbytes:
        if ws is not None:
            _retired_ws.append(ws)
        ws = _conv_ws[device.index] = torch.empty((max(nbytes // 4 + 64, 1 << 20),), device=device, dtype=torch.float32)
    return ws


def _wn_workspace(device, need_floats):
    """Scratch of the weight-norm kernels (same retirement rule as _workspace)."""
    ws = _wn_ws.get(device.index)
    if ws is None or ws.numel() < need_floats:
        if ws is not None:
            _retired_ws.append(ws)
        ws = _wn_ws[device.index] = torch.empty((max(need_floats, 64 * 32768),), device=device, dtype=torch.float32)
    return ws


def _pow2(v):
    return v > 0 and (v & (v - 1)) == 0


def conv_tc_supported(xshape, cout, kh, kw, stride, pad):
    """Shapes the tcgen05 implicit-GEMM kernels tile (fprop, dgrad and wgrad alike): channels in multiples of 128,
    power-of-two extents, and a batch that fills whole 128-pixel (fprop/dgrad) and 32-pixel (wgrad) boxes."""
    B, H, W, cin = xshape
    if pad != "SAME" or stride[0] != stride[1] or stride[0] not in (1, 2) or kh * kw > 32:
        return False
    s = stride[0]
    if cin % 128 or cout % 128 or H % s or W % s or kh < s or kw < s:        # k < stride: dgrad would have an empty parity class
        return False
    ho, wo = H // s, W // s
    if not (_pow2(ho) and _pow2(wo)) or wo > 32:
        return False
    img = ho * wo
    return (img >= 128 or (B * img) % 128 == 0) and (img >= 32 or (B * img) % 32 == 0)


class _ConvTC(torch.autograd.Function):
    """tf.nn.conv2d(x, W, [1,s,s,1], 'SAME') + bias_add on this library's tcgen05 implicit-GEMM kernels
    (otgan_conv2d_{fprop,dgrad,wgrad}_tf32, otgan_colsum_f32).  x: NHWC, wt: OHWI [Cout, kh*kw*Cin]."""

    @staticmethod
    def forward(ctx, x, wt, bias, geom, wt_ihwo=None):
        lib = _lib.load()
        kh, kw, s, pt, pl = geom
        ctx.wt_ihwo = wt_ihwo
        B, H, W, cin = x.shape
        cout = wt.shape[0]
        x, wt = x.contiguous(), wt.contiguous()
        if bias is not None and bias.data_ptr() % 16:
            bias = bias.clone()
        y = torch.empty((B, H // s, W // s, cout), device=x.device, dtype=torch.float32)
        ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_gemm(B, H // s, W // s, cout))
        rc = lib.otgan_conv2d_fprop_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, x.data_ptr(), wt.data_ptr(),
                                         bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                         ws.data_ptr(), ws.numel() * 4, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_conv2d_fprop_tf32")
        ctx.save_for_backward(x, wt)
        ctx.geom, ctx.has_bias = geom, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, wt = ctx.saved_tensors
        kh, kw, s, pt, pl = ctx.geom
        B, H, W, cin = x.shape
        cout = wt.shape[0]
        dy = dy.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        dx = dwt = db = None
        if ctx.needs_input_grad[0]:
            wt_t = ctx.wt_ihwo
            if wt_t is None:
                wt_t = torch.empty((cin, kh * kw * cout), device=x.device, dtype=torch.float32)
                _lib.check(lib.otgan_ohwi_to_ihwo_f32(cout, kh * kw, cin, wt.data_ptr(), wt_t.data_ptr(), stream), "otgan_ohwi_to_ihwo_f32")
            dx = torch.empty_like(x)
            ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_gemm(B, H, W, cin))
            rc = lib.otgan_conv2d_dgrad_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, dy.data_ptr(), wt_t.data_ptr(), dx.data_ptr(),
                                             ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_dgrad_tf32")
        if ctx.needs_input_grad[1]:
            need = lib.otgan_workspace_bytes_conv_wgrad(B, H, W, cin, cout, kh, kw, s)
            ws = _workspace(x.device, need)
            dwt = torch.empty_like(wt)
            rc = lib.otgan_conv2d_wgrad_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, dy.data_ptr(), x.data_ptr(), dwt.data_ptr(),
                                             ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_wgrad_tf32")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            P = dy.numel() // cout
            ws = _workspace(x.device, lib.otgan_workspace_bytes_colsum(P, cout))
            db = torch.empty((cout,), device=x.device, dtype=torch.float32)
            _lib.check(lib.otgan_colsum_f32(P, cout, dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream), "otgan_colsum_f32")
        return dx, dwt, db, None, None


def conv_narrow_supported(xshape, cout, kh, kw, stride, pad):
    """The two 3-channel layers of DCGAN (critic conv2d_0: 3 -> 128, generator conv2d_3: 128 -> 3): stride 1, one side of
    the layer has <= 16 channels, the other a multiple of 128."""
    B, H, W, cin = xshape
    if pad != "SAME" or tuple(stride) != (1, 1) or kh * kw > 32 or kh % 2 == 0 or kw % 2 == 0:
        return False
    if not ((cin <= 16 and cout == 128) or (cin == 128 and cout <= 16)) or kh * kw * min(cin, cout) > 128:
        return False
    if not (_pow2(H) and _pow2(W)) or W > 32:
        return False
    return H * W >= 128 or (B * H * W) % 128 == 0


def _pad_channels(t, c):
    """Zero-pad the last (channel) axis to c entries."""
    return t if t.shape[-1] == c else F.pad(t, (0, c - t.shape[-1]))


class _ConvNarrow(torch.autograd.Function):
    """The 3-channel convolutions (critic conv2d_0: 3 -> 128, generator conv2d_3: 128 -> 3) on this library's kernels.
    With 3 channels on one side a per-tap implicit GEMM re-reads the wide tensor 25 times with no reuse, so:
        narrow OUTPUT (generator fprop, critic dgrad): one 1x1 GEMM  z[px][t*C + c] = wide[px] . w2[t*C + c]  on the tcgen05
            kernel (the wide tensor is read once), then otgan_col2im_narrow_f32 shifts and sums the 25 taps;
        narrow INPUT of a filter gradient (both layers): otgan_im2col_narrow_f32 expands the 3-channel tensor to
            [pixels, 128] columns, then ONE 1x1 wgrad GEMM on the tcgen05 kernel (split over all SMs along the pixels);
        the remaining two passes (critic fprop, generator dgrad) run the GEMM kernel on the 3-channel tensor / filter
            zero-padded to 32 channels."""

    @staticmethod
    def _gemm1x1(lib, x, w2, B, H, W, stream):
        """z [B,H,W,128] = x [B,H,W,K] . w2[128, K]^T  (conv_gemm_tc_kernel as a 1x1 convolution)."""
        z = torch.empty((B, H, W, w2.shape[0]), device=x.device, dtype=torch.float32)
        rc = lib.otgan_conv2d_fprop_tf32(B, H, W, x.shape[3], w2.shape[0], 1, 1, 1, 0, 0, x.data_ptr(), w2.data_ptr(), None,
                                         z.data_ptr(), None, 0, stream)
        _lib.check(rc, "otgan_conv2d_fprop_tf32 (1x1)")
        return z

    @staticmethod
    def _wgrad1x1(lib, wide, col, B, H, W, stream):
        """out [Cw, 128] = sum_px wide[px][:]^T col[px][:]   (conv_wgrad_tc_kernel as a 1x1 convolution)."""
        cw = wide.shape[-1]
        need = lib.otgan_workspace_bytes_conv_wgrad(B, H, W, 128, cw, 1, 1, 1)
        ws = _workspace(wide.device, need)
        out = torch.empty((cw, 128), device=wide.device, dtype=torch.float32)
        rc = lib.otgan_conv2d_wgrad_tf32(B, H, W, 128, cw, 1, 1, 1, 0, 0, wide.data_ptr(), col.data_ptr(), out.data_ptr(),
                                         ws.data_ptr(), ws.numel() * 4, stream)
        _lib.check(rc, "otgan_conv2d_wgrad_tf32 (1x1)")
        return out

    @staticmethod
    def forward(ctx, x, wt, bias, geom, crelu=False):
        lib = _lib.load()
        kh, kw, s, pt, pl = geom
        B, H, W, cin = x.shape
        cout = wt.shape[0]
        taps = kh * kw
        x, wt = x.contiguous(), wt.contiguous()
        if bias is not None and bias.data_ptr() % 16:
            bias = bias.clone()
        stream = torch.cuda.current_stream().cuda_stream
        ctx.crelu = bool(crelu)
        y = torch.empty((B, H, W, 2 * cout if crelu else cout), device=x.device, dtype=torch.float32)
        if cin <= 16 and crelu:                            # critic conv2d_0 with the next layer's CReLU written by the epilogue
            xk = _pad_channels(x, 32)
            wk = _pad_channels(wt.view(cout, taps, cin), 32).reshape(cout, -1)
            rc = lib.otgan_conv2d_fprop_crelu_tf32(B, H, W, 32, cout, kh, kw, 1, pt, pl, xk.data_ptr(), wk.data_ptr(),
                                                   bias.data_ptr() if bias is not None else None, y.data_ptr(), stream)
            _lib.check(rc, "otgan_conv2d_fprop_crelu_tf32")
        elif cin <= 16:                                    # critic conv2d_0: image and filter zero-padded to 32 channels
            xk = _pad_channels(x, 32)
            wk = _pad_channels(wt.view(cout, taps, cin), 32).reshape(cout, -1)
            ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_gemm(B, H, W, cout))
            rc = lib.otgan_conv2d_fprop_tf32(B, H, W, 32, cout, kh, kw, 1, pt, pl, xk.data_ptr(), wk.data_ptr(),
                                             bias.data_ptr() if bias is not None else None, y.data_ptr(), ws.data_ptr(),
                                             ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_fprop_tf32")
        else:                                              # generator conv2d_3: 1x1 GEMM over the 75 (tap, co) columns + shift
            w2 = torch.zeros((128, cin), device=x.device, dtype=torch.float32)
            w2[:taps * cout] = wt.view(cout, taps, cin).permute(1, 0, 2).reshape(taps * cout, cin)
            z = _ConvNarrow._gemm1x1(lib, x, w2, B, H, W, stream)
            rc = lib.otgan_col2im_narrow_f32(B, H, W, cout, kh, kw, pt, pl, 0, z.data_ptr(), 128,
                                             bias.data_ptr() if bias is not None else None, y.data_ptr(), stream)
            _lib.check(rc, "otgan_col2im_narrow_f32")
        if crelu:
            assert cin <= 16, "the fused CReLU output exists for the narrow-INPUT layer only"
            ctx.save_for_backward(x, wt, y)
        else:
            ctx.save_for_backward(x, wt)
        ctx.geom, ctx.has_bias = geom, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, wt = ctx.saved_tensors[:2]
        kh, kw, s, pt, pl = ctx.geom
        B, H, W, cin = x.shape
        cout = wt.shape[0]
        taps = kh * kw
        dy = dy.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        if ctx.crelu:                                      # gradient of z = crelu(y) -> gradient of y
            dy = _crelu_bwd_from_z(lib, ctx.saved_tensors[2], dy, stream)
        dx = dwt = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if cin <= 16:                                  # critic conv2d_0: z[px][t*cin + ci] = dy[px] . W[:, t, ci], then shift
                w2 = torch.zeros((128, cout), device=x.device, dtype=torch.float32)
                w2[:taps * cin] = wt.t()
                z = _ConvNarrow._gemm1x1(lib, dy, w2, B, H, W, stream)
                _lib.check(lib.otgan_col2im_narrow_f32(B, H, W, cin, kh, kw, pt, pl, 1, z.data_ptr(), 128, None, dx.data_ptr(), stream),
                           "otgan_col2im_narrow_f32")
            else:                                          # generator conv2d_3: dy and the IHWO filter zero-padded to 32 channels
                dyk = _pad_channels(dy, 32)
                wt_t = _pad_channels(wt.view(cout, taps, cin).permute(2, 1, 0), 32).reshape(cin, -1).contiguous()
                ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_gemm(B, H, W, cin))
                rc = lib.otgan_conv2d_dgrad_tf32(B, H, W, cin, 32, kh, kw, 1, pt, pl, dyk.data_ptr(), wt_t.data_ptr(), dx.data_ptr(),
                                                 ws.data_ptr(), ws.numel() * 4, stream)
                _lib.check(rc, "otgan_conv2d_dgrad_tf32")
        if ctx.needs_input_grad[1]:
            col = torch.empty((B, H, W, 128), device=x.device, dtype=torch.float32)
            if cin <= 16:                                  # dW[co][t*cin + ci] = sum_px dy[px][co] x[px + off_t][ci]
                _lib.check(lib.otgan_im2col_narrow_f32(B, H, W, cin, kh, kw, pt, pl, 0, x.data_ptr(), col.data_ptr(), 128, stream),
                           "otgan_im2col_narrow_f32")
                out = _ConvNarrow._wgrad1x1(lib, dy, col, B, H, W, stream)
                dwt = out[:, :taps * cin].contiguous()
            else:                                          # dW[co][t][ci] = sum_px x[px][ci] dy[px - off_t][co]
                _lib.check(lib.otgan_im2col_narrow_f32(B, H, W, cout, kh, kw, pt, pl, 1, dy.data_ptr(), col.data_ptr(), 128, stream),
                           "otgan_im2col_narrow_f32")
                out = _ConvNarrow._wgrad1x1(lib, x, col, B, H, W, stream)          # [cin][t*cout + co]
                dwt = out[:, :taps * cout].reshape(cin, taps, cout).permute(2, 1, 0).reshape(cout, taps * cin)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if cout % 4 == 0:
                P = dy.numel() // cout
                ws = _workspace(x.device, lib.otgan_workspace_bytes_colsum(P, cout))
                db = torch.empty((cout,), device=x.device, dtype=torch.float32)
                _lib.check(lib.otgan_colsum_f32(P, cout, dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream), "otgan_colsum_f32")
            else:
                db = dy.reshape(-1, cout).sum(0)
        return dx, dwt, db, None, None


class Upsampled2x:
    """A 2x nearest-neighbour upsampled NHWC tensor that has NOT been materialised: `low` is the [B, H, W, C] tensor,
    the value is resize_nearest_neighbor(low, [2H, 2W]).  nn.conv2d consumes it with the fused upsample + convolution
    kernels (the generator's resize -> conv pairs, models/dcgan.py:37-46); anything else calls .materialize()."""

    def __init__(self, low):
        self.low = low

    @property
    def shape(self):
        B, H, W, C = self.low.shape
        return torch.Size((B, 2 * H, 2 * W, C))

    @property
    def is_cuda(self):
        return self.low.is_cuda

    @property
    def dtype(self):
        return self.low.dtype

    @property
    def device(self):
        return self.low.device

    def dim(self):
        return 4

    def materialize(self):
        return resize_nearest_neighbor(self.low, [2 * self.low.shape[1], 2 * self.low.shape[2]])


def upsample2x(x, lazy=True):
    """tf.image.resize_nearest_neighbor(x, [2H, 2W]) (models/dcgan.py:38,42,46).  With lazy=True on the GPU path the result
    is an Upsampled2x handle that the following nn.conv2d fuses with its convolution."""
    if lazy and UPSAMPLE_FUSION and CONV_BACKEND == "tcgen05" and x.is_cuda and x.dtype == torch.float32:
        return Upsampled2x(x.contiguous())
    return resize_nearest_neighbor(x, [2 * x.shape[1], 2 * x.shape[2]])


class CreluOut:
    """The CReLU-activated output of a convolution, relu(concat([y, -y], 3)) (utils/nn.py:198-200 on one tensor), produced by
    nn.conv2d(..., crelu_out=True): `z` is the [B, H, W, 2C] tensor, `shape` reports the RAW [B, H, W, C] like the tensor the
    reference passes between the layers.  The next nn.conv2d(..., pre_activation='crelu') consumes `z` as it is.  On the tcgen05
    path the producing convolution writes z from its epilogue (otgan_conv2d_fprop_crelu_tf32): the separate CReLU pass -- a read
    of y and a write of 2x its size per critic layer -- disappears."""

    def __init__(self, z):
        self.z = z

    @property
    def shape(self):
        B, H, W, C2 = self.z.shape
        return torch.Size((B, H, W, C2 // 2))

    @property
    def is_cuda(self):
        return self.z.is_cuda

    @property
    def dtype(self):
        return self.z.dtype

    @property
    def device(self):
        return self.z.device

    def dim(self):
        return 4


def _crelu_fusable(B, Ho, Wo, cout):
    """The fused CReLU epilogue is never split over the filter taps: use it when the launch has a tile per SM anyway."""
    if cout % 128 or not CRELU_FUSION:
        return False
    tn = 256 if cout % 256 == 0 else 128
    return (B * Ho * Wo // 128) * (cout // tn) >= CRELU_FUSION_MIN_TILES


def _crelu_bwd_from_z(lib, z, dz, stream):
    B, H, W, C2 = z.shape
    dy = torch.empty((B, H, W, C2 // 2), device=z.device, dtype=torch.float32)
    _lib.check(lib.otgan_crelu_bwd_from_activated_f32(B * H * W, C2 // 2, z.data_ptr(), dz.data_ptr(), dy.data_ptr(), stream),
               "otgan_crelu_bwd_from_activated_f32")
    return dy


def conv_up2_supported(lowshape, cout, kh, kw, stride, pad):
    """Shapes the fused upsample + convolution kernels tile."""
    B, H, W, cin = lowshape
    if pad != "SAME" or tuple(stride) != (1, 1) or kh != kw or kh % 2 == 0:
        return False
    if cin % 128 or cout % 128 or not (_pow2(H) and _pow2(W)) or W > 32:
        return False
    if _lib.load().otgan_up2_subtaps(kh, (kh - 1) // 2) == 0:
        return False
    img = H * W
    return (img >= 128 or (B * img) % 128 == 0) and (img >= 32 or (B * img) % 32 == 0)


class _ConvUp2TC(torch.autograd.Function):
    """conv2d(resize_nearest_neighbor(x_low, 2x), W, 'SAME') + bias on the fused kernels (otgan_conv2d_up2_{fprop,dgrad,wgrad}_tf32):
    four 3x3 convolutions of the low-resolution input with pre-summed sub-filters, one per output parity class."""

    @staticmethod
    def forward(ctx, x_low, wt, bias, geom):
        lib = _lib.load()
        kh, kw, pt, pl = geom
        B, H, W, cin = x_low.shape
        cout = wt.shape[0]
        n1 = lib.otgan_up2_subtaps(kh, pt)
        slots = n1 * n1
        x_low, wt = x_low.contiguous(), wt.contiguous()
        if bias is not None and bias.data_ptr() % 16:
            bias = bias.clone()
        stream = torch.cuda.current_stream().cuda_stream
        w_sub = torch.empty((4, cout, slots * cin), device=x_low.device, dtype=torch.float32)
        _lib.check(lib.otgan_up2_weight_presum_f32(cout, kh, kw, cin, pt, pl, wt.data_ptr(), w_sub.data_ptr(), stream), "otgan_up2_weight_presum_f32")
        y = torch.empty((B, 2 * H, 2 * W, cout), device=x_low.device, dtype=torch.float32)
        ws = _workspace(x_low.device, lib.otgan_workspace_bytes_conv_gemm(B, 2 * H, 2 * W, cout))
        rc = lib.otgan_conv2d_up2_fprop_tf32(B, H, W, cin, cout, kh, kw, pt, pl, x_low.data_ptr(), w_sub.data_ptr(),
                                             bias.data_ptr() if bias is not None else None, y.data_ptr(), ws.data_ptr(),
                                             ws.numel() * 4, stream)
        _lib.check(rc, "otgan_conv2d_up2_fprop_tf32")
        ctx.save_for_backward(x_low, w_sub)
        ctx.geom, ctx.has_bias, ctx.cout, ctx.slots = geom, bias is not None, cout, slots
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x_low, w_sub = ctx.saved_tensors
        kh, kw, pt, pl = ctx.geom
        B, H, W, cin = x_low.shape
        cout, slots = ctx.cout, ctx.slots
        dy = dy.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        dx = dwt = db = None
        if ctx.needs_input_grad[0]:
            w_sub_t = torch.empty((4, cin, slots * cout), device=dy.device, dtype=torch.float32)
            for c in range(4):
                _lib.check(lib.otgan_ohwi_to_ihwo_f32(cout, slots, cin, w_sub[c].data_ptr(), w_sub_t[c].data_ptr(), stream), "otgan_ohwi_to_ihwo_f32")
            dx = torch.empty_like(x_low)
            ws = _workspace(dy.device, lib.otgan_workspace_bytes_conv_gemm(B, H, W, cin))
            rc = lib.otgan_conv2d_up2_dgrad_tf32(B, H, W, cin, cout, kh, kw, pt, pl, dy.data_ptr(), w_sub_t.data_ptr(), dx.data_ptr(),
                                                 ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_up2_dgrad_tf32")
        if ctx.needs_input_grad[1]:
            ws = _workspace(dy.device, lib.otgan_workspace_bytes_conv_up2_wgrad(B, H, W, cin, cout, kh, kw, pt, pl))
            dw_sub = torch.empty_like(w_sub)
            rc = lib.otgan_conv2d_up2_wgrad_tf32(B, H, W, cin, cout, kh, kw, pt, pl, dy.data_ptr(), x_low.data_ptr(), dw_sub.data_ptr(),
                                                 ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_up2_wgrad_tf32")
            dwt = torch.empty((cout, kh * kw * cin), device=dy.device, dtype=torch.float32)
            _lib.check(lib.otgan_up2_weight_unsum_f32(cout, kh, kw, cin, pt, pl, dw_sub.data_ptr(), dwt.data_ptr(), stream), "otgan_up2_weight_unsum_f32")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            P = dy.numel() // cout
            ws = _workspace(dy.device, lib.otgan_workspace_bytes_colsum(P, cout))
            db = torch.empty((cout,), device=dy.device, dtype=torch.float32)
            _lib.check(lib.otgan_colsum_f32(P, cout, dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream), "otgan_colsum_f32")
        return dx, dwt, db, None


class _CreluL2Norm(torch.autograd.Function):
    """Critic head on this library's CUDA kernels (otgan_crelu_l2norm_{fwd,bwd}_f32)."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        B, H, W, C = x.shape
        y = torch.empty((B, H * W * 2 * C), device=x.device, dtype=torch.float32)
        inv = torch.empty((B,), device=x.device, dtype=torch.float32)
        rc = lib.otgan_crelu_l2norm_fwd_f32(B, H * W, C, x.data_ptr(), y.data_ptr(), inv.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_crelu_l2norm_fwd_f32")
        ctx.save_for_backward(x, y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, y, inv = ctx.saved_tensors
        B, H, W, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        rc = lib.otgan_crelu_l2norm_bwd_f32(B, H * W, C, x.data_ptr(), y.data_ptr(), inv.data_ptr(), dy.data_ptr(),
                                            dx.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_crelu_l2norm_bwd_f32")
        return dx


def crelu_l2norm(x):
    """The critic head of models/dcgan.py:16-19 / models/densenet.py:37-42:
        x = concat([relu(x), relu(-x)], 3); x = reshape(x, [B, -1]); x /= sqrt(reduce_sum(square(x), 1, keep_dims))
    One fused CUDA kernel (forward and backward) on the GPU; the literal op sequence for host-side (CPU tensor) tests."""
    if x.is_cuda and x.dtype == torch.float32:
        return _CreluL2Norm.apply(x.contiguous())
    x = torch.cat([torch.relu(x), torch.relu(-x)], 3)
    x = x.reshape(x.shape[0], -1)
    return x / torch.sqrt(torch.sum(torch.square(x), dim=1, keepdim=True))


class _CreluPad(torch.autograd.Function):
    """relu(concat([x, -x], 3)) written into a zero-padded NHWC buffer (otgan_crelu_pad_{fwd,bwd}_f32)."""

    @staticmethod
    def forward(ctx, x, pads):
        lib = _lib.load()
        B, H, W, C = x.shape
        pt, pl, pb, pr = pads
        z = torch.empty((B, H + pt + pb, W + pl + pr, 2 * C), device=x.device, dtype=torch.float32)
        rc = lib.otgan_crelu_pad_fwd_f32(B, H, W, C, pt, pl, pb, pr, x.data_ptr(), z.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_crelu_pad_fwd_f32")
        ctx.save_for_backward(x)
        ctx.pads = pads
        return z

    @staticmethod
    def backward(ctx, dz):
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        B, H, W, C = x.shape
        pt, pl, pb, pr = ctx.pads
        dz = dz.contiguous()
        dx = torch.empty_like(x)
        rc = lib.otgan_crelu_pad_bwd_f32(B, H, W, C, pt, pl, pb, pr, x.data_ptr(), dz.data_ptr(), dx.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_crelu_pad_bwd_f32")
        return dx, None


class _GluUp(torch.autograd.Function):
    """a * sigmoid(l) (+ 2x nearest-neighbour upsample), (a, l) = split(y, 2, 3)  (otgan_glu_up_{fwd,bwd}_f32)."""

    @staticmethod
    def forward(ctx, y, up):
        lib = _lib.load()
        B, H, W, C2 = y.shape
        C = C2 // 2
        out = torch.empty((B, up * H, up * W, C), device=y.device, dtype=torch.float32)
        rc = lib.otgan_glu_up_fwd_f32(B, H, W, C, up, y.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_glu_up_fwd_f32")
        ctx.save_for_backward(y)
        ctx.up = up
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        B, H, W, C2 = y.shape
        dout = dout.contiguous()
        dy = torch.empty_like(y)
        rc = lib.otgan_glu_up_bwd_f32(B, H, W, C2 // 2, ctx.up, y.data_ptr(), dout.data_ptr(), dy.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_glu_up_bwd_f32")
        return dy, None


def glu(y, upsample=False):
    """x, l = tf.split(y, 2, 3); x *= tf.nn.sigmoid(l)  [; x = resize_nearest_neighbor(x, 2x)]   (models/dcgan.py:39-48).
    One fused CUDA kernel on the GPU; the literal ops for CPU tensors.  With upsample=True on the GPU path the result is an
    un-materialised Upsampled2x handle (see upsample2x) so that the following nn.conv2d fuses the resize."""
    if y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and (y.shape[3] // 2) % 4 == 0:
        if upsample and UPSAMPLE_FUSION and CONV_BACKEND == "tcgen05":
            return Upsampled2x(_GluUp.apply(y.contiguous(), 1))
        return _GluUp.apply(y.contiguous(), 2 if upsample else 1)
    x, l = torch.chunk(y, 2, 3)
    x = x * torch.sigmoid(l)
    if upsample:
        x = resize_nearest_neighbor(x, [2 * x.shape[1], 2 * x.shape[2]])
    return x


class TransposedWeight:
    """W = g * V / ||V|| held output-channel-major ([C, K] == OHWI): what otgan_weightnorm_fwd_f32 writes and what the
    convolution / F.linear consume without another layout copy."""

    def __init__(self, wt, vshape, wt_ihwo=None):
        self.wt, self.vshape, self.wt_ihwo = wt, tuple(vshape), wt_ihwo       # wt_ihwo: cached [Cin, kh*kw*Cout] copy (dgrad)

    def as_oihw(self):
        kh, kw, ci, co = self.vshape
        return self.wt.view(co, kh, kw, ci).permute(0, 3, 1, 2)      # logical OIHW, channels-last memory: no copy


class _WeightNorm(torch.autograd.Function):
    """utils/nn.py:176-180 on this library's fused CUDA kernels (otgan_weightnorm_{fwd,bwd}_f32)."""

    @staticmethod
    def forward(ctx, V, g):
        lib = _lib.load()
        C = V.shape[-1]
        K = V.numel() // C
        Vc, gc = V.contiguous(), g.contiguous()
        wt = torch.empty((C, K), device=V.device, dtype=torch.float32)
        inv = torch.empty((C,), device=V.device, dtype=torch.float32)
        ws = _wn_workspace(V.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
        rc = lib.otgan_weightnorm_fwd_f32(K, C, Vc.data_ptr(), gc.data_ptr(), wt.data_ptr(), inv.data_ptr(), ws.data_ptr(),
                                          ws.numel() * 4, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_weightnorm_fwd_f32")
        ctx.save_for_backward(Vc, gc, inv)
        ctx.vshape = V.shape
        return wt

    @staticmethod
    def backward(ctx, dwt):
        lib = _lib.load()
        V, g, inv = ctx.saved_tensors
        C = V.shape[-1]
        K = V.numel() // C
        dwt = dwt.contiguous()
        dV, dg = torch.empty_like(V), torch.empty_like(g)
        ws = _wn_workspace(V.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
        rc = lib.otgan_weightnorm_bwd_f32(K, C, V.data_ptr(), g.data_ptr(), inv.data_ptr(), dwt.data_ptr(), dV.data_ptr(),
                                          dg.data_ptr(), ws.data_ptr(), ws.numel() * 4, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_weightnorm_bwd_f32")
        return dV.view(ctx.vshape), dg



# ------------------------------------------------------------------------------------------------ generic convolutions / DenseNet
def conv_gen_supported(xshape, cout, kh, kw, stride, pad):
    """Shapes the GENERIC mode of the tcgen05 kernels takes (otgan_conv2d_*_ex_tf32): 'SAME' padding, stride 1 or 2, power-of-two
    spatial extents; any batch and any channel counts (padded to multiples of 4 here)."""
    B, H, W, cin = xshape
    if pad != "SAME" or stride[0] != stride[1] or stride[0] not in (1, 2) or kh * kw > 40 or kh < stride[0] or kw < stride[0]:
        return False
    s = stride[0]
    return H % s == 0 and W % s == 0 and _pow2(H // s) and _pow2(W // s) and B >= 1


class _ConvGen(torch.autograd.Function):
    """tf.nn.conv2d(x, W, [1,s,s,1], 'SAME') + bias_add on the generic mode of this library's tcgen05 kernels: any batch / channel
    counts (multiples of 4; the caller pads), x may be any NHWC tensor including a crelu8 feature buffer.  wt: [Cout, kh*kw*Cin]."""

    @staticmethod
    def forward(ctx, x, wt, bias, geom):
        lib = _lib.load()
        kh, kw, s, pt, pl = geom
        B, H, W, cin = x.shape
        cout = wt.shape[0]
        x, wt = x.contiguous(), wt.contiguous()
        y = torch.empty((B, H // s, W // s, cout), device=x.device, dtype=torch.float32)
        rc = lib.otgan_conv2d_fprop_ex_tf32(B, H, W, cin, cin, cout, cout, kh, kw, s, pt, pl, x.data_ptr(), wt.data_ptr(),
                                            bias.data_ptr() if bias is not None else None, y.data_ptr(), 0,
                                            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_conv2d_fprop_ex_tf32")
        ctx.save_for_backward(x, wt)
        ctx.geom, ctx.has_bias = geom, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, wt = ctx.saved_tensors
        kh, kw, s, pt, pl = ctx.geom
        B, H, W, cin = x.shape
        cout = wt.shape[0]
        dy = dy.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        dx = dwt = db = None
        if ctx.needs_input_grad[0]:
            wt_t = torch.empty((cin, kh * kw * cout), device=x.device, dtype=torch.float32)
            _lib.check(lib.otgan_ohwi_to_ihwo_f32(cout, kh * kw, cin, wt.data_ptr(), wt_t.data_ptr(), stream), "otgan_ohwi_to_ihwo_f32")
            dx = torch.empty_like(x)
            rc = lib.otgan_conv2d_dgrad_ex_tf32(B, H, W, cin, cin, cout, cout, kh, kw, s, pt, pl, dy.data_ptr(), wt_t.data_ptr(),
                                                dx.data_ptr(), stream)
            _lib.check(rc, "otgan_conv2d_dgrad_ex_tf32")
        if ctx.needs_input_grad[1]:
            ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_wgrad_ex(B, H, W, cin, cout, kh, kw, s))
            dwt = torch.empty_like(wt)
            rc = lib.otgan_conv2d_wgrad_ex_tf32(B, H, W, cin, cin, cout, cout, kh, kw, s, pt, pl, dy.data_ptr(), x.data_ptr(),
                                                dwt.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_wgrad_ex_tf32")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            P = dy.numel() // cout
            ws = _workspace(x.device, lib.otgan_workspace_bytes_colsum(P, cout))
            db = torch.empty((cout,), device=x.device, dtype=torch.float32)
            _lib.check(lib.otgan_colsum_f32(P, cout, dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream), "otgan_colsum_f32")
        return dx, dwt, db, None


_perm_cache = {}


def crelu8_perm(elem_ch, taps, device):
    """Device int32 row permutation for a filter whose input is a crelu8 buffer of list elements with `elem_ch` channels:
    Wt[c][k_z] = V[perm[k_z]][c] (otgan_crelu8_perm_host).  Cached per (elements, taps, device)."""
    key = (tuple(int(c) for c in elem_ch), int(taps), device.index)
    t = _perm_cache.get(key)
    if t is None:
        import ctypes
        n = taps * 2 * sum(key[0])
        buf = (ctypes.c_int * n)()
        ch = (ctypes.c_int * len(key[0]))(*key[0])
        rc = _lib.load().otgan_crelu8_perm_host(len(key[0]), ch, taps, buf, n)
        if rc != n:
            _lib.check(rc if rc < 0 else -1, "otgan_crelu8_perm_host")
        t = _perm_cache[key] = torch.tensor(list(buf), dtype=torch.int32, device=device)
    return t


class _WeightNormPerm(torch.autograd.Function):
    """W = g V / ||V|| (utils/nn.py:176-180) written output-channel-major with the K rows permuted into crelu8 order."""

    @staticmethod
    def forward(ctx, V, g, perm):
        lib = _lib.load()
        C = V.shape[-1]
        K = V.numel() // C
        Vc, gc = V.contiguous(), g.contiguous()
        wt = torch.empty((C, K), device=V.device, dtype=torch.float32)
        inv = torch.empty((C,), device=V.device, dtype=torch.float32)
        ws = _wn_workspace(V.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
        rc = lib.otgan_weightnorm_fwd_ex_f32(K, C, Vc.data_ptr(), gc.data_ptr(), perm.data_ptr(), 0, 0, 0, wt.data_ptr(), inv.data_ptr(),
                                             ws.data_ptr(), ws.numel() * 4, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_weightnorm_fwd_ex_f32")
        ctx.save_for_backward(Vc, gc, inv, perm)
        ctx.vshape = V.shape
        return wt

    @staticmethod
    def backward(ctx, dwt):
        lib = _lib.load()
        V, g, inv, perm = ctx.saved_tensors
        C = V.shape[-1]
        K = V.numel() // C
        dwt = dwt.contiguous()
        dV, dg = torch.empty_like(V), torch.empty_like(g)
        ws = _wn_workspace(V.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
        rc = lib.otgan_weightnorm_bwd_ex_f32(K, C, V.data_ptr(), g.data_ptr(), inv.data_ptr(), perm.data_ptr(), 0, 0, 0, dwt.data_ptr(),
                                             dV.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "otgan_weightnorm_bwd_ex_f32")
        return dV.view(ctx.vshape), dg, None


class Crelu8Tensor:
    """The CReLU-activated concatenation of a Python list of NHWC tensors, held as ONE buffer in this library's crelu8 slot
    order (element i with c_i channels at raw offset o_i -> channels [2 o_i, 2 o_i + 2 c_i): blocks of 8 positive parts then
    their 8 negative parts).  Equal, up to that fixed channel permutation, to the reference's
    relu(concat([x_0, -x_0, x_1, -x_1, ...], 3)) (utils/nn.py:198-200); nn.conv2d consumes it with the filter's input channels
    permuted to match.  Produced by nn.dense_block."""

    def __init__(self, z, elem_ch):
        self.z, self.elem_ch = z, [int(c) for c in elem_ch]

    @property
    def raw_channels(self):
        return sum(self.elem_ch)

    @property
    def shape(self):
        B, H, W, _ = self.z.shape
        return torch.Size((B, H, W, self.raw_channels))

    @property
    def device(self):
        return self.z.device

    def upsample2x(self):
        """models/densenet.py:64-68 (`upsample`): tf.concat(x, 3) -> resize_nearest_neighbor -> conv2d(crelu).  The resize
        commutes with the element-wise CReLU, and because the reference CONCATENATES the list before that convolution its CReLU
        sees ONE tensor ([relu(cat), relu(-cat)], not the per-element interleave): the result is a single-element
        Crelu8Tensor.  The crelu8 layout itself is unchanged (blocks of 8 raw channels; element offsets are multiples of 8)."""
        return Crelu8Tensor(resize_nearest_neighbor(self.z, [2 * self.z.shape[1], 2 * self.z.shape[2]]), [self.raw_channels])


class _DenseBlock(torch.autograd.Function):
    """One dense block of models/densenet.py on otgan_dense_block_{fprop,bgrad}_tf32 (csrc/dense_block.cu).
    Inputs: geometry, the base list elements [B,H,W,c_i], then (V_r, g_r, b_r) of the L layers.  Output: Z (crelu8 buffer)."""

    @staticmethod
    def forward(ctx, n_base, L, *tensors):
        import ctypes
        lib = _lib.load()
        base = [t.contiguous() for t in tensors[:n_base]]
        Vs, gs, bs = tensors[n_base::3], tensors[n_base + 1::3], tensors[n_base + 2::3]
        B, H, W, _ = base[0].shape
        dev = base[0].device
        stream = torch.cuda.current_stream().cuda_stream
        geom = _lib.DenseGeom()
        geom.B, geom.H, geom.W, geom.n_base, geom.L, geom.growth = B, H, W, n_base, L, 16
        base_ch = [int(t.shape[3]) for t in base]
        for i, c in enumerate(base_ch):
            geom.base_ch[i] = c
        ctot = lib.otgan_dense_channels(ctypes.byref(geom))
        if ctot < 0:
            _lib.check(ctot, "otgan_dense_channels")
        Z = torch.empty((B, H, W, ctot), device=dev, dtype=torch.float32)
        P = B * H * W
        off = 0
        for t, c in zip(base, base_ch):
            _lib.check(lib.otgan_crelu8_fwd_f32(P, c, t.data_ptr(), c, Z.data_ptr() + 4 * 2 * off, ctot, stream), "otgan_crelu8_fwd_f32")
            off += c
        c0 = off
        # W_r = g V / ||V|| of all layers into ONE tensor [16L][9][Ctot] (input channels in crelu8 order, zero past cin_r)
        wf_all = torch.zeros((16 * L, 9, ctot), device=dev, dtype=torch.float32)
        invs = []
        for r in range(L):
            cin = 2 * (c0 + 16 * r)
            K = 9 * cin
            V, g = Vs[r].contiguous(), gs[r].contiguous()
            inv = torch.empty((16,), device=dev, dtype=torch.float32)
            perm = crelu8_perm(base_ch + [16] * r, 9, dev)
            ws = _wn_workspace(dev, lib.otgan_workspace_bytes_weightnorm(K, 16) // 4)
            rc = lib.otgan_weightnorm_fwd_ex_f32(K, 16, V.data_ptr(), g.data_ptr(), perm.data_ptr(), cin, ctot, 9 * ctot,
                                                 wf_all.data_ptr() + 4 * 16 * r * 9 * ctot, inv.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_weightnorm_fwd_ex_f32")
            invs.append(inv)
        bias_all = torch.cat([b.reshape(-1) for b in bs])
        S = torch.empty((B, H, W, 16 * L), device=dev, dtype=torch.float32)
        rc = lib.otgan_dense_block_fprop_tf32(ctypes.byref(geom), wf_all.data_ptr(), bias_all.data_ptr(), Z.data_ptr(), S.data_ptr(), stream)
        _lib.check(rc, "otgan_dense_block_fprop_tf32")
        ctx.geom, ctx.base_ch, ctx.L, ctx.n_base, ctx.ctot = geom, base_ch, L, n_base, ctot
        ctx.wf_all, ctx.invs = wf_all, invs
        ctx.save_for_backward(Z, *[t.contiguous() for t in Vs], *[t.contiguous() for t in gs])
        return Z

    @staticmethod
    def backward(ctx, dZ):
        import ctypes
        lib = _lib.load()
        saved = ctx.saved_tensors
        L, n_base, ctot, base_ch, geom = ctx.L, ctx.n_base, ctx.ctot, ctx.base_ch, ctx.geom
        Z, Vs, gs = saved[0], saved[1:1 + L], saved[1 + L:1 + 2 * L]
        B, H, W, _ = Z.shape
        dev = Z.device
        stream = torch.cuda.current_stream().cuda_stream
        dZ = dZ.contiguous()
        WB = torch.empty((lib.otgan_dense_wb_floats(ctypes.byref(geom)),), device=dev, dtype=torch.float32)
        _lib.check(lib.otgan_dense_build_wb_f32(ctypes.byref(geom), ctx.wf_all.data_ptr(), WB.data_ptr(), stream), "otgan_dense_build_wb_f32")
        dY = torch.empty((B, H, W, 16 * L), device=dev, dtype=torch.float32)
        dbase = [torch.empty((B, H, W, c), device=dev, dtype=torch.float32) for c in base_ch]
        need_w = any(ctx.needs_input_grad[2 + n_base + 3 * r + j] for r in range(L) for j in range(3))
        dW_all = torch.empty((16 * L, 9, ctot), device=dev, dtype=torch.float32) if need_w else None
        db_all = torch.empty((16 * L,), device=dev, dtype=torch.float32) if need_w else None
        ws = _workspace(dev, lib.otgan_workspace_bytes_dense_bgrad(ctypes.byref(geom)))
        rc = lib.otgan_dense_block_bgrad_tf32(ctypes.byref(geom), Z.data_ptr(), dZ.data_ptr(), WB.data_ptr(), dY.data_ptr(),
                                              _lib.ptr_array([t.data_ptr() for t in dbase]),
                                              dW_all.data_ptr() if need_w else None, db_all.data_ptr() if need_w else None,
                                              ws.data_ptr(), ws.numel() * 4, stream)
        _lib.check(rc, "otgan_dense_block_bgrad_tf32")
        c0 = sum(base_ch)
        grads = []
        for r in range(L):
            if not need_w:
                grads += [None, None, None]
                continue
            cin = 2 * (c0 + 16 * r)
            K = 9 * cin
            V, g = Vs[r], gs[r]
            dV, dg = torch.empty_like(V), torch.empty_like(g)
            perm = crelu8_perm(base_ch + [16] * r, 9, dev)
            wsn = _wn_workspace(dev, lib.otgan_workspace_bytes_weightnorm(K, 16) // 4)
            rc = lib.otgan_weightnorm_bwd_ex_f32(K, 16, V.data_ptr(), g.data_ptr(), ctx.invs[r].data_ptr(), perm.data_ptr(), cin, ctot,
                                                 9 * ctot, dW_all.data_ptr() + 4 * 16 * r * 9 * ctot, dV.data_ptr(), dg.data_ptr(),
                                                 wsn.data_ptr(), wsn.numel() * 4, stream)
            _lib.check(rc, "otgan_weightnorm_bwd_ex_f32")
            grads += [dV, dg, db_all[16 * r:16 * r + 16]]
        return (None, None, *dbase, *grads)


@add_arg_scope
def dense_block(x, layers_per_block, filters_per_layer, pre_activation="crelu", counters=None, init=False, ema=None,
                weight_norm=True, **kwargs):
    """models/densenet.py:10-15 / 56-61 (`block`):  for rep in range(layers_per_block): x.append(nn.conv2d(x, filters_per_layer,
    pre_activation=nonlinearity)).  Creates the same variables in the same order (conv2d_k/V, g, b) as those conv2d calls.
    On the GPU with CReLU and growth 16 the whole block runs on this library's dense-block kernels and the result is a
    Crelu8Tensor (the activated, concatenated list); otherwise the literal list code runs and the list is returned."""
    x = list(x) if isinstance(x, (list, tuple)) else [x]
    assert counters is not None, "dense_block needs the layer counters of the enclosing arg_scope"
    fused = (CONV_BACKEND == "tcgen05" and DENSE_BLOCK_FUSION and weight_norm and pre_activation == "crelu" and filters_per_layer == 16
             and not init and len(x) <= 4
             and all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[3] % 8 == 0 for t in x)
             and _pow2(int(x[0].shape[1])) and _pow2(int(x[0].shape[2])) and layers_per_block <= 32 and _tls.store.frozen)
    if not fused:
        for rep in range(layers_per_block):
            x.append(conv2d(x, filters_per_layer, pre_activation=pre_activation, counters=counters, init=init, ema=ema))
        return x
    tensors = list(x)
    for rep in range(layers_per_block):
        layer_name = get_name("conv2d", counters)
        prm = get_params(layer_name, None, False, ema, raw=True)
        tensors += [prm["V"], prm["g"], prm["b"]]
    z = _DenseBlock.apply(len(x), layers_per_block, *tensors)
    return Crelu8Tensor(z, [int(t.shape[3]) for t in x] + [16] * layers_per_block)


class LazyWN:
    """W = g * V / ||V|| not yet materialised: the tcgen05 convolution paths fuse the weight norm with the convolution
    (_ConvTCWN / _ConvUp2TCWN: the whole parameter-space pipeline stays in the variable's HWIO layout, one transpose per layer
    and step); every other consumer calls .materialize() and gets the TransposedWeight of the stand-alone weight-norm kernels."""

    def __init__(self, V, g, cached=None):
        self.V, self.g, self.vshape, self.cached = V, g, tuple(V.shape), cached      # cached: (wt, wt_ihwo) of the weight cache

    def materialize(self):
        if self.cached is not None:
            return TransposedWeight(self.cached[0], self.vshape, self.cached[1])
        return TransposedWeight(_WeightNorm.apply(self.V, self.g), self.vshape)


def _wn_fwd2(lib, V, g, taps, want_ihwo, stream):
    """otgan_weightnorm_fwd2_f32: (wt [C, K], wt_ihwo [Cin, taps * C] or None, inv [C])."""
    C = V.shape[-1]
    K = V.numel() // C
    wt = torch.empty((C, K), device=V.device, dtype=torch.float32)
    inv = torch.empty((C,), device=V.device, dtype=torch.float32)
    ihwo = torch.empty((K // taps, taps * C), device=V.device, dtype=torch.float32) if want_ihwo else None
    ws = _wn_workspace(V.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
    rc = lib.otgan_weightnorm_fwd2_f32(K, C, taps, V.data_ptr(), g.data_ptr(), wt.data_ptr(), ihwo.data_ptr() if want_ihwo else None,
                                       inv.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
    _lib.check(rc, "otgan_weightnorm_fwd2_f32")
    return wt, ihwo, inv


def _wn_bwd_hwio(lib, V, g, inv, dw_hwio, stream):
    C = V.shape[-1]
    K = V.numel() // C
    dV, dg = torch.empty_like(V), torch.empty_like(g)
    ws = _wn_workspace(V.device, lib.otgan_workspace_bytes_weightnorm(K, C) // 4)
    rc = lib.otgan_weightnorm_bwd_hwio_f32(K, C, V.data_ptr(), g.data_ptr(), inv.data_ptr(), dw_hwio.data_ptr(), dV.data_ptr(), dg.data_ptr(),
                                           ws.data_ptr(), ws.numel() * 4, stream)
    _lib.check(rc, "otgan_weightnorm_bwd_hwio_f32")
    return dV, dg


class _ConvTCWN(torch.autograd.Function):
    """Weight norm (utils/nn.py:176-180) + tf.nn.conv2d 'SAME' + bias_add as ONE autograd node on the tcgen05 kernels.
    Forward: V -> W (OHWI) and the IHWO dgrad operand in one pass; backward: wgrad writes the filter gradient in V's own HWIO
    layout and the weight-norm backward streams it -- no gradient transposes.  cached = (wt, wt_ihwo) skips the weight norm
    (calls that build no parameter gradients, VariableStore.refresh_weight_cache)."""

    @staticmethod
    def forward(ctx, x, V, g, bias, geom, cached, crelu=False):
        lib = _lib.load()
        kh, kw, s, pt, pl = geom
        stream = torch.cuda.current_stream().cuda_stream
        B, H, W, cin = x.shape
        cout = V.shape[-1]
        x = x.contiguous()
        ctx.crelu = bool(crelu)
        if cached is not None:
            wt, ihwo, inv = cached[0], cached[1], None
            Vc = gc = None
        else:
            Vc, gc = V.contiguous(), g.contiguous()
            wt, ihwo, inv = _wn_fwd2(lib, Vc, gc, kh * kw, ctx.needs_input_grad[0], stream)
        if bias is not None and bias.data_ptr() % 16:
            bias = bias.clone()
        if crelu:                                          # z = relu(concat([y, -y], 3)) straight from the convolution's epilogue
            y = torch.empty((B, H // s, W // s, 2 * cout), device=x.device, dtype=torch.float32)
            rc = lib.otgan_conv2d_fprop_crelu_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, x.data_ptr(), wt.data_ptr(),
                                                   bias.data_ptr() if bias is not None else None, y.data_ptr(), stream)
            _lib.check(rc, "otgan_conv2d_fprop_crelu_tf32")
        else:
            y = torch.empty((B, H // s, W // s, cout), device=x.device, dtype=torch.float32)
            ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_gemm(B, H // s, W // s, cout))
            rc = lib.otgan_conv2d_fprop_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, x.data_ptr(), wt.data_ptr(),
                                             bias.data_ptr() if bias is not None else None, y.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_fprop_tf32")
        ctx.geom, ctx.has_bias, ctx.cout = geom, bias is not None, cout
        ctx.wt, ctx.ihwo = wt, ihwo
        if crelu:
            ctx.save_for_backward(x, Vc, gc, inv, y)
        else:
            ctx.save_for_backward(x, Vc, gc, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, V, g, inv = ctx.saved_tensors[:4]
        kh, kw, s, pt, pl = ctx.geom
        B, H, W, cin = x.shape
        cout = ctx.cout
        dy = dy.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        if ctx.crelu:                                      # gradient of z = crelu(y) -> gradient of y
            dy = _crelu_bwd_from_z(lib, ctx.saved_tensors[4], dy, stream)
        dx = dV = dg = db = None
        if ctx.needs_input_grad[0]:
            wt_t = ctx.ihwo
            if wt_t is None:
                wt_t = torch.empty((cin, kh * kw * cout), device=x.device, dtype=torch.float32)
                _lib.check(lib.otgan_ohwi_to_ihwo_f32(cout, kh * kw, cin, ctx.wt.data_ptr(), wt_t.data_ptr(), stream), "otgan_ohwi_to_ihwo_f32")
            dx = torch.empty_like(x)
            ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_gemm(B, H, W, cin))
            rc = lib.otgan_conv2d_dgrad_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, dy.data_ptr(), wt_t.data_ptr(), dx.data_ptr(),
                                             ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_dgrad_tf32")
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            ws = _workspace(x.device, lib.otgan_workspace_bytes_conv_wgrad(B, H, W, cin, cout, kh, kw, s))
            dw = torch.empty((kh * kw * cin, cout), device=x.device, dtype=torch.float32)               # HWIO: dV's layout
            rc = lib.otgan_conv2d_wgrad_hwio_tf32(B, H, W, cin, cout, kh, kw, s, pt, pl, dy.data_ptr(), x.data_ptr(), dw.data_ptr(),
                                                  ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_wgrad_hwio_tf32")
            dV, dg = _wn_bwd_hwio(lib, V, g, inv, dw, stream)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            P = dy.numel() // cout
            ws = _workspace(x.device, lib.otgan_workspace_bytes_colsum(P, cout))
            db = torch.empty((cout,), device=x.device, dtype=torch.float32)
            _lib.check(lib.otgan_colsum_f32(P, cout, dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream), "otgan_colsum_f32")
        return dx, dV, dg, db, None, None, None


class _ConvUp2TCWN(torch.autograd.Function):
    """Weight norm + (2x nearest-neighbour upsample -> conv 'SAME') + bias as one node (generator, models/dcgan.py:37-46):
    V -> W -> 4 pre-summed sub-filters forward; backward in V's layout: sub-filter dgrad operand from W_ihwo, wgrad and the chain
    rule of the pre-sum on HWIO gradients, streaming weight-norm backward."""

    @staticmethod
    def forward(ctx, x_low, V, g, bias, geom):
        lib = _lib.load()
        kh, kw, pt, pl = geom
        stream = torch.cuda.current_stream().cuda_stream
        B, H, W, cin = x_low.shape
        cout = V.shape[-1]
        n1 = lib.otgan_up2_subtaps(kh, pt)
        slots = n1 * n1
        x_low, Vc, gc = x_low.contiguous(), V.contiguous(), g.contiguous()
        wt, ihwo, inv = _wn_fwd2(lib, Vc, gc, kh * kw, ctx.needs_input_grad[0], stream)
        if bias is not None and bias.data_ptr() % 16:
            bias = bias.clone()
        w_sub = torch.empty((4, cout, slots * cin), device=x_low.device, dtype=torch.float32)
        _lib.check(lib.otgan_up2_weight_presum_f32(cout, kh, kw, cin, pt, pl, wt.data_ptr(), w_sub.data_ptr(), stream), "otgan_up2_weight_presum_f32")
        y = torch.empty((B, 2 * H, 2 * W, cout), device=x_low.device, dtype=torch.float32)
        ws = _workspace(x_low.device, lib.otgan_workspace_bytes_conv_gemm(B, 2 * H, 2 * W, cout))
        rc = lib.otgan_conv2d_up2_fprop_tf32(B, H, W, cin, cout, kh, kw, pt, pl, x_low.data_ptr(), w_sub.data_ptr(),
                                             bias.data_ptr() if bias is not None else None, y.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
        _lib.check(rc, "otgan_conv2d_up2_fprop_tf32")
        ctx.geom, ctx.has_bias, ctx.cout, ctx.slots, ctx.ihwo = geom, bias is not None, cout, slots, ihwo
        ctx.save_for_backward(x_low, Vc, gc, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x_low, V, g, inv = ctx.saved_tensors
        kh, kw, pt, pl = ctx.geom
        B, H, W, cin = x_low.shape
        cout, slots = ctx.cout, ctx.slots
        dy = dy.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        dx = dV = dg = db = None
        if ctx.needs_input_grad[0]:
            w_sub_t = torch.empty((4, cin, slots * cout), device=dy.device, dtype=torch.float32)
            _lib.check(lib.otgan_up2_weight_presum_ihwo_f32(cout, kh, kw, cin, pt, pl, ctx.ihwo.data_ptr(), w_sub_t.data_ptr(), stream),
                       "otgan_up2_weight_presum_ihwo_f32")
            dx = torch.empty_like(x_low)
            ws = _workspace(dy.device, lib.otgan_workspace_bytes_conv_gemm(B, H, W, cin))
            rc = lib.otgan_conv2d_up2_dgrad_tf32(B, H, W, cin, cout, kh, kw, pt, pl, dy.data_ptr(), w_sub_t.data_ptr(), dx.data_ptr(),
                                                 ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_up2_dgrad_tf32")
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            ws = _workspace(dy.device, lib.otgan_workspace_bytes_conv_up2_wgrad(B, H, W, cin, cout, kh, kw, pt, pl))
            dw_sub = torch.empty((4, slots * cin, cout), device=dy.device, dtype=torch.float32)
            rc = lib.otgan_conv2d_up2_wgrad_hwio_tf32(B, H, W, cin, cout, kh, kw, pt, pl, dy.data_ptr(), x_low.data_ptr(), dw_sub.data_ptr(),
                                                      ws.data_ptr(), ws.numel() * 4, stream)
            _lib.check(rc, "otgan_conv2d_up2_wgrad_hwio_tf32")
            dw = torch.empty((kh * kw * cin, cout), device=dy.device, dtype=torch.float32)
            _lib.check(lib.otgan_up2_weight_unsum_hwio_f32(cout, kh, kw, cin, pt, pl, dw_sub.data_ptr(), dw.data_ptr(), stream),
                       "otgan_up2_weight_unsum_hwio_f32")
            dV, dg = _wn_bwd_hwio(lib, V, g, inv, dw, stream)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            P = dy.numel() // cout
            ws = _workspace(dy.device, lib.otgan_workspace_bytes_colsum(P, cout))
            db = torch.empty((cout,), device=dy.device, dtype=torch.float32)
            _lib.check(lib.otgan_colsum_f32(P, cout, dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream), "otgan_colsum_f32")
        return dx, dV, dg, db, None

# ------------------------------------------------------------------------------------------------ get_params
def get_params(layer_name, x=None, init=False, ema=None, use_W=True, use_g=True, use_b=True, f=None, weight_norm=True,
               init_scale=1.0, filter_size=None, num_units=None, pre_activation=None, raw=False):
    """utils/nn.py:103-183: variables V, g, b of one layer (created on the init pass), W = g * V / ||V||.
    raw=True returns the variables themselves ({"V", "g", "b"}, EMA-substituted) for kernels that fuse the weight norm."""
    store = _tls.store
    scope = store.name + "/" + layer_name
    ema_src = ema.shadow if ema is not None else None
    params = {}
    if raw:
        return {"V": store.get(scope + "/V", ema_src), "g": store.get(scope + "/g", ema_src), "b": store.get(scope + "/b", ema_src)}
    if init:
        xl = _as_list(x)
        nr_in = sum(int(xi.shape[-1]) for xi in xl)
        if num_units is None:
            num_units = nr_in
        if use_W:
            if pre_activation in ("celu", "crelu"):
                nr_in *= 2
            vshape = list(filter_size) + [nr_in, num_units] if filter_size is not None else [nr_in, num_units]
            store.create(scope + "/V", vshape, lambda s: torch.randn(s) * 0.05)                      # :123-127
        if use_g:
            store.create(scope + "/g", [num_units], lambda s: torch.ones(s))                          # :143
        if use_b:
            store.create(scope + "/b", [num_units], lambda s: torch.zeros(s))                         # :160
        if DATA_DEPENDENT_INIT and use_W:
            V = store.get(scope + "/V")
            Wn = l2_normalize(V, list(range(V.dim() - 1))) if weight_norm else V
            x_init = f(x, Wn)
            red = list(range(x_init.dim() - 1))
            m_init, v_init = x_init.mean(red), x_init.var(red, unbiased=False)                        # tf.nn.moments :138
            init_g = init_scale / torch.sqrt(v_init)
            if use_g:
                store.get(scope + "/g").copy_(init_g)
            if use_b:
                store.get(scope + "/b").copy_(-m_init * init_g)
    if use_b:
        params["b"] = store.get(scope + "/b", ema_src)
    g = store.get(scope + "/g", ema_src) if use_g else None
    if use_W:
        V = store.get(scope + "/V", ema_src)
        cached = store.cached_weight(scope) if (weight_norm and use_g and ema is None and store.frozen) else None
        if weight_norm and use_g and V.is_cuda and store.frozen and V.dim() == 4 and WN_FUSION and CONV_BACKEND == "tcgen05":
            params["W"] = LazyWN(V, g, cached)                                                        # :176-180 fused into the convolution
        elif cached is not None:
            params["W"] = TransposedWeight(cached[0], V.shape, cached[1])                             # reuse (no param grads)
        elif weight_norm and use_g and V.is_cuda and store.frozen:
            params["W"] = TransposedWeight(_WeightNorm.apply(V, g), V.shape)                          # :176-180 fused
        else:
            W = l2_normalize(V, list(range(V.dim() - 1))) if weight_norm else V                       # :176
            if use_g:
                W = W * g.view([1] * (V.dim() - 1) + [V.shape[-1]])                                   # :180
            params["W"] = W
    elif use_g:
        params["g"] = g
    return params


# ------------------------------------------------------------------------------------------------ layers
def _dense(x, W, pre_activation=None, bias=None):
    """tf.matmul(x, W) (+ bias when given: the caller's `x + b` rides in the GEMM epilogue on the tcgen05 path)."""
    x = apply_pre_activation(x, pre_activation, 1)
    if isinstance(W, TransposedWeight):
        if (CONV_BACKEND == "tcgen05" and DENSE_ON_TCGEN05 and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
                and x.shape[1] % 4 == 0 and W.wt.shape[0] % 4 == 0):
            # tf.matmul(x, W) (utils/nn.py:208-209) as a 1x1 convolution of a [B, 1, 1, K] image on the generic tcgen05 kernels:
            # W.wt [units, K] is exactly the OHWI filter matrix; the filter gradient is the same kernels' wgrad
            B, K = x.shape
            return _ConvGen.apply(x.contiguous().view(B, 1, 1, K), W.wt, bias, (1, 1, 1, 0, 0)).view(B, W.wt.shape[0])
        y = F.linear(x, W.wt)
        return y if bias is None else y + bias
    y = x @ W                                                                                         # :208-209
    return y if bias is None else y + bias


def _conv2d_fused_crelu(x, W, stride, pad, dilate, pre_activation, upsample, bias):
    """z = crelu(conv(x) + b) from the convolution's own epilogue, or None when this call has no such form."""
    if (CONV_BACKEND != "tcgen05" or upsample or dilate != 1 or pre_activation is not None or isinstance(x, (list, tuple, Upsampled2x))
            or not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4)):
        return None
    kh, kw, cin, cout = W.vshape if isinstance(W, (LazyWN, TransposedWeight)) else (0, 0, 0, 0)
    if not kh:
        return None
    B, H, Wd, _ = x.shape
    s = stride[0]
    if stride[0] != stride[1] or H % s or Wd % s or not _crelu_fusable(B, H // s, Wd // s, cout):
        return None
    geom = (kh, kw, s, same_padding(H, kh, s)[0], same_padding(Wd, kw, s)[0])
    if isinstance(W, LazyWN) and conv_tc_supported((B, H, Wd, cin), cout, kh, kw, stride, pad):
        return _ConvTCWN.apply(x.contiguous(), W.V, W.g, bias, geom, W.cached, True)
    if CONV_NARROW and cin <= 16 and conv_narrow_supported((B, H, Wd, cin), cout, kh, kw, stride, pad):
        Wm = W.materialize() if isinstance(W, LazyWN) else W
        return _ConvNarrow.apply(x.contiguous(), Wm.wt, bias, geom, True)
    return None


def _conv2d(x, W, stride=(1, 1), pad="SAME", dilate=1, pre_activation=None, upsample=False, bias=None, crelu_out=False):
    """utils/nn.py:234-275 (__list_conv2d): optional NN-upsample of the concatenated list, pre-activation, conv.
    crelu_out: return CreluOut(relu(concat([y, -y], 3))) -- the next layer's CReLU -- instead of y (fused into the epilogue of the
    tcgen05 convolution where the launch has a tile per SM; the literal ops otherwise)."""
    if isinstance(x, Crelu8Tensor):
        raise TypeError("a Crelu8Tensor is consumed by nn.conv2d(..., pre_activation='crelu') on the GPU path only")
    if isinstance(x, CreluOut):                            # already activated by the producing layer
        if pre_activation != "crelu" or upsample:
            raise ValueError("a CreluOut input needs pre_activation='crelu' (it IS the activated tensor)")
        x, pre_activation = x.z, None
    if crelu_out:
        y = _conv2d_fused_crelu(x, W, stride, pad, dilate, pre_activation, upsample, bias)
        if y is None:                                      # no fused form for this call: convolution, then the CReLU pass
            y = _conv2d(x, W, stride, pad, dilate, pre_activation, upsample, bias)
            if y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and y.shape[3] % 4 == 0:
                y = _CreluPad.apply(y.contiguous(), (0, 0, 0, 0))
            else:
                y = torch.relu(torch.cat([y, -y], 3))
        return CreluOut(y)
    if isinstance(x, Upsampled2x):
        if (pre_activation is None and not upsample and isinstance(W, (TransposedWeight, LazyWN)) and CONV_BACKEND == "tcgen05"
                and conv_up2_supported(tuple(x.low.shape), W.vshape[3], W.vshape[0], W.vshape[1], stride, pad)):
            kh, kw = W.vshape[0], W.vshape[1]
            geom_up = (kh, kw, (kh - 1) // 2, (kw - 1) // 2)
            if isinstance(W, LazyWN) and W.cached is None and W.vshape[3] % 4 == 0 and (kh * kw * W.vshape[2]) % 4 == 0:
                return _ConvUp2TCWN.apply(x.low, W.V, W.g, bias, geom_up)
            if isinstance(W, LazyWN):
                W = W.materialize()
            return _ConvUp2TC.apply(x.low, W.wt, bias, geom_up)
        x = x.materialize()
    if isinstance(W, LazyWN):
        xs_ = _as_list(x)
        x0 = xs_[0]
        fused_ok = (len(xs_) == 1 and not isinstance(x0, Upsampled2x) and x0.is_cuda and x0.dtype == torch.float32 and x0.dim() == 4
                    and dilate == 1 and not upsample and pre_activation in (None, "crelu") and CONV_BACKEND == "tcgen05")
        if fused_ok:
            kh, kw, cin, cout = W.vshape
            B, H, Wd, C = x0.shape
            if conv_tc_supported((B, H, Wd, cin), cout, kh, kw, stride, pad):
                z = x0.contiguous() if pre_activation is None else _CreluPad.apply(x0.contiguous(), (0, 0, 0, 0))
                geom = (kh, kw, stride[0], same_padding(H, kh, stride[0])[0], same_padding(Wd, kw, stride[1])[0])
                return _ConvTCWN.apply(z, W.V, W.g, bias, geom, W.cached)
        W = W.materialize()
    xl = _as_list(x)
    xl = [xi.materialize() if isinstance(xi, Upsampled2x) else xi for xi in xl]
    if dilate != 1:
        raise NotImplementedError("dilated conv is dead code in the reference (SURVEY 2.1 #16)")
    if upsample:
        xc = torch.cat(xl, 3) if len(xl) > 1 else xl[0]
        xl = [resize_nearest_neighbor(xc, [2 * xc.shape[1], 2 * xc.shape[2]])]
    on_gpu = len(xl) == 1 and xl[0].is_cuda and xl[0].dtype == torch.float32 and xl[0].dim() == 4
    if on_gpu and CONV_BACKEND == "tcgen05" and isinstance(W, TransposedWeight) and pre_activation in (None, "crelu"):
        kh, kw, cin, cout = W.vshape
        B, H, Wd, C = xl[0].shape
        if conv_tc_supported((B, H, Wd, cin), cout, kh, kw, stride, pad):
            # TensorFlow 'SAME' zero padding costs nothing here: the kernels' TMA boxes are zero-filled outside the tensor
            z = xl[0].contiguous() if pre_activation is None else _CreluPad.apply(xl[0].contiguous(), (0, 0, 0, 0))
            geom = (kh, kw, stride[0], same_padding(H, kh, stride[0])[0], same_padding(Wd, kw, stride[1])[0])
            return _ConvTC.apply(z, W.wt, bias, geom, W.wt_ihwo)
        if pre_activation is None and CONV_NARROW and conv_narrow_supported((B, H, Wd, cin), cout, kh, kw, stride, pad):
            geom = (kh, kw, 1, same_padding(H, kh, 1)[0], same_padding(Wd, kw, 1)[0])
            return _ConvNarrow.apply(xl[0].contiguous(), W.wt, bias, geom)
        if conv_gen_supported((B, H, Wd, cin), cout, kh, kw, stride, pad):
            # every other shape: the generic mode of the same kernels (channel counts padded to multiples of 4)
            z = xl[0].contiguous() if pre_activation is None else _CreluPad.apply(xl[0].contiguous(), (0, 0, 0, 0))
            return _conv_gen(z, W.wt, bias, kh, kw, stride[0], cin, cout)
    if (pre_activation == "crelu" and len(xl) == 1 and pad == "SAME" and xl[0].is_cuda and xl[0].dtype == torch.float32
            and xl[0].shape[3] % 4 == 0):
        # CReLU written straight into the TensorFlow-'SAME'-padded input of the convolution (one fused kernel)
        kh, kw = (W.vshape[0], W.vshape[1]) if isinstance(W, TransposedWeight) else (W.shape[0], W.shape[1])
        pt, pb = same_padding(xl[0].shape[1], kh, stride[0])
        pl, pr = same_padding(xl[0].shape[2], kw, stride[1])
        z = _CreluPad.apply(xl[0].contiguous(), (pt, pl, pb, pr))
        return _conv2d_nhwc(z, W, list(stride), "VALID", bias)
    return _conv2d_nhwc(apply_pre_activation(xl, pre_activation, 3), W, list(stride), pad, bias)


def _conv_gen(z, wt, bias, kh, kw, s, cin, cout):
    """_ConvGen with the channel counts padded to multiples of 4 (TMA strides are multiples of 16 bytes): zero input channels /
    zero filters change nothing; the padded output channels are sliced away."""
    B, H, Wd, _ = z.shape
    cin4, cout4 = -(-cin // 4) * 4, -(-cout // 4) * 4
    if cin4 != cin:
        z = F.pad(z, (0, cin4 - cin))
        wt = F.pad(wt.view(cout, kh * kw, cin), (0, cin4 - cin)).reshape(cout, -1)
    if cout4 != cout:
        wt = F.pad(wt, (0, 0, 0, cout4 - cout))
        bias = F.pad(bias, (0, cout4 - cout)) if bias is not None else None
    geom = (kh, kw, s, same_padding(H, kh, s)[0], same_padding(Wd, kw, s)[0])
    y = _ConvGen.apply(z, wt, bias, geom)
    return y[..., :cout] if cout4 != cout else y


def _conv2d_crelu8(x, V, g, bias, stride, pad):
    """nn.conv2d(list, ..., pre_activation='crelu') where the activated list is a Crelu8Tensor: the filter's input channels are
    permuted into crelu8 order while W = g V / ||V|| is built; the convolution reads the buffer as is."""
    kh, kw, c2, cout = V.shape
    assert c2 == 2 * x.raw_channels, "filter does not match the crelu8 buffer"
    B, H, Wd, _ = x.z.shape
    if not conv_gen_supported((B, H, Wd, c2), cout, kh, kw, stride, pad):
        raise _lib.OtganError("conv2d on a crelu8 buffer: shape not supported (needs 'SAME', stride 1/2, power-of-two extents)")
    cout4 = -(-cout // 4) * 4
    if cout4 != cout:                                  # e.g. the generator's 3-channel output layer
        V = F.pad(V, (0, cout4 - cout))
        g = F.pad(g, (0, cout4 - cout), value=1.0)
        bias = F.pad(bias, (0, cout4 - cout)) if bias is not None else None
    wt = _WeightNormPerm.apply(V, g, crelu8_perm(x.elem_ch, kh * kw, x.z.device))
    geom = (kh, kw, stride[0], same_padding(H, kh, stride[0])[0], same_padding(Wd, kw, stride[1])[0])
    y = _ConvGen.apply(x.z, wt, bias, geom)
    return y[..., :cout] if cout4 != cout else y


@add_arg_scope
def dense(x, num_units, pre_activation="celu", init_scale=1.0, counters={}, init=False, ema=None, weight_norm=True,
          use_b=True, use_g=True, **kwargs):
    """utils/nn.py:315-325"""
    layer_name = get_name("dense", counters)
    f = lambda x, W: _dense(x, W, pre_activation)
    params = get_params(layer_name, x, init, ema, use_W=True, use_g=use_g, use_b=use_b, f=f, weight_norm=weight_norm,
                        init_scale=init_scale, num_units=num_units, pre_activation=pre_activation)
    return _dense(x, params["W"], pre_activation, bias=params["b"] if use_b else None)      # x @ W + b  (:322-324)


@add_arg_scope
def conv2d(x, num_filters, pre_activation="celu", filter_size=[3, 3], stride=[1, 1], pad="SAME", dilate=1,
           upsample=False, init_scale=1.0, counters={}, init=False, ema=None, weight_norm=True, use_b=True, use_g=True,
           crelu_out=False, **kwargs):
    """utils/nn.py:328-338.  crelu_out (addition): hand the NEXT layer's CReLU to this convolution -- returns a CreluOut."""
    layer_name = get_name("conv2d", counters)
    if isinstance(x, Crelu8Tensor):
        if not (pre_activation == "crelu" and weight_norm and use_g and not upsample and dilate == 1 and not init):
            raise ValueError("a Crelu8Tensor input needs pre_activation='crelu' with weight norm (models/densenet.py usage)")
        prm = get_params(layer_name, None, False, ema, raw=True)
        return _conv2d_crelu8(x, prm["V"], prm["g"], prm["b"] if use_b else None, stride, pad)
    f = lambda x, W: _conv2d(x, W, stride, pad, dilate, pre_activation, upsample)
    params = get_params(layer_name, x, init, ema, use_W=True, use_g=use_g, use_b=use_b, f=f, weight_norm=weight_norm,
                        init_scale=init_scale, filter_size=list(filter_size), num_units=num_filters,
                        pre_activation=pre_activation)
    # tf.nn.bias_add (:337) rides in the convolution's epilogue instead of a separate pass over the activations
    return _conv2d(x, params["W"], stride, pad, dilate, pre_activation, upsample, bias=params["b"] if use_b else None,
                   crelu_out=crelu_out and not init)


# ------------------------------------------------------------------------------------------------ optimisers
class _Updates:
    """What tf.group(*updates) is to sess.run: call .run(grads, lr) to apply one update."""

    def __init__(self, fn):
        self.run = fn


def _flat_of(params):
    """params: a Template (preferred) or a flat leaf tensor."""
    return params.flat if isinstance(params, Template) else params


def adam_updates(params, cost_or_grads=None, lr=0.001, mom1=0.9, mom2=0.999, ema=None, shard=None):
    """utils/nn.py:50-73 on the flat variable buffer, as ONE fused CUDA kernel (otgan_adam_ema_f32):
        v = mom1 v + (1-mom1) g ; v_hat = v / (1 - mom1^t) ; mg = mom2 mg + (1-mom2) g^2 ; mg_hat = mg / (1 - mom2^t)
        p = p - lr * v_hat / sqrt(mg_hat + 1e-8)        (epsilon INSIDE the root, t starts at 1)
    and, when `ema` is given, shadow = decay*shadow + (1-decay)*p in the same pass (train.py:64,223).
    Returns an object whose .run(grads_flat, lr=None) applies one step (lr may be negative: the critic ascends, train.py:143).
    shard = (rank, world): optimiser state sharded over the data-parallel ranks (ZeRO-1): this rank keeps the Adam moments of, and
    updates, only its 1/world slice of the flat buffer; .run then expects the REDUCE-SCATTERED gradient slice and the caller
    all-gathers the updated parameters (train.GradSync / Trainer): same NVLink volume as the all-reduce it replaces, the 7-pass Adam
    + EMA kernel at 1/world of the parameters.  The EMA shadow is updated on the own slice only (Trainer.sync_ema gathers it)."""
    flat = _flat_of(params)
    lo, hi = 0, flat.numel()
    if shard is not None:
        rank, world = shard
        assert flat.numel() % (4 * world) == 0, "flat parameter buffer must split into float4-sized shards"
        n = flat.numel() // world
        lo, hi = rank * n, (rank + 1) * n
    nloc = hi - lo
    state = {"t": 1, "mg": torch.zeros(nloc, device=flat.device), "v": torch.zeros(nloc, device=flat.device) if mom1 > 0 else None,
             "range": (lo, hi)}
    default_lr = lr

    def hyper(lr=None):
        """(lr, d1, d2) of the NEXT update -- the step-dependent scalars of utils/nn.py:63,68 -- and advance t.  Used with
        run(..., hyper_dev=...) when the update is replayed from a CUDA graph."""
        t = state["t"]
        state["t"] = t + 1
        return (float(default_lr if lr is None else lr), (1.0 - mom1 ** t) if mom1 > 0 else 1.0, 1.0 - mom2 ** t)

    def run(grads, lr=None, hyper_dev=None):
        lib = _lib.load()
        step_lr = default_lr if lr is None else lr
        g = grads if grads is not None else flat.grad
        if isinstance(params, Template):
            params.store.version += 1                       # cached W = g V/||V|| of this template is stale from here on
        if ema is not None and ema.shadow is None:
            ema.attach(params)
        stream = torch.cuda.current_stream().cuda_stream
        v_ptr = state["v"].data_ptr() if state["v"] is not None else None
        e_ptr = (ema.shadow.data_ptr() + 4 * lo) if ema is not None else None
        decay = float(ema.decay) if ema is not None else 0.0
        p_ptr = flat.data_ptr() + 4 * lo
        assert g.numel() == nloc, "adam_updates: gradient size does not match the (sharded) parameter range"
        if hyper_dev is not None:                           # scalars come from device memory; t is advanced by hyper()
            with torch.no_grad():
                rc = lib.otgan_adam_ema_dev_f32(nloc, p_ptr, g.data_ptr(), v_ptr, state["mg"].data_ptr(), e_ptr,
                                                hyper_dev.data_ptr(), float(mom1), float(mom2), decay, stream)
            _lib.check(rc, "otgan_adam_ema_dev_f32")
            return
        t = state["t"]
        c1 = (1.0 - mom1 ** t) if mom1 > 0 else 1.0        # bias-correction denominators (utils/nn.py:63,68)
        c2 = 1.0 - mom2 ** t
        with torch.no_grad():
            rc = lib.otgan_adam_ema_f32(nloc, p_ptr, g.data_ptr(), v_ptr, state["mg"].data_ptr(), e_ptr,
                                        float(step_lr), float(mom1), float(mom2), float(c1), float(c2), decay, stream)
        _lib.check(rc, "otgan_adam_ema_f32")
        state["t"] = t + 1

    u = _Updates(run)
    u.state = state
    u.hyper = hyper
    return u


def adamax_updates(params, cost_or_grads=None, lr=0.001, mom1=0.9, mom2=0.999):
    """utils/nn.py:29-48 (torch ops; not on the benchmarked path)."""
    flat = _flat_of(params)
    state = {"mg": torch.zeros_like(flat), "v": torch.zeros_like(flat) if mom1 > 0 else None}
    default_lr = lr

    def run(grads, lr=None):
        step_lr = default_lr if lr is None else lr
        g = grads if grads is not None else flat.grad
        if isinstance(params, Template):
            params.store.version += 1
        with torch.no_grad():
            if mom1 > 0:
                state["v"].mul_(mom1).add_(g, alpha=1.0 - mom1)
                v_t = state["v"]
            else:
                v_t = g
            state["mg"] = torch.maximum(mom2 * state["mg"] + 1e-8, g.abs())
            flat.sub_(step_lr * v_t / state["mg"])

    return _Updates(run)


def nesterov_updates(params, cost_or_grads=None, lr=0.01, mom1=0.9):
    """utils/nn.py:75-87 (torch ops; not on the benchmarked path)."""
    flat = _flat_of(params)
    state = {"v": torch.zeros_like(flat)}
    default_lr = lr

    def run(grads, lr=None):
        step_lr = default_lr if lr is None else lr
        g = grads if grads is not None else flat.grad
        if isinstance(params, Template):
            params.store.version += 1
        with torch.no_grad():
            v_new = mom1 * state["v"] - step_lr * g
            flat.add_(-mom1 * state["v"] + (1.0 + mom1) * v_new)
            state["v"] = v_new

    return _Updates(run)
