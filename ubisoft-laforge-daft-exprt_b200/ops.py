"""Autograd functions of the Daft-Exprt hot path: every forward/backward here is a sequence of C-ABI calls into
libdaftexprt_b200.so (hand-written sm_100a CUDA).  PyTorch provides device memory, streams and the autograd tape only.

Granularity = one Function per sub-layer of the reference (`model.py`), so that the tape has ~40 nodes per step and
the fusions live below the ABI:
  AttentionSubLayer   MultiHeadAttention.forward + masked_fill            model.py:171-193, :258-259
  ConvFFSubLayer      PositionWiseConvFF.forward + masked_fill            model.py:220-237, :261-262
  PreNet              ProsodyEncoder.convs                                model.py:341-363, :408-409
  FrameInput          + energy/pitch embeddings + PE + mask               model.py:400-414 / :696-701
  MeanPoolSpeaker     time pooling + speaker embedding                    model.py:419-424
  FilmHead            gammas/betas predictors + post-multipliers          model.py:427-461
  Linear              LinearNorm (+ReLU), gradient reversal folded in     model.py:27-38, :276-283
  EmbedPE             symbols embedding + PE + mask                       model.py:497-504
  PredictorBlock/Head LocalProsodyPredictor.forward (any nb_blocks)       model.py:549-575
  GaussUpsample       GaussianUpsamplingModule.forward                    model.py:608-662
  MelProjection       Linear(128->80) + mask + transpose                  model.py:706-708
  Loss                DaftExprtLoss.forward                               loss.py:30-106
"""
import ctypes
import threading
import struct
import weakref

import torch

from . import cabi

_state = threading.local()
_weights_epoch = 0
_pack_cache = {}
_desc_tables = {}
_seed_counter = [0]
_backend = [cabi.DX_GEMM_TCGEN05_BF16X3]   # same default as the library (cabi.cu: g_backend)
launch_count = [0]   # number of C-ABI compute calls issued (each launches >= 1 of our kernels)


def lib():
    return cabi.load()


BACKENDS = {'fp32': cabi.DX_GEMM_FP32_CUDA_CORES, 'tf32': cabi.DX_GEMM_TCGEN05_TF32, 'bf16x3': cabi.DX_GEMM_TCGEN05_BF16X3}


def set_backend(name):
    """GEMM backend for Conv1d/Linear forward, dgrad and wgrad:
    'bf16x3' tcgen05 tensor cores on bf16 hi/lo operand planes, 3 passes, fp32-grade results (default; weight gradients summed
             over >= 4096 rows take one pass, see set_gemm_passes);
    'tf32'   tcgen05 kind::tf32 on fp32 tiles, one pass, ~1e-3 per GEMM (wgrad still runs bf16x3);
    'fp32'   exact fp32 on CUDA cores (strict parity mode)."""
    be = BACKENDS[name]
    cabi.check(lib().dx_set_gemm_backend(be), 'dx_set_gemm_backend')
    _backend[0] = be
    invalidate_packed_weights()


def set_gemm_passes(conv=3, wgrad=0):
    """bf16x3 backend: tensor-core passes per K-step of the GEMMs issued from now on (3 = parity mode; 2 / 1 = reduced precision;
    wgrad 0 = the default: one pass for weight gradients summed over >= 4096 rows, three below — see dx_set_gemm_passes).  `conv` applies to forward AND input-gradient GEMMs: switch it between a forward and its backward to
    give them different pass counts (tools/pass_ablation.py does)."""
    cabi.check(lib().dx_set_gemm_passes(int(conv), int(wgrad)), 'dx_set_gemm_passes')


def get_backend():
    return {v: k for k, v in BACKENDS.items()}[_backend[0]]


def invalidate_packed_weights():
    """Call after parameters were modified through raw pointers (e.g. dx_adam_step), which does not bump tensor versions."""
    global _weights_epoch
    _weights_epoch += 1
    for k in [k for k, hit in _pack_cache.items() if hit[3]() is None]:   # the pack buffers of live weights are kept and refilled
        del _pack_cache[k]
    if len(_desc_tables) > 16:
        _desc_tables.clear()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _call(name, *args):
    launch_count[0] += 1
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed (rc={rc}): {cabi.last_error()}')


def next_seed():
    _seed_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


def _check_input(t, dtype=torch.float32):
    if t.device.type != 'cuda':
        raise RuntimeError('daft_exprt_b200 ops need CUDA tensors: there is no CPU fallback')
    if t.dtype != dtype:
        raise RuntimeError(f'expected {dtype}, got {t.dtype}')
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------------------
# packed weights (derived, cached, version-checked; the parameters keep the reference's own layout)
# ----------------------------------------------------------------------------------------------------------------------
class PackedWeight:
    """fp32 packed layout [KW, N, K] + (bf16x3 backend) cached bf16 hi|lo planes of the same layout."""
    __slots__ = ('w', 'planes', 'shape', 'gen')

    def __init__(self, w, planes):
        self.w, self.planes, self.shape = w, planes, tuple(w.shape)
        self.gen = 0   # bumped every time the buffers are refilled in place with new weight values


def hold_packs(ctx, *packs):
    """Remember the dgrad packs a backward will consume together with their fill generation."""
    ctx.packed = packs if len(packs) != 1 else packs[0]
    ctx.packed_gen = tuple(p.gen for p in packs)


def held_packs(ctx):
    """The packs saved by `hold_packs`; raises when one of them was refilled with NEWER weights since the forward (forward(A),
    optimiser step, forward(B), backward(A) would otherwise silently use the new weights for A's dgrad)."""
    packs = ctx.packed if isinstance(ctx.packed, (tuple, list)) else (ctx.packed,)
    for p, g in zip(packs, ctx.packed_gen):
        if p.gen != g:
            raise RuntimeError('daft_exprt_b200: a packed weight was refilled (the parameters changed) between this forward and '
                               'its backward; run backward before the next optimiser step + forward of the same model')
    return ctx.packed


def packed(weight):
    """weight [Cout, Cin, KW] or [Cout, Cin] -> (fwd [KW, Cout, Cin], dgrad [KW, Cin, Cout] taps flipped) PackedWeights;
    one fused launch makes both fp32 packs and (bf16x3 backend) their bf16 hi|lo planes."""
    w = weight.detach()
    key = id(weight)
    ver = (w.data_ptr(), w._version, _weights_epoch, _backend[0])
    hit = _pack_cache.get(key)
    if hit is not None and hit[3]() is not weight:   # id() can be recycled: check object identity too
        hit = None
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    w = _check_input(w)
    cout, cin = w.shape[0], w.shape[1]
    kw = w.shape[2] if w.dim() == 3 else 1
    rnd = 1 if _backend[0] == cabi.DX_GEMM_TCGEN05_TF32 else 0
    if hit is not None and hit[0][3] == ver[3] and hit[1].w.device == w.device:
        fwd, dgrad = hit[1], hit[2]                  # same weight, same backend: refill the existing buffers
        _call('dx_pack_conv_weight', _p(w), _p(fwd.w), _p(dgrad.w), _p(fwd.planes), _p(dgrad.planes), cout, cin, kw, rnd, _st())
        fwd.gen += 1
        dgrad.gen += 1
    else:
        fw = torch.empty(kw, cout, cin, device=w.device, dtype=torch.float32)
        dg = torch.empty(kw, cin, cout, device=w.device, dtype=torch.float32)
        fp = dp = None
        if _backend[0] == cabi.DX_GEMM_TCGEN05_BF16X3:
            fp = torch.empty(2 * fw.numel(), device=w.device, dtype=torch.bfloat16)
            dp = torch.empty(2 * fw.numel(), device=w.device, dtype=torch.bfloat16)
        _call('dx_pack_conv_weight', _p(w), _p(fw), _p(dg), _p(fp), _p(dp), cout, cin, kw, rnd, _st())
        fwd, dgrad = PackedWeight(fw, fp), PackedWeight(dg, dp)
    _pack_cache[key] = (ver, fwd, dgrad, weakref.ref(weight))
    return fwd, dgrad


def prepack(weights):
    """Refresh, with ONE launch, every stale pack among `weights` that `packed()` has built before (an optimiser step makes
    all ~60 of them stale at once).  Weights never seen by `packed()`, or whose storage moved, are left to `packed()`.
    The pack buffers are refilled IN PLACE: a forward of a model invalidates the packs held by an older, not yet
    back-propagated forward of the same model if the weights changed in between."""
    stale = []
    for wt in weights:
        hit = _pack_cache.get(id(wt))
        if hit is None or hit[3]() is not wt:
            continue
        w = wt.detach()
        ver = (w.data_ptr(), w._version, _weights_epoch, _backend[0])
        if hit[0] == ver or hit[0][0] != ver[0] or hit[0][3] != ver[3] or not w.is_contiguous() or hit[1].shape[0] > 4:
            continue
        stale.append((wt, w, ver, hit))
    if len(stale) < 2:
        return
    key = tuple((w.data_ptr(), hit[1].w.data_ptr()) for _, w, _, hit in stale)
    table = _desc_tables.get(key)
    if table is None:
        raw, block0 = bytearray(), 0
        for _, w, _, hit in stale:
            kw, cout, cin = hit[1].shape
            raw += struct.pack('<5Q4i8x', w.data_ptr(), _p(hit[1].w) or 0, _p(hit[2].w) or 0, _p(hit[1].planes) or 0,
                               _p(hit[2].planes) or 0, cout, cin, kw, block0)
            block0 += ((cout + 31) // 32) * ((cin + 31) // 32)
        dev_table = torch.frombuffer(raw, dtype=torch.uint8).to(stale[0][1].device)
        table = (dev_table, block0)
        _desc_tables[key] = table
    rnd = 1 if _backend[0] == cabi.DX_GEMM_TCGEN05_TF32 else 0
    _call('dx_pack_conv_weights_batched', _p(table[0]), len(stale), table[1], rnd, _st())
    for wt, _, ver, hit in stale:
        hit[1].gen += 1
        hit[2].gen += 1
        _pack_cache[id(wt)] = (ver, hit[1], hit[2], hit[3])


def make_planes(x, rows, C, ld=None, want_colsum=False):
    """bf16 hi|lo operand planes of an fp32 activation [rows, C] (made ONCE, shared by every tensor-core GEMM that consumes
    the activation: forward + wgrad, or dgrad + wgrad).  None when the active backend does not use planes.
    want_colsum: also return the column sums of x (the bias gradient when x is an output gradient), fused into the same pass."""
    if _backend[0] == cabi.DX_GEMM_FP32_CUDA_CORES or C % 8 != 0 or C < 16:
        return (None, None) if want_colsum else None
    planes = torch.empty(2, rows, C, device=x.device, dtype=torch.bfloat16)
    cs = torch.empty(C, device=x.device, dtype=torch.float32) if want_colsum else None
    _call('dx_split_planes', _p(x), C if ld is None else ld, _p(planes), _p(cs), rows, C, _st())
    return (planes, cs) if want_colsum else planes


def conv_gemm(x, wp, bias, B, S, relu=False, relu_src=None, add_src=None, alpha=1.0, round_out=False, ldx=None, x_planes=None,
              lens=None, halo=0, relu_src_hi=None, emit_planes=False, want_y=True, want_colsum=False):
    """x [B,S,Cin] (row stride ldx) · packed weight [KW,Cout,Cin] -> [B,S,Cout].
    lens/halo: padding skip — output rows s >= lens[b] + halo are declared irrelevant by the caller (written as zeros).
    Plane hand-over (`plane_handover()`): emit_planes -> returns (y, planes [2, B*S, Cout] bf16 hi|lo written by the epilogue);
    want_y=False drops the fp32 copy; x may be None when x_planes is given; relu_src_hi = hi plane used as the ReLU mask.
    want_colsum (tensor-core backends): also returns the [Cout] column sums of the output, accumulated by the epilogue."""
    kw, cout, cin = wp.shape
    dev = wp.w.device
    y = torch.empty(B, S, cout, device=dev, dtype=torch.float32) if want_y else None
    yP = torch.empty(2, B * S, cout, device=dev, dtype=torch.bfloat16) if emit_planes else None
    ycs = torch.empty(cout, device=dev, dtype=torch.float32) if want_colsum else None
    rnd = 1 if (round_out and _backend[0] == cabi.DX_GEMM_TCGEN05_TF32) else 0
    nbytes = lib().dx_conv_gemm_workspace(B, S, cin, cout, kw, int(x_planes is not None), int(wp.planes is not None), -1)
    ws = torch.empty(nbytes + 256, device=dev, dtype=torch.uint8) if nbytes else None
    _call('dx_conv_gemm', _p(x), _p(x_planes), _p(wp.w), _p(wp.planes), _p(bias), _p(relu_src), _p(relu_src_hi), _p(add_src), _p(y),
          _p(yP), _p(ycs), _p(ws), ws.numel() if ws is not None else 0, _p(lens), int(halo), B, S, cin, cout, kw,
          cin if ldx is None else ldx, cout, float(alpha), int(relu), rnd, -1, _st())
    out = (y, yP) if emit_planes else y
    if want_colsum:
        out = (out + (ycs,)) if emit_planes else (out, ycs)
    return out


_wgrad_deferred = [False]
_wgrad_keepalive = []


def set_wgrad_deferral(on):
    """Opt-in (gather-mode gradient buckets only: nobody may read a weight gradient before `flush_wgrad()`): the ~58 split-K
    reductions of a backward pass are recorded and executed by ONE launch."""
    _call('dx_wgrad_defer', int(bool(on)))
    _wgrad_deferred[0] = bool(on)
    _wgrad_keepalive.clear()


def flush_wgrad():
    if _wgrad_deferred[0]:
        _call('dx_wgrad_flush', _st())
        _wgrad_keepalive.clear()


def plane_handover(cin, cout):
    """True when a GEMM [.., cin] -> [.., cout] can hand its output to the next GEMM as bf16 hi|lo planes (no fp32 copy)."""
    return _backend[0] == cabi.DX_GEMM_TCGEN05_BF16X3 and cin % 8 == 0 and cin >= 16 and cout % 32 == 0


def colsum_planes(planes, rows, C):
    db = torch.empty(C, device=planes.device, dtype=torch.float32)
    _call('dx_colsum_planes', _p(planes), _p(db), rows, C, _st())
    return db


def linear_rows(x2d, wp, bias, **kw):
    """Linear over a [R, Cin] matrix (KW == 1: rows are independent, so the batch structure is irrelevant)."""
    R = x2d.shape[0]
    return conv_gemm(x2d, wp, bias, 1, R, **kw).view(R, -1)


_grad_slots = {}          # id(parameter) -> view of the flat gradient bucket (ddp.FlatGradSync, mode 'gather')
_use_grad_slots = [False]


def register_grad_slots(slots):
    """ddp.FlatGradSync: where each parameter's gradient lives in the flat bucket.  While `use_grad_slots(True)`, the weight-gradient
    GEMMs of the big weights write THERE (their result tensor is a fresh alias of the slot, which autograd's AccumulateGrad adopts as
    `p.grad`), so packing the bucket before the gradient exchange has nothing left to copy for them."""
    _grad_slots.clear()
    _grad_slots.update(slots)


def use_grad_slots(on):
    _use_grad_slots[0] = bool(on)


def grad_slot(w):
    return _grad_slots.get(id(w)) if _use_grad_slots[0] else None


def conv_wgrad(x, dy, B, S, cin, cout, kw, shape, want_bias=True, ldx=None, alpha=1.0, x_planes=None, dy_planes=None, dbias=None,
               lens=None, halo=0, slot=None):
    """-> (dw in the parameter's own layout `shape`, dbias [Cout] or None).  `dbias`: already computed (fused into make_planes).
    slot: the parameter's view of the flat gradient bucket (grad_slot()): dw is written there."""
    dev = (dy if dy is not None else dy_planes).device   # x / dy may be None when their planes are handed over
    if slot is not None and tuple(slot.shape) == tuple(shape) and slot.is_contiguous():
        dw = slot.detach()        # a fresh alias: AccumulateGrad adopts it without a copy when nothing else refers to it
    else:
        dw = torch.empty(tuple(shape), device=dev, dtype=torch.float32)
    if dbias is not None:
        want_bias = False
    db = torch.empty(cout, device=dev, dtype=torch.float32) if want_bias else None
    nbytes = lib().dx_conv_wgrad_workspace(B, S, cin, cout, kw, int(x_planes is not None), int(dy_planes is not None), -1)
    ws = torch.empty(max(nbytes, 16) // 4 + 4, device=dev, dtype=torch.float32)
    _call('dx_conv_wgrad', _p(x), _p(x_planes), _p(dy), _p(dy_planes), _p(dw), _p(db), _p(ws), ws.numel() * 4, _p(lens), int(halo),
          B, S, cin, cout, kw, cin if ldx is None else ldx, float(alpha), -1, _st())
    if _wgrad_deferred[0]:
        _wgrad_keepalive.append(ws)   # the split-K partials are reduced by flush_wgrad()
    return dw, (dbias if dbias is not None else db)


_fused_inproj = [True]
_fused_ln = [True]


def fused_ln(on=None):
    """Get / set: the GEMMs that end a sub-layer (attention out-projection, conv2 of the feed-forward) finish dropout + residual +
    LayerNorm + FiLM + mask in their own epilogue (default on; off = separate dx_ln_fwd launch, kept for A/B tests)."""
    if on is not None:
        _fused_ln[0] = bool(on)
    return _fused_ln[0] and _backend[0] == cabi.DX_GEMM_TCGEN05_BF16X3


def gemm_ln(x_planes, wp, bias, res, ln_w, ln_b, film, film_stride, lens, B, S, p_in=0.0, seed_in=0):
    """(y, xhat, rstd) of dx_conv_gemm_ln; y carries its operand planes (attach_planes)."""
    kw, cout, cin = wp.shape
    dev = wp.w.device
    y = torch.empty(B, S, cout, device=dev, dtype=torch.float32)
    xhat = torch.empty(B, S, cout, device=dev, dtype=torch.float32)
    rstd = torch.empty(B * S, device=dev, dtype=torch.float32)
    yP = torch.empty(2, B * S, cout, device=dev, dtype=torch.bfloat16)
    _call('dx_conv_gemm_ln', _p(x_planes), _p(wp.planes), _p(bias), _p(res), _p(ln_w), _p(ln_b), _p(film), film_stride, _p(lens), _p(y),
          _p(yP), _p(xhat), _p(rstd), B, S, cin, kw, float(p_in), seed_in, _st())
    attach_planes(y, yP)
    return y, xhat, rstd


def can_fuse_ln(x_planes, wp, bias, D):
    return fused_ln() and D == 128 and x_planes is not None and wp.planes is not None and bias is not None and wp.shape[0] in (1, 3)


def fused_inproj(on=None):
    """Get / set: in-projection GEMM writes the attention operand planes directly (default on; off = fp32 qkv + conversion pass,
    kept for A/B tests)."""
    if on is not None:
        _fused_inproj[0] = bool(on)
    return _fused_inproj[0] and _backend[0] == cabi.DX_GEMM_TCGEN05_BF16X3


def attention_planes(B, S, H, dh, device):
    """Workspace for the tensor-core attention's operand planes; None when dx_attention_fwd will run the exact-fp32 kernels
    (fp32 backend, or a head layout the tensor-core kernels do not cover), which never write planes."""
    if not lib().dx_attention_uses_planes(H, dh):
        return None
    return torch.empty(lib().dx_attention_planes_bytes(B, S, H, dh), device=device, dtype=torch.uint8)


def _uses_planes(C):
    return _backend[0] != cabi.DX_GEMM_FP32_CUDA_CORES and C % 8 == 0 and C >= 16


def attach_planes(t, planes):
    """Side channel between sub-layers: the producer of an activation already wrote its bf16 hi|lo operand planes."""
    if planes is not None:
        t._dx_planes = (planes, t._version)
    return t


def planes_of(x, rows, C):
    """Operand planes of x: the ones its producer attached (LayerNorm / attention epilogues), else one split pass."""
    tag = getattr(x, '_dx_planes', None)
    if tag is not None and tag[1] == x._version and tuple(tag[0].shape) == (2, rows, C) and _uses_planes(C):
        return tag[0]
    return make_planes(x, rows, C)


def ln_fwd(a, res, ln_w, ln_b, film, film_stride, lens, B, S, D, p_in=0.0, seed_in=0, p_out=0.0, seed_out=0, emit_planes=False):
    """emit_planes: y is also written as bf16 hi|lo operand planes (attached to y for `planes_of`)."""
    y = torch.empty(B, S, D, device=a.device, dtype=torch.float32)
    xhat = torch.empty(B, S, D, device=a.device, dtype=torch.float32)
    rstd = torch.empty(B * S, device=a.device, dtype=torch.float32)
    yP = torch.empty(2, B * S, D, device=a.device, dtype=torch.bfloat16) if (emit_planes and _uses_planes(D)) else None
    _call('dx_ln_fwd', _p(a), _p(res), _p(ln_w), _p(ln_b), _p(film), film_stride, _p(lens), _p(y), _p(xhat), _p(rstd), _p(yP),
          B, S, D, float(p_in), seed_in, float(p_out), seed_out, _st())
    attach_planes(y, yP)
    return y, xhat, rstd


def ln_bwd(dy, xhat, rstd, ln_w, ln_b, film, film_stride, lens, B, S, D, relu_src=None, p_in=0.0, seed_in=0, p_out=0.0,
           seed_out=0, want_film=False, emit_planes=False):
    """-> (dv, da, dln_w, dln_b, dfilm[, planes, colsum]).  emit_planes (D in {128, 256}): the fused kernel also writes the
    operand planes and the column sums of da (the gradient entering the GEMM that produced `a`); (None, None) when the
    backend has no planes."""
    dv = torch.empty(B, S, D, device=dy.device, dtype=torch.float32)
    da = torch.empty(B, S, D, device=dy.device, dtype=torch.float32) if p_in > 0 else None
    dw = torch.empty(D, device=dy.device, dtype=torch.float32)
    db = torch.empty(D, device=dy.device, dtype=torch.float32)
    dfilm = torch.empty(B, 2 * D, device=dy.device, dtype=torch.float32) if want_film else None
    fuse = emit_planes and _uses_planes(D) and D in (128, 256, 1024)
    gP = torch.empty(2, B * S, D, device=dy.device, dtype=torch.bfloat16) if fuse else None
    gcs = torch.empty(D, device=dy.device, dtype=torch.float32) if fuse else None
    _call('dx_ln_bwd', _p(dy), _p(xhat), _p(rstd), _p(ln_w), _p(ln_b), _p(film), film_stride, _p(lens), _p(relu_src),
          _p(dv), _p(da), _p(dw), _p(db), _p(dfilm), _p(gP), _p(gcs), B, S, D, float(p_in), seed_in, float(p_out), seed_out, _st())
    out = (dv, (da if da is not None else dv), dw, db, dfilm)
    if emit_planes:
        if not fuse:
            gP, gcs = make_planes(out[1], B * S, D, want_colsum=True)
        out = out + (gP, gcs)
    return out


# ----------------------------------------------------------------------------------------------------------------------
class AttentionSubLayer(torch.autograd.Function):
    """y = mask(LN(dropout(out_proj(SDPA(in_proj(x)))) + x))"""

    @staticmethod
    def forward(ctx, x, lens, in_w, in_b, out_w, out_b, ln_w, ln_b, nb_heads, p_drop):
        x = _check_input(x)
        B, S, D = x.shape
        dh = D // nb_heads
        in_wp, in_wd = packed(in_w)
        out_wp, out_wd = packed(out_w)
        xP = planes_of(x, B * S, D)
        att = torch.empty(B, S, D, device=x.device, dtype=torch.float32)
        lse = torch.empty(B, nb_heads, S, device=x.device, dtype=torch.float32)
        seed_attn, seed_out = (next_seed(), next_seed()) if p_drop > 0 else (0, 0)
        planes = attention_planes(B, S, nb_heads, dh, x.device)
        attP = torch.empty(2, B * S, D, device=x.device, dtype=torch.bfloat16) if (planes is not None and _uses_planes(D)) else None
        if fused_inproj() and planes is not None and xP is not None and in_wp.planes is not None:
            # the in-projection's epilogue writes the attention operand planes itself (per-head bf16 hi|lo, q pre-scaled): no fp32
            # qkv tensor, no conversion pass (north_star: QKV projection fused into the attention operand hand-over)
            qkv = None
            _call('dx_inproj_head_planes', _p(xP), _p(in_wp.planes), _p(in_b), _p(planes), _p(lens), B, S, D, nb_heads, dh, _st())
        else:
            qkv = conv_gemm(x, in_wp, in_b, B, S, x_planes=xP, lens=lens)   # rows >= len: keys masked, queries skipped
        _call('dx_attention_fwd', _p(qkv), _p(lens), _p(att), _p(lse), _p(planes), _p(attP), B, S, nb_heads, dh, float(p_drop), seed_attn,
              _st())   # the kernel writes ctx and its operand planes
        if can_fuse_ln(attP, out_wp, out_b, D):   # out-projection + dropout + residual + LayerNorm + mask in ONE kernel
            y, xhat, rstd = gemm_ln(attP, out_wp, out_b, x, ln_w, ln_b, None, 0, lens, B, S, p_in=p_drop, seed_in=seed_out)
        else:
            proj = conv_gemm(att, out_wp, out_b, B, S, x_planes=attP, lens=lens)   # rows >= len are masked by the LayerNorm kernel
            y, xhat, rstd = ln_fwd(proj, x, ln_w, ln_b, None, 0, lens, B, S, D, p_in=p_drop, seed_in=seed_out, emit_planes=True)
        ctx.save_for_backward(x, lens, qkv, att, lse, xhat, rstd, ln_w, ln_b, xP, attP, planes)
        hold_packs(ctx, in_wd, out_wd)
        ctx.slots = (grad_slot(in_w), grad_slot(out_w))
        ctx.cfg = (B, S, D, nb_heads, dh, float(p_drop), seed_attn, seed_out, in_w.shape, out_w.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, lens, qkv, att, lse, xhat, rstd, ln_w, ln_b, xP, attP, planes = ctx.saved_tensors
        in_wd, out_wd = held_packs(ctx)
        B, S, D, H, dh, p, seed_attn, seed_out, in_shape, out_shape = ctx.cfg
        dy = _check_input(dy)
        dv, dproj, dln_w, dln_b, _, dprojP, dpb = ln_bwd(dy, xhat, rstd, ln_w, ln_b, None, 0, lens, B, S, D, p_in=p, seed_in=seed_out,
                                                         emit_planes=True)
        scratch = torch.empty(lib().dx_attention_bwd_scratch_bytes(B, S, H, dh), device=dy.device, dtype=torch.uint8)
        if fused_inproj() and planes is not None and dprojP is not None and out_wd.planes is not None:
            # the out-projection's input-gradient GEMM writes the dO operand planes and delta = rowsum(dO * O) itself: no fp32 d(ctx)
            datt = None
            _call('dx_outproj_dgrad_head_planes', _p(dprojP), _p(out_wd.planes), _p(att), _p(scratch), _p(lens), B, S, D, H, dh, _st())
        else:
            datt = conv_gemm(dproj, out_wd, None, B, S, x_planes=dprojP, lens=lens)   # dproj == 0 beyond len: exact
        d_out_w, d_out_b = conv_wgrad(att, dproj, B, S, D, D, 1, out_shape, x_planes=attP, dy_planes=dprojP, dbias=dpb, lens=lens,
                                      slot=ctx.slots[1])
        dqkv = torch.empty(B, S, 3 * D, device=dy.device, dtype=torch.float32)
        _call('dx_attention_bwd', _p(qkv), _p(planes), _p(lens), _p(att), _p(lse), _p(datt), _p(dqkv), _p(scratch), B, S, H, dh, p,
              seed_attn, _st())
        dqkvP, dqb = make_planes(dqkv, B * S, 3 * D, want_colsum=True)
        dx = conv_gemm(dqkv, in_wd, None, B, S, add_src=dv, x_planes=dqkvP, lens=lens)   # dqkv == dv == 0 beyond len: exact
        d_in_w, d_in_b = conv_wgrad(x, dqkv, B, S, D, 3 * D, 1, in_shape, x_planes=xP, dy_planes=dqkvP, dbias=dqb, lens=lens,
                                    slot=ctx.slots[0])
        return dx, None, d_in_w, d_in_b, d_out_w, d_out_b, dln_w, dln_b, None, None


class ConvFFSubLayer(torch.autograd.Function):
    """y = mask(gamma * LN(dropout(conv2(relu(conv1(x)))) + x) + beta);  film = (tensor [B, stride], column offset) or None"""

    @staticmethod
    def forward(ctx, x, lens, w1, b1, w2, b2, ln_w, ln_b, film, p_drop):
        x = _check_input(x)
        B, S, D = x.shape
        w1p, w1d = packed(w1)
        w2p, w2d = packed(w2)
        C = w1.shape[0]
        xP = planes_of(x, B * S, D)
        k2 = (w2.shape[2] - 1) // 2
        if plane_handover(D, C) and xP is not None:
            # the hidden activation (8x wider than the model) only ever feeds GEMMs: it exists as operand planes only,
            # written by conv1's epilogue; its ReLU mask in backward is read from the hi plane
            h, hP = conv_gemm(x, w1p, b1, B, S, relu=True, x_planes=xP, lens=lens, halo=k2, emit_planes=True, want_y=False)
        else:
            h = conv_gemm(x, w1p, b1, B, S, relu=True, round_out=True, x_planes=xP, lens=lens, halo=k2)
            hP = make_planes(h, B * S, C)
        seed = next_seed() if p_drop > 0 else 0
        if film is not None:
            film = _check_input(film)
            assert film.shape[1] == 2 * D   # reference model.py:232
        if can_fuse_ln(hP, w2p, b2, D):   # conv2 + dropout + residual + LayerNorm + FiLM + mask in ONE kernel
            y, xhat, rstd = gemm_ln(hP, w2p, b2, x, ln_w, ln_b, film, 2 * D, lens, B, S, p_in=p_drop, seed_in=seed)
        else:
            o = conv_gemm(h, w2p, b2, B, S, x_planes=hP, lens=lens)   # rows >= len are masked by the LayerNorm kernel
            y, xhat, rstd = ln_fwd(o, x, ln_w, ln_b, film, 2 * D, lens, B, S, D, p_in=p_drop, seed_in=seed, emit_planes=True)
        ctx.save_for_backward(x, lens, h, xhat, rstd, ln_w, ln_b, film, xP, hP)
        hold_packs(ctx, w1d, w2d)
        ctx.slots = (grad_slot(w1), grad_slot(w2))
        ctx.cfg = (B, S, D, C, float(p_drop), seed, w1.shape, w2.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, lens, h, xhat, rstd, ln_w, ln_b, film, xP, hP = ctx.saved_tensors
        w1d, w2d = held_packs(ctx)
        B, S, D, C, p, seed, w1_shape, w2_shape = ctx.cfg
        dy = _check_input(dy)
        dv, do, dln_w, dln_b, dfilm, doP, dob = ln_bwd(dy, xhat, rstd, ln_w, ln_b, film, 2 * D, lens, B, S, D, p_in=p, seed_in=seed,
                                                       want_film=film is not None, emit_planes=True)
        k2 = (w2_shape[2] - 1) // 2
        if h is None:   # plane hand-over: dh exists as planes only; its column sums (conv1 bias gradient) are taken from them
            dh, dhP, dhb = conv_gemm(do, w2d, None, B, S, relu_src_hi=hP, x_planes=doP, lens=lens, halo=k2, emit_planes=True, want_y=False,
                                     want_colsum=True)   # conv1's bias gradient = column sums of dh, taken by the epilogue
        else:
            dh = conv_gemm(do, w2d, None, B, S, relu_src=h, round_out=True, x_planes=doP, lens=lens, halo=k2)   # do == 0 beyond len: exact
            dhP, dhb = make_planes(dh, B * S, C, want_colsum=True)
        dw2, db2 = conv_wgrad(h, do, B, S, C, D, w2_shape[2], w2_shape, x_planes=hP, dy_planes=doP, dbias=dob, lens=lens, slot=ctx.slots[1])
        dx = conv_gemm(dh, w1d, None, B, S, add_src=dv, x_planes=dhP, lens=lens)   # rows >= len are masked by the producer's LN bwd
        dw1, db1 = conv_wgrad(x, dh, B, S, D, C, w1_shape[2], w1_shape, x_planes=xP, dy_planes=dhP, dbias=dhb, lens=lens, halo=k2,
                              slot=ctx.slots[0])
        return dx, None, dw1, db1, dw2, db2, dln_w, dln_b, dfilm, None


class PreNet(torch.autograd.Function):
    """3 x (Conv1d k3 -> ReLU -> LayerNorm -> Dropout) on the reference mel-spec [B, M, T] -> [B, T, D]"""

    @staticmethod
    def forward(ctx, mel, lens, w0, b0, g0, e0, w1, b1, g1, e1, w2, b2, g2, e2, p_drop):
        mel = _check_input(mel)
        B, M, T = mel.shape
        full = torch.full((B,), T, device=mel.device, dtype=torch.int64)
        x = torch.empty(B, T, M, device=mel.device, dtype=torch.float32)
        _call('dx_mask_transpose_bwd', _p(mel), _p(full), _p(x), B, T, M, _st())   # plain [B,M,T] -> [B,T,M]
        saved, cur = [x], x
        seeds, wds, in_planes = [], [], []
        layers = ((w0, b0, g0, e0), (w1, b1, g1, e1), (w2, b2, g2, e2))
        pads = [(w.shape[2] - 1) // 2 for w, _, _, _ in layers]
        # padding skip: rows >= len + (receptive field of the convs still to come) cannot reach a valid pre-net output row
        halos = [sum(pads[i + 1:]) for i in range(3)]
        for i, (w, b, g, e) in enumerate(layers):
            wp, wd = packed(w)
            curP = planes_of(cur, B * T, w.shape[1])
            in_planes.append(curP)
            a = conv_gemm(cur, wp, b, B, T, relu=True, x_planes=curP, lens=lens, halo=halos[i])
            seed = next_seed() if p_drop > 0 else 0
            y, xhat, rstd = ln_fwd(a, None, g, e, None, 0, None, B, T, w.shape[0], p_out=p_drop, seed_out=seed, emit_planes=i < 2)
            saved += [a, xhat, rstd, g, e]
            seeds.append(seed)
            wds.append(wd)
            cur = y
            saved.append(y)
        saved.append(lens)
        ctx.halos = halos
        ctx.save_for_backward(*saved)
        ctx.in_planes = in_planes   # bf16 planes of each layer's input (not autograd tensors of interest: plain buffers)
        hold_packs(ctx, *wds)
        ctx.slots = (grad_slot(w0), grad_slot(w1), grad_slot(w2))
        ctx.cfg = (B, T, M, float(p_drop), seeds, (w0.shape, w1.shape, w2.shape))
        return cur

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        B, T, M, p, seeds, shapes = ctx.cfg
        x, lens = saved[0], saved[-1]
        halos = ctx.halos
        layers = [saved[1 + 6 * i: 7 + 6 * i] for i in range(3)]   # a, xhat, rstd, g, e, y
        grads = [None] * 12
        d = _check_input(dy)
        for i in (2, 1, 0):
            a, xhat, rstd, g, e, _y = layers[i]
            wd = held_packs(ctx)[i]
            cout, cin, kw = shapes[i]
            inp = x if i == 0 else layers[i - 1][5]
            dpre, _, dg, de, _, dpreP, dpreb = ln_bwd(d, xhat, rstd, g, e, None, 0, None, B, T, cout, relu_src=a, p_out=p,
                                                      seed_out=seeds[i], emit_planes=True)
            # the gradient of the masked pre-net output is zero beyond len, so dpre of layer i is exactly zero beyond len + halos[i]
            dw, db = conv_wgrad(inp, dpre, B, T, cin, cout, kw, shapes[i], x_planes=ctx.in_planes[i], dy_planes=dpreP, dbias=dpreb,
                                lens=lens, halo=halos[i], slot=ctx.slots[i])
            grads[4 * i: 4 * i + 4] = [dw, db, dg, de]
            if i > 0:
                d = conv_gemm(dpre, wd, None, B, T, x_planes=dpreP, lens=lens, halo=halos[i - 1])
        return (None, None, *grads, None)


class FrameInput(torch.autograd.Function):
    """y = mask * (x + PE [+ conv3(energy) + conv3(pitch)])"""

    @staticmethod
    def forward(ctx, x, lens, pe, energy, pitch, we, be, wp, bp):
        x = _check_input(x)
        B, T, D = x.shape
        y = torch.empty_like(x)
        has = energy is not None
        if has:
            energy, pitch = _check_input(energy), _check_input(pitch)
        _call('dx_frame_input_fwd', _p(x), _p(energy), _p(pitch), _p(we), _p(be), _p(wp), _p(bp), _p(pe), _p(lens), _p(y),
              B, T, D, _st())
        ctx.save_for_backward(lens, energy, pitch)
        ctx.cfg = (B, T, D, has)
        return y

    @staticmethod
    def backward(ctx, dy):
        lens, energy, pitch = ctx.saved_tensors
        B, T, D, has = ctx.cfg
        dy = _check_input(dy)
        dx = torch.empty_like(dy)
        dwe = dbe = dwp = dbp = None
        if has:
            dwe = torch.empty(D, 1, 3, device=dy.device, dtype=torch.float32)
            dwp = torch.empty(D, 1, 3, device=dy.device, dtype=torch.float32)
            dbe = torch.empty(D, device=dy.device, dtype=torch.float32)
            dbp = torch.empty(D, device=dy.device, dtype=torch.float32)
        _call('dx_frame_input_bwd', _p(dy), _p(energy), _p(pitch), _p(lens), _p(dx), _p(dwe), _p(dbe), _p(dwp), _p(dbp),
              B, T, D, _st())
        return dx, None, None, None, None, dwe, dbe, dwp, dbp


class MeanPoolSpeaker(torch.autograd.Function):
    """pooled = sum_t x / len ;  h = pooled + spk_embedding[ids]   -> (pooled, h)"""

    @staticmethod
    def forward(ctx, x, lens, spk_ids, spk_emb):
        x = _check_input(x)
        B, S, D = x.shape
        pooled = torch.empty(B, D, device=x.device, dtype=torch.float32)
        h = torch.empty(B, D, device=x.device, dtype=torch.float32)
        _call('dx_meanpool_fwd', _p(x), _p(lens), _p(pooled), B, S, D, _st())
        _call('dx_add_speaker_fwd', _p(pooled), _p(spk_ids), _p(spk_emb), _p(h), B, D, spk_emb.shape[0], _st())
        ctx.save_for_backward(lens, spk_ids)
        ctx.cfg = (B, S, D, spk_emb.shape[0])
        return pooled, h

    @staticmethod
    def backward(ctx, dpooled, dh):
        lens, spk_ids = ctx.saved_tensors
        B, S, D, n_spk = ctx.cfg
        dh = _check_input(dh)
        demb = torch.empty(n_spk, D, device=dh.device, dtype=torch.float32)
        _call('dx_add_speaker_bwd', _p(dh), _p(spk_ids), _p(demb), B, D, n_spk, _st())
        tot = (dh + dpooled).contiguous()   # [B, D] host-side plumbing: both consumers of `pooled` feed one gradient
        dx = torch.empty(B, S, D, device=dh.device, dtype=torch.float32)
        _call('dx_meanpool_bwd', _p(tot), _p(lens), _p(dx), B, S, D, _st())
        return dx, None, None, demb


class FilmHead(torch.autograd.Function):
    """gammas/betas = Linear(h); split per module/block, scalar post-multipliers, gamma in the delta regime.
    Returns one contiguous [B, 2*ch] tensor per FiLM-ed block (module order: encoder blocks, predictor, decoder blocks)."""

    @staticmethod
    def forward(ctx, h, gw, gb, bw, bb, post, nb_blocks, channels):
        h = _check_input(h)
        B = h.shape[0]
        gwp, gwd = packed(gw)
        bwp, bwd = packed(bw)
        graw = linear_rows(h, gwp, gb)
        braw = linear_rows(h, bwp, bb)
        NF = graw.shape[1]
        film = torch.empty(B, 2 * NF, device=h.device, dtype=torch.float32)
        nb_arr = (ctypes.c_int * len(nb_blocks))(*nb_blocks)
        ch_arr = (ctypes.c_int * len(channels))(*channels)
        _call('dx_film_assemble_fwd', _p(graw), _p(braw), _p(post), _p(film), B, len(nb_blocks), ctypes.addressof(nb_arr),
              ctypes.addressof(ch_arr), _st())
        outs, col = [], 0
        for nb, ch in zip(nb_blocks, channels):
            for _ in range(nb):
                outs.append(film[:, col: col + 2 * ch].contiguous())
                col += 2 * ch
        ctx.save_for_backward(h, graw, braw, post)
        hold_packs(ctx, gwd, bwd)
        ctx.cfg = (B, NF, tuple(nb_blocks), tuple(channels), gw.shape, bw.shape)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        h, graw, braw, post = ctx.saved_tensors
        gwd, bwd = held_packs(ctx)
        B, NF, nb_blocks, channels, gshape, bshape = ctx.cfg
        parts, k = [], 0
        for nb, ch in zip(nb_blocks, channels):
            for _ in range(nb):
                d = douts[k]
                parts.append(d if d is not None else torch.zeros(B, 2 * ch, device=h.device, dtype=torch.float32))
                k += 1
        dfilm = torch.cat(parts, dim=1).contiguous()
        dgraw = torch.empty_like(graw)
        dbraw = torch.empty_like(braw)
        dpost = torch.empty_like(post) if post is not None else None
        nb_arr = (ctypes.c_int * len(nb_blocks))(*nb_blocks)
        ch_arr = (ctypes.c_int * len(channels))(*channels)
        _call('dx_film_assemble_bwd', _p(dfilm), _p(graw), _p(braw), _p(post), _p(dgraw), _p(dbraw), _p(dpost), B,
              len(nb_blocks), ctypes.addressof(nb_arr), ctypes.addressof(ch_arr), _st())
        D = h.shape[1]
        dh1 = linear_rows(dgraw, gwd, None)
        dh = linear_rows(dbraw, bwd, None, add_src=dh1)
        dgw, dgb = conv_wgrad(h, dgraw, 1, B, D, NF, 1, gshape)
        dbw, dbb = conv_wgrad(h, dbraw, 1, B, D, NF, 1, bshape)
        return dh, dgw, dgb, dbw, dbb, dpost, None, None


class Linear(torch.autograd.Function):
    """y = relu?(x W^T + b); the input gradient is multiplied by `grad_in_scale` (gradient reversal: -lambda)."""

    @staticmethod
    def forward(ctx, x, w, b, relu, grad_in_scale):
        x = _check_input(x)
        wp, wd = packed(w)
        y = linear_rows(x, wp, b, relu=relu)
        ctx.save_for_backward(x, y if relu else None)
        hold_packs(ctx, wd)
        ctx.cfg = (relu, float(grad_in_scale), w.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        wd = held_packs(ctx)
        relu, scale, wshape = ctx.cfg
        dy = _check_input(dy)
        R, cout, cin = x.shape[0], wshape[0], wshape[1]
        if relu:
            dpre = torch.empty_like(dy)
            _call('dx_relu_bwd', _p(dy), _p(y), _p(dpre), dy.numel(), _st())
            dy = dpre
        dx = linear_rows(dy, wd, None, alpha=scale)
        dw, db = conv_wgrad(x, dy, 1, R, cin, cout, 1, wshape)
        return dx, dw, db, None, None


class EmbedPE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, symbols, lens, emb, pe):
        B, L = symbols.shape
        D = emb.shape[1]
        y = torch.empty(B, L, D, device=emb.device, dtype=torch.float32)
        symbols = _check_input(symbols, torch.int64)
        _call('dx_embed_pe_fwd', _p(symbols), _p(lens), _p(emb), _p(pe), _p(y), B, L, D, emb.shape[0], _st())
        ctx.save_for_backward(symbols, lens)
        ctx.cfg = (B, L, D, emb.shape[0])
        return y

    @staticmethod
    def backward(ctx, dy):
        symbols, lens = ctx.saved_tensors
        B, L, D, n = ctx.cfg
        dy = _check_input(dy)
        demb = torch.empty(n, D, device=dy.device, dtype=torch.float32)
        _call('dx_embed_pe_bwd', _p(symbols), _p(lens), _p(dy), _p(demb), B, L, D, n, _st())
        return None, None, demb, None


class PredictorBlock(torch.autograd.Function):
    """One block of LocalProsodyPredictor (model.py:526-543,559-566): conv(k3)+ReLU -> LN -> drop -> conv(k3)+ReLU -> LN -> drop -> FiLM,
    followed by the padding mask when `lens` is given (the reference masks once, after the LAST block, model.py:567-568)."""

    @staticmethod
    def forward(ctx, x, lens, film, w0, b0, g0, e0, w1, b1, g1, e1, p_drop):
        x = _check_input(x)
        B, L, D = x.shape
        C = w0.shape[0]
        w0p, w0d = packed(w0)
        w1p, w1d = packed(w1)
        film = _check_input(film)
        assert film.shape[1] == 2 * C   # reference model.py:561
        s0, s1 = (next_seed(), next_seed()) if p_drop > 0 else (0, 0)
        a0 = conv_gemm(x, w0p, b0, B, L, relu=True)
        y0, xh0, rs0 = ln_fwd(a0, None, g0, e0, None, 0, None, B, L, C, p_out=p_drop, seed_out=s0)
        a1 = conv_gemm(y0, w1p, b1, B, L, relu=True)
        y1, xh1, rs1 = ln_fwd(a1, None, g1, e1, film, 2 * C, lens, B, L, C, p_out=p_drop, seed_out=s1)
        ctx.save_for_backward(x, lens, film, a0, y0, xh0, rs0, a1, xh1, rs1, g0, e0, g1, e1)
        hold_packs(ctx, w0d, w1d)
        ctx.cfg = (B, L, D, C, float(p_drop), s0, s1, w0.shape, w1.shape)
        return y1

    @staticmethod
    def backward(ctx, dy1):
        x, lens, film, a0, y0, xh0, rs0, a1, xh1, rs1, g0, e0, g1, e1 = ctx.saved_tensors
        w0d, w1d = held_packs(ctx)
        B, L, D, C, p, s0, s1, w0s, w1s = ctx.cfg
        dy1 = _check_input(dy1)
        dpre1, _, dg1, de1, dfilm = ln_bwd(dy1, xh1, rs1, g1, e1, film, 2 * C, lens, B, L, C, relu_src=a1, p_out=p,
                                           seed_out=s1, want_film=True)
        dy0 = conv_gemm(dpre1, w1d, None, B, L)
        dw1, db1 = conv_wgrad(y0, dpre1, B, L, C, C, w1s[2], w1s)
        dpre0, _, dg0, de0, _ = ln_bwd(dy0, xh0, rs0, g0, e0, None, 0, None, B, L, C, relu_src=a0, p_out=p, seed_out=s0)
        dx = conv_gemm(dpre0, w0d, None, B, L)
        dw0, db0 = conv_wgrad(x, dpre0, B, L, D, C, w0s[2], w0s)
        return dx, None, dfilm, dw0, db0, dg0, de0, dw1, db1, dg1, de1, None


class PredictorHead(torch.autograd.Function):
    """Linear(C -> 3) + mask (model.py:569-575) on the masked block output.  Returns preds [3, B, L] (duration, energy, pitch planes)."""

    @staticmethod
    def forward(ctx, y, lens, pw, pb):
        y = _check_input(y)
        B, L, C = y.shape
        NO = pw.shape[0]
        out = torch.empty(NO, B, L, device=y.device, dtype=torch.float32)
        _call('dx_narrow_linear_fwd', _p(y), _p(pw), _p(pb), _p(lens), _p(out), B, L, C, NO, _st())
        ctx.save_for_backward(y, lens, pw)
        ctx.cfg = (B, L, C, NO)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, lens, pw = ctx.saved_tensors
        B, L, C, NO = ctx.cfg
        dout = _check_input(dout)
        dy = torch.empty(B, L, C, device=dout.device, dtype=torch.float32)
        dpw = torch.empty(NO, C, device=dout.device, dtype=torch.float32)
        dpb = torch.empty(NO, device=dout.device, dtype=torch.float32)
        _call('dx_narrow_linear_bwd', _p(dout), _p(y), _p(pw), _p(lens), _p(dy), _p(dpw), _p(dpb), B, L, C, NO, _st())
        return dy, None, dpw, dpb


class GaussUpsample(torch.autograd.Function):
    """-> (x_upsamp [B,T,D], weights [B,L,T], csum [B,L] int64, totals [B] int64)"""

    @staticmethod
    def forward(ctx, x, dur_f, dur_i, energy, pitch, lens, wd, bd, we, be, wp, bp, rw, rb, T):
        ctx.set_materialize_grads(False)
        x = _check_input(x)
        B, L, D = x.shape
        dev = x.device
        dur_f, energy, pitch = _check_input(dur_f), _check_input(energy), _check_input(pitch)
        dur_i = _check_input(dur_i, torch.int64)
        xp = torch.empty(B, L, D, device=dev, dtype=torch.float32)
        z = torch.empty(B, L, device=dev, dtype=torch.float32)
        sigma = torch.empty(B, L, device=dev, dtype=torch.float32)
        mu = torch.empty(B, L, device=dev, dtype=torch.float32)
        csum = torch.empty(B, L, device=dev, dtype=torch.int64)
        total = torch.empty(B, device=dev, dtype=torch.int64)
        _call('dx_gauss_prep', _p(x), _p(dur_f), _p(dur_i), _p(energy), _p(pitch), _p(lens), _p(wd), _p(bd), _p(we), _p(be),
              _p(wp), _p(bp), _p(rw), _p(rb), _p(xp), _p(z), _p(sigma), _p(mu), _p(csum), _p(total), B, L, D, _st())
        if T is None:   # inference: T_max = max(sum(durations_int)) needs one read-back (the reference syncs here too)
            T = int(total.max().item())
        up = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        weights = torch.empty(B, L, T, device=dev, dtype=torch.float32)
        _call('dx_gauss_upsample_fwd', _p(xp), _p(mu), _p(sigma), _p(lens), _p(up), _p(weights), B, L, T, D, _st())
        ctx.save_for_backward(up, weights, xp, z, mu, sigma, dur_f, energy, pitch, lens, wd, bd, rw)
        ctx.cfg = (B, L, T, D)
        ctx.mark_non_differentiable(csum, total)
        return up, weights, csum, total

    @staticmethod
    def backward(ctx, dup, dweights, _dc, _dt):
        up, weights, xp, z, mu, sigma, dur_f, energy, pitch, lens, wd, bd, rw = ctx.saved_tensors
        B, L, T, D = ctx.cfg
        dev = up.device
        dup = _check_input(dup) if dup is not None else torch.zeros_like(up)
        dweights = _check_input(dweights) if dweights is not None else None
        dx = torch.empty(B, L, D, device=dev, dtype=torch.float32)
        g = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        dwd, dwe, dwp = g(D, 1, 3), g(D, 1, 3), g(D, 1, 3)
        dbd, dbe, dbp = g(D), g(D), g(D)
        drw, drb = g(1, D), g(1)
        scratch = g(B * L + 2 * B * T)
        _call('dx_gauss_upsample_bwd', _p(dup), _p(dweights), _p(up), _p(weights), _p(xp), _p(z), _p(mu), _p(sigma),
              _p(dur_f), _p(energy), _p(pitch), _p(lens), _p(wd), _p(bd), _p(rw), _p(dx), _p(dwd), _p(dbd), _p(dwe), _p(dbe),
              _p(dwp), _p(dbp), _p(drw), _p(drb), _p(scratch), B, L, T, D, _st())
        return dx, None, None, None, None, None, dwd, dbd, dwe, dbe, dwp, dbp, drw, drb, None


class MelProjection(torch.autograd.Function):
    """mel [B, M, T] = transpose(mask * (x W^T + b))"""

    @staticmethod
    def forward(ctx, x, lens, w, b):
        x = _check_input(x)
        B, T, D = x.shape
        M = w.shape[0]
        wp, wd = packed(w)
        y = conv_gemm(x, wp, b, B, T, lens=lens)
        mel = torch.empty(B, M, T, device=x.device, dtype=torch.float32)
        _call('dx_mask_transpose_fwd', _p(y), _p(lens), _p(mel), B, T, M, _st())
        ctx.save_for_backward(x, lens)
        hold_packs(ctx, wd)
        ctx.cfg = (B, T, D, M, w.shape)
        return mel

    @staticmethod
    def backward(ctx, dmel):
        x, lens = ctx.saved_tensors
        wd = held_packs(ctx)
        B, T, D, M, wshape = ctx.cfg
        dmel = _check_input(dmel)
        dy = torch.empty(B, T, M, device=dmel.device, dtype=torch.float32)
        _call('dx_mask_transpose_bwd', _p(dmel), _p(lens), _p(dy), B, T, M, _st())
        dx = conv_gemm(dy, wd, None, B, T, lens=lens)   # dy == 0 beyond len: exact
        dw, db = conv_wgrad(x, dy, B, T, D, M, 1, wshape, lens=lens)
        return dx, None, dw, db


class Loss(torch.autograd.Function):
    """-> out[8] = {speaker, post_mult, duration, energy, pitch, mel_l1, mel_l2, total}; only out[7] is differentiable."""

    @staticmethod
    def forward(ctx, spk_logits, post, dur_p, energy_p, pitch_p, mel_p, spk_ids, dur_t, energy_t, pitch_t, mel_t, in_lens,
                out_lens, weights):
        spk_logits, dur_p, energy_p, pitch_p, mel_p = (_check_input(t) for t in (spk_logits, dur_p, energy_p, pitch_p, mel_p))
        dur_t, energy_t, pitch_t, mel_t = (_check_input(t) for t in (dur_t, energy_t, pitch_t, mel_t))
        post_c = _check_input(post) if post is not None else None
        B, NS = spk_logits.shape
        L = dur_p.shape[1]
        M, T = mel_p.shape[1], mel_p.shape[2]
        NP = post_c.numel() if post_c is not None else 0
        acc = torch.empty(B, 8, device=mel_p.device, dtype=torch.float32)
        out = torch.empty(8, device=mel_p.device, dtype=torch.float32)
        w = tuple(float(v) for v in weights)
        _call('dx_loss_fwd', _p(spk_logits), _p(spk_ids), _p(post_c), _p(dur_p), _p(energy_p), _p(pitch_p), _p(dur_t),
              _p(energy_t), _p(pitch_t), _p(mel_p), _p(mel_t), _p(in_lens), _p(out_lens), B, L, T, M, NS, NP, *w, _p(acc),
              _p(out), _st())
        ctx.save_for_backward(spk_logits, spk_ids, post_c, dur_p, energy_p, pitch_p, dur_t, energy_t, pitch_t, mel_p, mel_t,
                              in_lens, out_lens)
        ctx.cfg = (B, L, T, M, NS, NP, w, post.shape if post is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        (spk_logits, spk_ids, post, dur_p, energy_p, pitch_p, dur_t, energy_t, pitch_t, mel_p, mel_t, in_lens,
         out_lens) = ctx.saved_tensors
        B, L, T, M, NS, NP, w, post_shape = ctx.cfg
        dev = mel_p.device
        gout = _check_input(dout)[7:8].contiguous()   # d(total); the 7 individual terms are reporting-only (loss.py:102-104)
        g = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        dlog, ddur, den, dpi, dmel = g(B, NS), g(B, L), g(B, L), g(B, L), g(B, M, T)
        dpost = g(*post_shape) if post is not None else None
        _call('dx_loss_bwd', _p(gout), _p(spk_logits), _p(spk_ids), _p(post), _p(dur_p), _p(energy_p), _p(pitch_p), _p(dur_t),
              _p(energy_t), _p(pitch_t), _p(mel_p), _p(mel_t), _p(in_lens), _p(out_lens), B, L, T, M, NS, NP, *w, _p(dlog),
              _p(dpost), _p(ddur), _p(den), _p(dpi), _p(dmel), _st())
        return dlog, dpost, ddur, den, dpi, dmel, None, None, None, None, None, None, None, None
