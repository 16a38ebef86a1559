"""Score network through the single native entry point ``sd_scorenet_forward`` (csrc/scorenet_forward.cu).

``NativeScoreNet(bound)`` packs the GEMM-layout weights of a bound ``ScoreNet`` (models/ddpm.py) into the one device blob
the C entry expects and exposes the reference's ``model_fn(t, x, y)`` call shape (cifar/models/utils.py:86-96).  The
blob and the descriptor fields are everything a non-Python host needs: ``save(path)`` writes them to disk
(``<path>.bin`` = the blob, ``<path>.json`` = the configuration), see INTEGRATION.md for the C side.
"""
import ctypes
import json

import torch

from . import _lib

ALIGN = 256
PRECISION_BF16, PRECISION_FP32_FAITHFUL = 1, 2


def _arrays(bound):
    """The blob's arrays in the order csrc/scorenet_forward.cu::layout walks."""
    b = bound
    out = [b.temb_w0, b.temb_b0, b.temb_w1, b.temb_b1]
    if b.class_emb is not None:
        out.append(b.class_emb)
    if not b.conv_in_tc:
        raise NotImplementedError("sd_scorenet_forward needs the tensor-core first conv (<= 3 channels, image size % 16 == 0)")
    out += [b.conv_in_w64, b.conv_in_b, b.dense_w, b.dense_b]
    for r in b.res:
        out += [r["g1"], r["be1"], r["w1"], r["g2"], r["be2"], r["w2"], r["b2"]]
    for a in b.attn:
        out += [a["g"], a["be"], a["w_q2"], a["b_q2"], a["w_voT"], a["b_vo"]]
    for d in b.down:
        out += [d["w"], d["b"]]
    for u in b.up:
        out += [u["w4"], u["b"]]
    out += [b.out_g, b.out_be, b.out_w, b.out_b]
    return out


def pack_weights(bound):
    """One uint8 device tensor: every array of `_arrays` on a 256-byte boundary."""
    arrs = _arrays(bound)
    total = 0
    offs = []
    for t in arrs:
        offs.append(total)
        total = (total + t.numel() * t.element_size() + ALIGN - 1) // ALIGN * ALIGN
    blob = torch.zeros(total, dtype=torch.uint8, device=bound.device)
    for t, o in zip(arrs, offs):
        n = t.numel() * t.element_size()
        blob[o:o + n] = t.contiguous().view(-1).view(torch.uint8)
    return blob


def make_desc(config, blob=None, precision=PRECISION_BF16):
    m, d = config.model, config.data
    desc = _lib.ScoreNetDesc()
    desc.precision = int(precision)
    desc.image_size, desc.channels, desc.nf = int(d.image_size), int(d.num_channels), int(m.nf)
    desc.num_res_blocks = int(m.num_res_blocks)
    ch_mult, attn = tuple(m.ch_mult), tuple(m.attn_resolutions)
    desc.n_levels = len(ch_mult)
    for i, v in enumerate(ch_mult):
        desc.ch_mult[i] = int(v)
    desc.n_attn_res = len(attn)
    for i, v in enumerate(attn):
        desc.attn_resolutions[i] = int(v)
    desc.conditioned = int(bool(m.conditioned))
    desc.num_classes = int(getattr(d, "num_classes", 0) or 0)
    if blob is not None:
        desc.weights = blob.data_ptr()
        desc.weights_bytes = blob.numel()
    return desc


class NativeScoreNet:
    """model_fn(t, x, y) over sd_scorenet_forward: one C call per forward, activations in a caller-owned workspace."""

    def __init__(self, bound):
        self.config = bound.config
        self.device = bound.device
        self.blob = pack_weights(bound)
        self.precision = PRECISION_FP32_FAITHFUL if getattr(bound, "split", False) else PRECISION_BF16
        self.desc = make_desc(self.config, self.blob, self.precision)
        lib = _lib.load()
        need = ctypes.c_size_t()
        _lib.check(lib.sd_scorenet_weights_bytes(ctypes.byref(self.desc), ctypes.byref(need)), "sd_scorenet_weights_bytes")
        if need.value != self.blob.numel():
            raise RuntimeError(f"weight blob is {self.blob.numel()} bytes, the native layout expects {need.value}")
        self._ws = None

    def workspace_bytes(self, B, t_stride=0):
        need = ctypes.c_size_t()
        _lib.check(_lib.load().sd_scorenet_workspace_bytes(ctypes.byref(self.desc), int(B), int(t_stride), ctypes.byref(need)),
                   "sd_scorenet_workspace_bytes")
        return need.value

    def __call__(self, t, x, y=None, out=None):
        _lib.require_device()
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise ValueError("x must be a contiguous float32 CUDA tensor (NHWC)")
        B = x.shape[0]
        if not torch.is_tensor(t):
            t = torch.full((1,), float(t), device=x.device, dtype=torch.float32)
        t = t.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
        stride = 0 if t.numel() == 1 else 1
        if stride and t.numel() != B:
            raise ValueError("t must have one entry per sample")
        labels = None
        if self.desc.conditioned:
            if y is None:
                raise ValueError("conditioned score-net needs labels")
            labels = y.to(device=x.device, dtype=torch.int32).contiguous()
        need = self.workspace_bytes(B, stride)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        if out is None:
            out = torch.empty_like(x)
        rc = _lib.load().sd_scorenet_forward(ctypes.byref(self.desc), t.data_ptr(), stride, x.data_ptr(),
                                             labels.data_ptr() if labels is not None else None, B, out.data_ptr(),
                                             self._ws.data_ptr(), self._ws.numel(), self.precision,
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "sd_scorenet_forward")
        return out

    def forward_sched(self, sched, step_counter, x, y=None, out=None):
        """model_fn at t = sched[*step_counter].sigma, read on the device (sd_scorenet_forward_sched): every argument is
        fixed across timesteps, so the call can sit in a CUDA graph that replays for the whole loop."""
        _lib.require_device()
        B = x.shape[0]
        labels = None
        if self.desc.conditioned:
            if y is None:
                raise ValueError("conditioned score-net needs labels")
            labels = y.to(device=x.device, dtype=torch.int32).contiguous()
        need = self.workspace_bytes(B, 0)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        if out is None:
            out = torch.empty_like(x)
        rc = _lib.load().sd_scorenet_forward_sched(ctypes.byref(self.desc), sched.data_ptr(), step_counter.data_ptr(), x.data_ptr(),
                                                   labels.data_ptr() if labels is not None else None, B, out.data_ptr(),
                                                   self._ws.data_ptr(), self._ws.numel(), self.precision,
                                                   torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "sd_scorenet_forward_sched")
        return out

    def save(self, path):
        """<path>.bin: the weight blob; <path>.json: the sd_scorenet_desc fields."""
        self.blob.cpu().numpy().tofile(path + ".bin")
        d = self.desc
        with open(path + ".json", "w") as fh:
            json.dump({"image_size": d.image_size, "channels": d.channels, "nf": d.nf, "num_res_blocks": d.num_res_blocks,
                       "ch_mult": list(d.ch_mult)[:d.n_levels], "attn_resolutions": list(d.attn_resolutions)[:d.n_attn_res],
                       "conditioned": d.conditioned, "num_classes": d.num_classes, "weights_bytes": int(d.weights_bytes),
                       "precision": int(d.precision)}, fh)
