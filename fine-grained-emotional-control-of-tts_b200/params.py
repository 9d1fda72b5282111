"""Parameter storage for the drop-in FastSpeech2: one flat fp32 buffer (so AdamW and the NCCL gradient
all-reduce are single passes) exposed as nn.Parameters under exactly the reference's state_dict keys
(`/root/reference/emo_rank_tts/fastspeech2/model.py:187-276`), plus the packed
(tap-major, operand-dtype) weight copies that the implicit-GEMM kernels read.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L


class PackedWeight:
    """A GEMM weight in the packed operand buffer: rows = cout, row = [k taps][cin]."""

    __slots__ = ("name", "cout", "cin", "k", "off", "src_ld", "src_col0")

    def __init__(self, name, cout, cin, k, src_ld=None, src_col0=0):
        self.name, self.cout, self.cin, self.k = name, cout, cin, k
        self.off = 0
        self.src_ld = src_ld if src_ld is not None else cin * k
        self.src_col0 = src_col0


def _container(root: nn.Module, path: str) -> nn.Module:
    mod = root
    for part in path.split("."):
        if part not in mod._modules:
            mod.add_module(part, nn.Module())
        mod = mod._modules[part]
    return mod


class ParamStore:
    """Builds the parameter tree of `root` inside one flat buffer."""

    def __init__(self, root: nn.Module):
        self.root = root
        self.specs = []          # (key, shape, init)
        self.packed = {}         # key -> PackedWeight
        self.offsets = {}        # key -> (offset, numel)
        self.total = 0
        self.flat = None
        self.flat_grad = None
        self.params = {}         # key -> nn.Parameter
        self.packed_buf = None
        self.packed_total = 0
        self._items_dev = None
        self._packed_dtype = None

    # ------------------------------------------------------------------ declaration
    def add(self, key, shape, init):
        self.specs.append((key, tuple(shape), init))

    def add_packed(self, key, cout, cin, k, src_ld=None, src_col0=0, name=None):
        self.packed[name or key] = (key, PackedWeight(name or key, cout, cin, k, src_ld, src_col0))

    def linear(self, prefix, cout, cin, bias=True):
        self.add(prefix + ".weight", (cout, cin), ("kaiming_uniform", cin))
        if bias:
            self.add(prefix + ".bias", (cout,), ("uniform_fan", cin))
        self.add_packed(prefix + ".weight", cout, cin, 1)

    def conv(self, prefix, cout, cin, k, pack=True):
        self.add(prefix + ".weight", (cout, cin, k), ("kaiming_uniform", cin * k))
        self.add(prefix + ".bias", (cout,), ("uniform_fan", cin * k))
        if pack:
            self.add_packed(prefix + ".weight", cout, cin, k)

    def layernorm(self, prefix, c):
        self.add(prefix + ".weight", (c,), ("ones",))
        self.add(prefix + ".bias", (c,), ("zeros",))

    # ---------------------------------------------------------------------- build
    def build(self, device="cpu"):
        off = 0
        for key, shape, _ in self.specs:
            n = int(math.prod(shape))
            self.offsets[key] = (off, n)
            off += (n + 3) // 4 * 4          # 16-byte aligned parameter starts (float4 kernels)
        self.total = off
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        for key, shape, init in self.specs:
            o, n = self.offsets[key]
            view = self.flat[o:o + n].view(shape)
            self._init(view, init)
            p = nn.Parameter(view)
            path, _, leaf = key.rpartition(".")
            _container(self.root, path).register_parameter(leaf, p)
            self.params[key] = p
        poff = 0
        for name, (key, pw) in self.packed.items():
            pw.off = poff
            poff += (pw.cout * pw.cin * pw.k + 63) // 64 * 64   # 128-byte aligned rows for TMA
        self.packed_total = poff

    @staticmethod
    def _init(view, init):
        kind = init[0]
        with torch.no_grad():
            if kind == "ones":
                view.fill_(1.0)
            elif kind == "zeros":
                view.zero_()
            elif kind == "normal":
                view.normal_(0.0, 1.0)
            elif kind == "kaiming_uniform":      # torch Linear/Conv default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                bound = 1.0 / math.sqrt(init[1])
                view.uniform_(-bound, bound)
            elif kind == "uniform_fan":
                bound = 1.0 / math.sqrt(init[1])
                view.uniform_(-bound, bound)
            elif kind == "xavier_uniform":       # nn.MultiheadAttention in_proj_weight
                fan_out, fan_in = view.shape
                bound = math.sqrt(6.0 / (fan_in + fan_out))
                view.uniform_(-bound, bound)
            else:
                raise ValueError(kind)

    # ------------------------------------------------------------- device movement
    def reflatten(self):
        """Re-establish the flat-buffer views after nn.Module._apply replaced parameter storage."""
        first = next(iter(self.params.values()))
        dev, dt = first.device, first.dtype
        if dt != torch.float32:
            raise RuntimeError("fs2_b200: master parameters must stay fp32 (precision is chosen with precision=...)")
        new_flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        for key, p in self.params.items():
            o, n = self.offsets[key]
            new_flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = new_flat[o:o + n].view(p.shape)
            p.grad = None
        self.flat = new_flat
        self.flat_grad = None
        self.packed_buf = None
        self._items_dev = None

    def views_intact(self):
        base = self.flat.data_ptr()
        for key, p in self.params.items():
            o, _ = self.offsets[key]
            if p.data_ptr() != base + 4 * o:
                return False
        return True

    # ------------------------------------------------------------------ gradients
    def ensure_grads(self):
        """Make every p.grad a view of one flat fp32 buffer; returns True when the buffer was (re)zeroed.
        `optimizer.zero_grad()` (set_to_none) is honoured: missing grads mean "start from zero"."""
        if self.flat_grad is None or self.flat_grad.device != self.flat.device:
            self.flat_grad = torch.zeros_like(self.flat)
            fresh = True
        else:
            fresh = False
        need_zero = fresh
        base = self.flat_grad.data_ptr()
        for key, p in self.params.items():
            o, n = self.offsets[key]
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * o:
                if g is not None and not fresh:
                    raise RuntimeError("fs2_b200: parameter .grad tensors were replaced by foreign storage")
                need_zero = True
        if need_zero:
            if not fresh:
                L.call("fs2_memset", self.flat_grad, 0, self.flat_grad.numel() * 4)
            for key, p in self.params.items():
                o, n = self.offsets[key]
                p.grad = self.flat_grad[o:o + n].view(p.shape)
        return need_zero

    def grad(self, key):
        o, n = self.offsets[key]
        return self.flat_grad[o:o + n]

    # -------------------------------------------------------------------- packing
    def pack(self, bf16: bool):
        """Refresh the packed operand copies from the fp32 masters (one kernel launch)."""
        dt = torch.bfloat16 if bf16 else torch.float32
        if self.packed_buf is None or self._packed_dtype != dt or self.packed_buf.device != self.flat.device:
            self.packed_buf = torch.zeros(self.packed_total, dtype=dt, device=self.flat.device)
            self._packed_dtype = dt
            items = (L.Fs2PackItem * len(self.packed))()
            for i, (name, (key, pw)) in enumerate(self.packed.items()):
                o, _ = self.offsets[key]
                items[i].src_off = o + pw.src_col0
                items[i].dst_off = pw.off
                items[i].src_ld = pw.src_ld
                items[i].cout, items[i].cin, items[i].k = pw.cout, pw.cin, pw.k
            raw = bytes(items)
            host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            self._items_dev = host.to(self.flat.device)
        L.call("fs2_pack_weights", self._items_dev, len(self.packed), self.flat, self.packed_buf, int(bf16))

    def pw(self, name) -> PackedWeight:
        return self.packed[name][1]
