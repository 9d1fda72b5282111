"""Parameter storage for the drop-in FastSpeech2: one flat fp32 buffer (so AdamW and the NCCL gradient
all-reduce are single passes) exposed as nn.Parameters under exactly the reference's state_dict keys
(`/root/reference/emo_rank_tts/fastspeech2/model.py:187-276`).

Layout of a Conv1d weight with k > 1: the reference's parameter is (Cout, Cin, k).  The implicit-GEMM kernels read a
weight tap-major, [Cout][k][Cin] (one K-slab of Cin per tap).  The fp32 MASTER is therefore stored tap-major inside the
flat buffer and the nn.Parameter is the permuted view of it -- same key, same shape, same values, strides (k*Cin, 1, Cin).
state_dict()/load_state_dict()/torch.optim see an ordinary (Cout, Cin, k) tensor; the kernels see their operand layout
without a per-step re-packing pass:
  * fp32 path: the GEMMs read the master directly;
  * bf16 path: a bf16 MIRROR of the whole flat buffer (same element offsets) is the operand; FusedAdamW writes it in the
    same pass that updates the master, any other change of the parameters is followed by one cast pass
    (`sync_operands`).
The only operand that is not a whole parameter -- the token columns [0:384) of `concat_proj.w.weight` (384 x 773, row
pitch not TMA-aligned) -- is gathered into a small tail region of the mirror (bf16) / a side buffer (fp32).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L


ALIGN = 64      # floats: parameter starts are 256-byte aligned, so the same offset in the bf16 mirror is 128-byte aligned (TMA)


class PackedWeight:
    """A GEMM weight operand: rows = cout, row = [k taps][cin].  `gather`: a column block of a wider matrix (copied)."""

    __slots__ = ("name", "key", "cout", "cin", "k", "off", "src_ld", "src_col0", "gather", "goff")

    def __init__(self, name, key, cout, cin, k, src_ld=None, src_col0=0):
        self.name, self.key, self.cout, self.cin, self.k = name, key, cout, cin, k
        self.off = 0                 # element offset of the parameter in the flat buffer / mirror
        self.src_ld = src_ld if src_ld is not None else cin * k
        self.src_col0 = src_col0
        self.gather = src_ld is not None and (src_ld != cin * k or src_col0 != 0)
        self.goff = 0                # offset inside the gather region


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
        self.mirror = None           # bf16 copy of flat (+ gather tail): the tensor-core operand buffer
        self.gather32 = None         # fp32 gather region (precision="fp32")
        self.gather_total = 0
        self.mirror_valid = False    # set by FusedAdamW.step (it wrote the mirror), consumed by sync_operands
        self._mirror_token = -1
        self.tapmajor = set()        # keys stored [Cout][k][Cin]
        self._items_dev = None

    # ------------------------------------------------------------------ declaration
    def add(self, key, shape, init):
        self.specs.append((key, tuple(shape), init))

    def add_packed(self, key, cout, cin, k, src_ld=None, src_col0=0, name=None):
        self.packed[name or key] = (key, PackedWeight(name or key, key, cout, cin, k, src_ld, src_col0))

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
            if k > 1 and cin > 1:
                self.tapmajor.add(prefix + ".weight")

    def layernorm(self, prefix, c):
        self.add(prefix + ".weight", (c,), ("ones",))
        self.add(prefix + ".bias", (c,), ("zeros",))

    # ---------------------------------------------------------------------- build
    def build(self, device="cpu"):
        off = 0
        for key, shape, _ in self.specs:
            n = int(math.prod(shape))
            self.offsets[key] = (off, n)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.shapes = {key: shape for key, shape, _ in self.specs}
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        for key, shape, init in self.specs:
            view = self.view_of(self.flat, key)
            if key in self.tapmajor:
                tmp = torch.empty(shape)         # same RNG stream and values as a contiguous (Cout, Cin, k) parameter
                self._init(tmp, init)
                with torch.no_grad():
                    view.copy_(tmp)
            else:
                self._init(view, init)
            p = nn.Parameter(view)
            path, _, leaf = key.rpartition(".")
            _container(self.root, path).register_parameter(leaf, p)
            self.params[key] = p
        goff = 0
        for name, (key, pw) in self.packed.items():
            pw.off = self.offsets[key][0]
            if pw.gather:
                pw.goff = goff
                goff += (pw.cout * pw.cin * pw.k + ALIGN - 1) // ALIGN * ALIGN
        self.gather_total = goff

    def view_of(self, flat, key):
        """The parameter-shaped view of `flat` (the master, the gradient buffer, ...) for `key`."""
        o, n = self.offsets[key]
        shape = self.shapes[key]
        if key in self.tapmajor:
            cout, cin, k = shape
            return flat[o:o + n].view(cout, k, cin).permute(0, 2, 1)
        return flat[o:o + n].view(shape)

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
            v = self.view_of(new_flat, key)
            v.copy_(p.data)
            p.data = v
            p.grad = None
        self.flat = new_flat
        self.flat_grad = None
        self.mirror = None
        self.gather32 = None
        self.mirror_valid = False
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
                p.grad = self.view_of(self.flat_grad, key)
        return need_zero

    def grad(self, key):
        o, n = self.offsets[key]
        return self.flat_grad[o:o + n]

    # ------------------------------------------------------------------- operands
    def _gather_items(self):
        gl = [pw for _, (_, pw) in self.packed.items() if pw.gather]
        if self._items_dev is None and gl:
            items = (L.Fs2PackItem * len(gl))()
            for i, pw in enumerate(gl):
                items[i].src_off = pw.off + pw.src_col0
                items[i].dst_off = pw.goff
                items[i].src_ld = pw.src_ld
                items[i].cout, items[i].cin, items[i].k = pw.cout, pw.cin, pw.k
            host = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8)
            self._items_dev = host.to(self.flat.device)
        return gl

    def sync_operands(self, bf16: bool):
        """Make the GEMM operand buffers current.  bf16: the mirror is rewritten by one cast pass unless FusedAdamW.step
        has just produced it together with the parameter update (`mirror_valid`, consumed here: any parameter change
        this class cannot see -- torch.optim, load_state_dict, .data edits -- is therefore followed by a fresh cast).
        fp32: the masters are the operands; only the gathered column block is copied."""
        dev = self.flat.device
        gl = self._gather_items()
        if bf16:
            if self.mirror is None or self.mirror.device != dev:
                self.mirror = torch.zeros(self.total + self.gather_total, dtype=torch.bfloat16, device=dev)
                self.mirror_valid = False
            if self.mirror_valid and self._mirror_token == self.version_token():
                self.mirror_valid = False
                return
            self.mirror_valid = False
            L.call("fs2_cast_bf16", self.flat, self.mirror, self.total)
            if gl:
                L.call("fs2_pack_weights", self._items_dev, len(gl), self.flat, self.mirror[self.total:], 1)
        elif gl:
            if self.gather32 is None or self.gather32.device != dev:
                self.gather32 = torch.zeros(self.gather_total, dtype=torch.float32, device=dev)
            L.call("fs2_pack_weights", self._items_dev, len(gl), self.flat, self.gather32, 0)

    def version_token(self):
        """Sum of the parameters' autograd version counters: changes with every in-place update made through torch
        (optimizers, load_state_dict, p.copy_() ...); edits through `.data` are invisible to it."""
        return sum(p._version for p in self.params.values())

    def mark_mirror_current(self):
        self.mirror_valid = True
        self._mirror_token = self.version_token()

    def pack(self, bf16: bool):          # former name
        self.sync_operands(bf16)

    def operand(self, name, bf16: bool):
        """(buffer, element offset, row pitch) of a GEMM weight operand in the requested operand dtype."""
        pw = self.packed[name][1]
        if pw.gather:
            return (self.mirror, self.total + pw.goff, pw.k * pw.cin) if bf16 else (self.gather32, pw.goff, pw.k * pw.cin)
        return (self.mirror if bf16 else self.flat), pw.off, pw.k * pw.cin

    def adamw_gather(self):
        """The single gathered operand FusedAdamW keeps current inside its own pass: (src_off, src_ld, rows, cols, dst_off)
        in flat / mirror elements, or None."""
        gl = [pw for _, (_, pw) in self.packed.items() if pw.gather]
        if len(gl) != 1 or gl[0].k != 1:
            return None
        pw = gl[0]
        return (pw.off + pw.src_col0, pw.src_ld, pw.cout, pw.cin, self.total + pw.goff)

    def pw(self, name) -> PackedWeight:
        return self.packed[name][1]
