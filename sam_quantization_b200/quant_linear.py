"""QuantLinear -- drop-in for ``gptq_triton.quant_linear`` on B200.

Same constructor, buffer names/dtypes/shapes and checkpoint layout as the reference
(/root/reference/gptq_triton/quant_linear.py:66-116); ``forward`` launches the
hand-written sm_100a dequant-GEMM through the C ABI (``samq_qlinear_fwd``) instead of
the Triton ``matmul4_kernel`` (quant_linear.py:231-437).  Additive extensions
(SURVEY 8(b)): ``pack()`` (semantics of ``pack_linear``, gptq4sam.py:434-497), an
optional ``g_idx`` buffer (act-order groups) and bits in {2, 3, 4, 8}.

Differences from the reference, all deliberate:
  * outputs are freshly allocated (the reference returns views of one global 40 MiB
    workspace, quant_linear.py:13,431 -- not re-entrant, overflows for M*N > 20971520);
  * bias / GELU / residual are fused into the GEMM epilogue and added in fp32 before
    the single fp16 rounding (the reference rounds the GEMM to fp16, then adds the
    bias in fp16, quant_linear.py:352,434-435);
  * there is no autotune: tile shapes are static (``autotune_warmup`` is a no-op kept
    for API compatibility).
"""
from __future__ import annotations

import math
import weakref
from typing import Optional

import warnings

import torch
import torch.nn as nn

from . import _lib, ops

__all__ = ["QuantLinear", "make_quant", "matmul4", "triton_matmul4", "autotune_warmup", "pack_fields",
           "unpack_fields"]


def pack_fields(vals: torch.Tensor, bits: int) -> torch.Tensor:
    """Pack integer fields along dim 0 of ``vals[L, C]`` into int32 words.

    2/4/8 bit: word r = OR_j vals[r*f + j] << (bits*j), f = 32/bits, fields NOT masked
    (exactly what the reference's accumulation does, gptq4sam.py:472-477, 490-495 --
    a field of -1 sign-fills the rest of its word).  3 bit: 32 fields form a 96-bit
    little-endian stream over 3 words (quant.py:160-180); fields are masked to 3 bits.
    """
    assert vals.dim() == 2
    length, cols = vals.shape
    v = vals.to(torch.int64)
    if bits in (2, 4, 8):
        f = 32 // bits
        assert length % f == 0, "field count must fill whole int32 words"
        v = v.view(length // f, f, cols)
        words = torch.zeros((length // f, cols), dtype=torch.int64)
        for j in range(f):
            words |= v[:, j, :] << (bits * j)
        words &= 0xFFFFFFFF
    elif bits == 3:
        assert length % 32 == 0, "3-bit packing needs a multiple of 32 fields"
        v = (v & 7).view(length // 32, 32, cols)
        w = [torch.zeros((length // 32, cols), dtype=torch.int64) for _ in range(3)]
        for j in range(32):
            p = 3 * j
            word, off = p // 32, p % 32
            w[word] |= (v[:, j, :] << off) & 0xFFFFFFFF
            if off + 3 > 32:
                w[word + 1] |= v[:, j, :] >> (32 - off)
        words = torch.stack(w, dim=1).reshape(length // 32 * 3, cols)
    else:
        raise NotImplementedError("Only 2,3,4,8 bits are supported.")
    # two's-complement reinterpretation of the low 32 bits
    words = torch.where(words >= 2**31, words - 2**32, words)
    return words.to(torch.int32)


def unpack_fields(words: torch.Tensor, bits: int) -> torch.Tensor:
    """Inverse of :func:`pack_fields` for properly masked fields: int32 words ``[R, C]`` ->
    int64 fields ``[R * 32 / bits, C]`` (CPU; one-time checkpoint re-layout only)."""
    w = words.detach().cpu().to(torch.int64) & 0xFFFFFFFF
    rows, cols = w.shape
    if bits in (2, 4, 8):
        f = 32 // bits
        out = torch.stack([(w >> (bits * j)) & ((1 << bits) - 1) for j in range(f)], dim=1)
        return out.reshape(rows * f, cols)
    if bits == 3:
        assert rows % 3 == 0
        w = w.view(rows // 3, 3, cols)
        fields = []
        for j in range(32):
            p = 3 * j
            word, off = p // 32, p % 32
            v = w[:, word] >> off
            if off + 3 > 32:
                v = v | (w[:, word + 1] << (32 - off))
            fields.append(v & 7)
        return torch.stack(fields, dim=1).reshape(rows // 3 * 32, cols)
    raise NotImplementedError("Only 2,3,4,8 bits are supported.")


def _no_chain():
    return None


class WeightPrefetchChain:
    """Weight prefetch over a fixed execution order of int4 ``QuantLinear`` layers (the four linears of
    every encoder block).  For long M a layer runs as "unpack the weight once into an fp16 scratch, then
    a dense tcgen05 GEMM"; done back to back, the 6 us unpack kernel sits BETWEEN two GEMMs with the GPU
    nearly idle.  With a chain, right after layer i's GEMM has been enqueued the unpack of layer i + 1
    is launched programmatically behind it (``samq_qlinear_prefetch``): a small persistent grid that
    runs next to that GEMM and writes one of ``kBuffers`` rotating scratch buffers; layer i + 1 then
    only runs its GEMM.  Same kernels' arithmetic, same results bit for bit.

    Why the early writes are safe: a scratch buffer is rewritten kBuffers - 1 prefetches after the GEMM
    that read it, and in the encoder at most two GEMMs follow each other without a normally launched
    kernel (LayerNorm, attention) between them, which completes everything before it.  The buffers are
    allocated once, before any operand of a kernel that can overlap a prefetch (so the caching
    allocator cannot hand them memory such a kernel still reads), and never re-allocated.

    State lives here and in a weak map, not on the modules: ``state_dict``, ``deepcopy`` and pickling of
    the model are unaffected."""

    kBuffers = 4

    def __init__(self, layers):
        self.layers = list(layers)
        self.key = tuple(id(m) for m in self.layers)
        self.bufs = None
        self.slot = 0
        self.ready = None          # (index, buffer, weight version key) of the one prefetched layer
        for i, m in enumerate(self.layers):
            _PREFETCH[m] = (self, i)

    def __deepcopy__(self, memo):   # a copied encoder links its own chain on first use
        return None

    def __reduce__(self):           # ... and so does an unpickled one (no scratch buffers in the pickle)
        return (_no_chain, ())

    @staticmethod
    def _version(m):
        return (m.qweight.data_ptr(), m.qweight._version, m.qzeros.data_ptr(), m.qzeros._version,
                m.scales.data_ptr(), m.scales._version)

    @staticmethod
    def _prefetchable(m) -> bool:
        return (isinstance(m, QuantLinear) and m.bits == 4 and m.g_idx is None and m.groupsize % 16 == 0
                and m.infeatures % 64 == 0 and m.outfeatures % 8 == 0 and m.qweight.is_cuda)

    def begin(self, device) -> None:
        """Start of a pass over the chain: drop stale state, make sure the scratch buffers exist."""
        self.ready = None
        numel = max((m.infeatures * m.outfeatures for m in self.layers if self._prefetchable(m)), default=0)
        if numel and (self.bufs is None or self.bufs[0].numel() < numel or self.bufs[0].device != device):
            self.bufs = [torch.empty(numel, dtype=torch.float16, device=device) for _ in range(self.kBuffers)]
            if not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(device).synchronize()   # one-time: nothing in flight reads that memory

    def take(self, i: int):
        """The scratch holding layer i's unpacked weight if it was prefetched (and is still current)."""
        r, self.ready = self.ready, None
        if r is not None and r[0] == i and r[2] == self._version(self.layers[i]):
            return r[1]
        return None

    def after(self, i: int, rows: int) -> None:
        """Layer i's GEMM has just been enqueued: unpack layer i + 1 next to it."""
        nxt = i + 1
        if (nxt >= len(self.layers) or self.bufs is None or rows < ops.PREFETCH_MIN_M
                or not _lib.OPTIONS["prefetch"] or _lib.OPTIONS["gemm"] == "fused"):
            return
        m = self.layers[nxt]
        if not self._prefetchable(m) or m.qweight.device != self.bufs[0].device \
                or m.infeatures * m.outfeatures > self.bufs[0].numel():
            return
        buf = self.bufs[self.slot]
        self.slot = (self.slot + 1) % self.kBuffers
        ops.qlinear_prefetch(m.qweight, m.qzeros, m.scales, m.bits, m.groupsize, buf)
        self.ready = (nxt, buf, self._version(m))


_PREFETCH = weakref.WeakKeyDictionary()   # QuantLinear -> (WeightPrefetchChain, index)


class QuantLinear(nn.Module):
    """GPTQ-packed linear layer (reference: quant_linear.py:66-116)."""

    def __init__(self, bits: int, groupsize: int, infeatures: int, outfeatures: int, bias: bool):
        super().__init__()
        if bits not in (2, 3, 4, 8):
            # the reference supports 4 only (quant_linear.py:72-73); 2/3/8 are extensions
            raise NotImplementedError("Only 2, 3, 4 and 8 bits are supported.")
        groupsize = infeatures if groupsize == -1 else groupsize

        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.bits = bits
        self.groupsize = groupsize

        if bits == 3:
            assert infeatures % 32 == 0 and outfeatures % 32 == 0, \
                "3-bit packing needs infeatures and outfeatures to be multiples of 32"
            rows, zcols = infeatures // 32 * 3, outfeatures // 32 * 3
        else:
            features_per_int = 32 // bits
            assert outfeatures % features_per_int == 0, \
                "outfeatures must be a multiple of features_per_int"   # quant_linear.py:84-86
            rows, zcols = infeatures // features_per_int, outfeatures // features_per_int
        groups = math.ceil(infeatures / groupsize)
        self.register_buffer("qweight", torch.empty((rows, outfeatures), dtype=torch.int32))
        self.register_buffer("qzeros", torch.empty((groups, zcols), dtype=torch.int32))
        self.register_buffer("scales", torch.empty((groups, outfeatures), dtype=torch.float16))
        if bias:
            self.register_buffer("bias", torch.empty(outfeatures, dtype=torch.float16))
        else:
            self.register_parameter("bias", None)
        # act-order extension: absent (None) means contiguous groups k // groupsize
        self.register_buffer("g_idx", None)

    # ------------------------------------------------------------------ forward
    def sorted_pack(self):
        """Act-order layers on the fused in-SM dequant path: ``(perm, qweight_sorted)`` with the
        input features re-ordered so that every group is contiguous -- ``perm = argsort(g_idx)``
        (stable), ``qweight_sorted`` = the same integer fields, rows ``perm`` -- so that
        ``x[:, perm] @ dequant(qweight_sorted, contiguous groups) == x @ dequant(qweight, g_idx)``
        (same products, summed in a different order).  Needs every group to have exactly
        ``groupsize`` members, which GPTQ's act-order guarantees (``g_idx = invperm // groupsize``);
        returns ``None`` otherwise.  Computed once per weight version (CPU re-pack), cached."""
        if self.g_idx is None:
            return None
        key = (self.qweight.data_ptr(), self.qweight._version, self.g_idx.data_ptr(), self.g_idx._version)
        hit = getattr(self, "_sorted_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        g = self.g_idx.detach().cpu().to(torch.int64)
        groups = self.infeatures // self.groupsize
        ok = (self.infeatures % self.groupsize == 0 and int(g.min()) >= 0 and int(g.max()) < groups
              and bool((torch.bincount(g, minlength=groups) == self.groupsize).all()))
        result = None
        if ok:
            perm = torch.argsort(g, stable=True)
            fields = unpack_fields(self.qweight, self.bits)[perm]
            dev = self.qweight.device
            result = (perm.to(torch.int32).to(dev), pack_fields(fields, self.bits).contiguous().to(dev))
        self._sorted_cache = (key, result)
        return result

    def _use_sorted(self, rows: int):
        """(perm, qweight_sorted) when the call should take gather + fused kernel, else None."""
        if self.g_idx is None or rows >= ops.TWO_KERNEL_MIN_M or _lib.OPTIONS["gemm"] == "dense" \
                or self.groupsize % 64 != 0:
            return None
        return self.sorted_pack()

    def forward(self, x: torch.Tensor, epilogue: int = _lib.EPI_NONE,
                residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        sp = self._use_sorted(x.numel() // max(1, x.shape[-1])) if x.is_cuda else None
        if sp is not None:      # act-order, short M: gather x's columns, contiguous groups, fused kernel
            return ops.qlinear(ops.gather_cols(x, sp[0]), sp[1], self.qzeros, self.scales, self.bits,
                               self.groupsize, self.bias, None, epilogue, residual)
        pf = _PREFETCH.get(self) if x.is_cuda else None
        wt = pf[0].take(pf[1]) if pf else None
        y = ops.qlinear(x, self.qweight, self.qzeros, self.scales, self.bits, self.groupsize,
                        self.bias, self.g_idx, epilogue, residual, wt_ready=wt)
        if pf:
            pf[0].after(pf[1], x.numel() // max(1, x.shape[-1]))
        return y

    def forward_unpartition(self, x: torch.Tensor, shortcut: torch.Tensor, window_size: int) -> torch.Tensor:
        """``shortcut + window_unpartition(self(x))`` in one kernel (x: windowed tokens)."""
        sp = self._use_sorted(x.numel() // max(1, x.shape[-1])) if x.is_cuda else None
        if sp is not None:
            return ops.qlinear_unpartition(ops.gather_cols(x, sp[0]), sp[1], self.qzeros, self.scales, self.bits,
                                           self.groupsize, self.bias, shortcut, window_size, None)
        pf = _PREFETCH.get(self) if x.is_cuda else None
        wt = pf[0].take(pf[1]) if pf else None
        y = ops.qlinear_unpartition(x, self.qweight, self.qzeros, self.scales, self.bits, self.groupsize,
                                    self.bias, shortcut, window_size, self.g_idx, wt_ready=wt)
        if pf:
            pf[0].after(pf[1], x.numel() // max(1, x.shape[-1]))
        return y

    def forward_partition(self, x: torch.Tensor, window_size: int) -> torch.Tensor:
        """``self(window_partition(x))`` in one kernel: ``x[B, H, W, K]`` in image order ->
        ``[B*nWin, ws, ws, N]``; the zero-padding tokens are not multiplied (their rows = bias)."""
        sp = self._use_sorted(x.numel() // max(1, x.shape[-1])) if x.is_cuda else None
        if sp is not None:
            return ops.qlinear_partition(ops.gather_cols(x, sp[0]), sp[1], self.qzeros, self.scales, self.bits,
                                         self.groupsize, self.bias, window_size, None)
        pf = _PREFETCH.get(self) if x.is_cuda else None
        wt = pf[0].take(pf[1]) if pf else None
        y = ops.qlinear_partition(x, self.qweight, self.qzeros, self.scales, self.bits, self.groupsize,
                                  self.bias, window_size, self.g_idx, wt_ready=wt)
        if pf:
            pf[0].after(pf[1], x.numel() // max(1, x.shape[-1]))
        return y

    def dequantize(self, transposed: bool = False) -> torch.Tensor:
        """fp16 ``W[K, N]`` (``[N, K]`` if transposed) via ``samq_unpack_dequant``."""
        return ops.unpack_dequant(self.qweight, self.qzeros, self.scales, self.bits, self.groupsize,
                                  self.g_idx, transposed)

    # --------------------------------------------------------------------- pack
    @torch.no_grad()
    def pack(self, linear, scales: torch.Tensor, zeros: torch.Tensor,
             g_idx: Optional[torch.Tensor] = None) -> None:
        """Pack fake-quantised weights into this layer's buffers.

        ``linear``: an ``nn.Linear`` (weight ``[N, K]`` + bias) or a weight tensor ``[N, K]``;
        ``scales``, ``zeros``: ``[N, G]`` as produced by the GPTQ ``Quantizer``.
        Semantics of ``pack_linear`` (gptq4sam.py:434-497): integer grid
        ``round((W + zero*scale) / scale)`` in fp32, fields LSB-first, ``qzeros`` stores
        ``zero - 1``.  Runs on the CPU (one-time, offline) like the reference.
        """
        if isinstance(linear, nn.Module):
            weight, lbias = linear.weight.data, (linear.bias.data if linear.bias is not None else None)
        else:
            weight, lbias = linear, None
        weight = weight.detach().cpu()
        n, k = weight.shape
        assert (n, k) == (self.outfeatures, self.infeatures), "weight shape does not match the layer"
        scales_t = scales.detach().cpu().to(torch.float32).t().contiguous()   # [G, N]
        zeros_t = zeros.detach().cpu().to(torch.float32).t().contiguous()     # [G, N]
        scale_zeros = zeros_t * scales_t
        if g_idx is None:
            gi = torch.arange(k) // self.groupsize
        else:
            gi = g_idx.detach().cpu().to(torch.int64)
            assert gi.numel() == k
        wf = weight.to(torch.float32).t()                                      # [K, N]
        intweight = torch.round((wf + scale_zeros[gi]) / scales_t[gi]).to(torch.int32)
        dev = self.qweight.device
        self.qweight = pack_fields(intweight, self.bits).contiguous().to(dev)
        zeros_m1 = (zeros_t - 1).to(torch.int32)
        if bool((zeros_m1 < 0).any()):
            # zero == 0 is not representable: the checkpoint stores zero-1 (gptq4sam.py:484), the
            # reference packs the -1 unmasked and its sign bits overwrite the neighbouring columns'
            # zero points.  Bug-compatible bits are kept (tests pin them), but say so.
            warnings.warn(f"QuantLinear.pack: {int((zeros_m1 < 0).sum())} group zero points are 0; the reference's "
                          f"`zero - 1` storage cannot hold them and corrupts neighbouring qzeros fields "
                          f"(use sym=True or a Quantizer whose grid keeps zero >= 1)", RuntimeWarning)
        self.qzeros = pack_fields(zeros_m1.t().contiguous(), self.bits).t().contiguous().to(dev)
        self.scales = scales_t.to(torch.float16).to(dev)
        if self.bias is not None:
            # a bare weight tensor carries no bias: zero it instead of leaving torch.empty garbage
            self.bias = (torch.zeros(n, dtype=torch.float16) if lbias is None
                         else lbias.detach().clone().to(torch.float16)).to(dev)
        self.g_idx = None if g_idx is None else gi.to(torch.int32).to(dev)

    def extra_repr(self) -> str:
        return (f"bits={self.bits}, groupsize={self.groupsize}, infeatures={self.infeatures}, "
                f"outfeatures={self.outfeatures}, bias={self.bias is not None}, "
                f"g_idx={self.g_idx is not None}")


def make_quant(model: nn.Module, bits: int, groupsize: int) -> None:
    """Replace every ``nn.Linear`` in ``model`` by a ``QuantLinear`` (quant_linear.py:15-36);
    a module literally named ``lm_head`` is skipped like in the reference (:24-25)."""
    for name, m in list(model.named_modules()):
        if not isinstance(m, nn.Linear):
            continue
        if name == "lm_head":
            continue
        qlayer = QuantLinear(bits, groupsize, m.in_features, m.out_features, m.bias is not None)
        if "." in name:
            parent_name, child = name.rsplit(".", 1)
            parent = model.get_submodule(parent_name)
        else:
            parent, child = model, name
        setattr(parent, child, qlayer)


def matmul4(groupsize: int, a: torch.Tensor, qweight: torch.Tensor, scales: torch.Tensor,
            qzeros: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``C = A x dequant(B) + bias`` for int4 -- the call signature of the reference's
    ``triton_matmul4`` (quant_linear.py:355-437), including its shape assertions."""
    assert a.shape[-1] == qweight.shape[0] * 8, "A must be a multiple of 8 in the last dimension"
    assert a.is_contiguous(), "A must be contiguous"
    K, N = a.shape[-1], qweight.shape[1]
    assert K % 128 == 0, "K must be a multiple of 16, 32, 64, and 128"
    assert N % 256 == 0, "N must be a multiple of 16, 32, 64, 128, and 256"
    assert groupsize % 128 == 0, "groupsize must be a multiple of 32, 64, and 128"
    return ops.qlinear(a, qweight, qzeros, scales, 4, groupsize, bias)


# the reference's public name (gptq_triton/__init__.py:12); nothing here uses Triton
triton_matmul4 = matmul4


def autotune_warmup(model: nn.Module):
    """API-compatibility no-op: the CUDA kernels have static tile shapes, nothing to tune
    (reference: quant_linear.py:39-63 returns one warm-up closure per unique (K, N))."""
    return iter(())
