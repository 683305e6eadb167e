"""Data-parallel launcher: shard the image batch over the GPUs of one box.

The reference has no multi-GPU inference path (device hard-coded, batch asserted 1:
/root/reference/gptq4sam_infer.py:169,176).  Images are independent -- no op
of the encoder crosses the batch dimension (image_encoder.py:106-118) -- so the path
shards as REPLICAS: one process per GPU (torchrun), full copy of the packed weights
(ViT-H int4 g128 = 328 MB) per rank, contiguous split of the global batch, and no
collective on the hot path.  The only communication is an optional final
``all_gather`` of the embeddings ``[B_local, 256, 64, 64]`` (2 MiB/image in fp16) over
NCCL / NVLink, outside the per-block loop.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "init_distributed", "ShardedEncoder", "GraphedEncoder"]


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n`` items: the first ``n % world_size`` ranks get one extra."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the default
    process group when WORLD_SIZE > 1 (NCCL on GPUs, gloo on CPU)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


class ShardedEncoder:
    """Runs ``encoder`` on this rank's contiguous shard of a global batch."""

    def __init__(self, encoder, rank: int = 0, world_size: int = 1, micro_batch: Optional[int] = None):
        self.encoder = encoder
        self.rank = rank
        self.world_size = world_size
        self.micro_batch = micro_batch

    def local_slice(self, global_batch: int) -> slice:
        lo, hi = shard_bounds(global_batch, self.world_size, self.rank)
        return slice(lo, hi)

    @torch.no_grad()
    def encode_local(self, images: torch.Tensor) -> torch.Tensor:
        """Encode already-local images, optionally in micro-batches."""
        mb = self.micro_batch or max(1, images.shape[0])
        outs: List[torch.Tensor] = [self.encoder(images[i:i + mb]) for i in range(0, images.shape[0], mb)]
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    @torch.no_grad()
    def __call__(self, images_global: torch.Tensor, gather: bool = False) -> torch.Tensor:
        """``images_global`` is the full batch (same on every rank); returns this rank's
        embeddings, or -- with ``gather`` -- the full batch's embeddings on every rank."""
        n = images_global.shape[0]
        local = self.encode_local(images_global[self.local_slice(n)])
        if not gather or self.world_size == 1:
            return local
        sizes = [shard_bounds(n, self.world_size, r) for r in range(self.world_size)]
        max_len = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((max_len,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        bufs = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(bufs, pad)
        return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


class GraphedEncoder:
    """CUDA-graph replay of a (quantized, fused) encoder for one fixed input shape.

    The 32-block loop is ~230 kernel launches per step; at small batch the launch gaps are
    visible, so the whole forward is captured once (static input/output buffers, TMA
    descriptors keyed on the stable buffer addresses) and replayed.  ``__call__`` copies the
    input into the static buffer (device-to-device, or host-to-device for pinned input) and
    returns the static output tensor (valid until the next call); ``run_host`` is the pipelined
    host-buffer form.
    """

    def __init__(self, encoder, example: torch.Tensor, warmup: int = 2):
        assert example.is_cuda, "GraphedEncoder needs a CUDA example input"
        self.encoder = encoder
        self.static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                encoder(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        from . import _lib

        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = encoder(self.static_in)
        #: libsamq kernels recorded in the graph (= launched by every replay)
        self.kernels_per_replay = _lib.launch_count() - before

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        self.static_in.copy_(images, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def _init_pipeline(self) -> None:
        dev = self.static_in.device
        self._h2d, self._d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._stage_in = [torch.empty_like(self.static_in) for _ in range(2)]
        self._stage_out = [torch.empty_like(self.static_out) for _ in range(2)]
        self._in_ready, self._in_free, self._out_ready, self._out_free = (
            [torch.cuda.Event() for _ in range(2)] for _ in range(4))
        self._n = 0

    @torch.no_grad()
    def run_host(self, host_in: torch.Tensor, host_out: torch.Tensor) -> torch.cuda.Event:
        """Serving-loop form of ``__call__`` for PINNED host buffers: asynchronous, and pipelined
        across consecutive calls.  The host-to-device copy of call i+1 and the device-to-host copy
        of call i-1 run on two copy streams underneath the encoder of call i (staging buffers are
        double-buffered; a device-to-device copy moves them in and out of the graph's static
        tensors, 0.2 ms per call at batch 32 against ~5 ms of PCIe time).  Returns the event after
        which ``host_out`` holds this call's embeddings (``host_in`` may be overwritten by then
        too)."""
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise ValueError("run_host needs pinned host tensors (torch.Tensor.pin_memory())")
        if not hasattr(self, "_h2d"):
            self._init_pipeline()
        k, first = self._n & 1, self._n < 2
        self._n += 1
        cur = torch.cuda.current_stream(self.static_in.device)
        with torch.cuda.stream(self._h2d):
            if not first:
                self._h2d.wait_event(self._in_free[k])      # the encoder two calls back has consumed it
            self._stage_in[k].copy_(host_in, non_blocking=True)
            self._in_ready[k].record(self._h2d)
        cur.wait_event(self._in_ready[k])
        self.static_in.copy_(self._stage_in[k])
        self._in_free[k].record(cur)
        self.graph.replay()
        if not first:
            cur.wait_event(self._out_free[k])               # its previous content has reached the host
        self._stage_out[k].copy_(self.static_out)
        self._out_ready[k].record(cur)
        with torch.cuda.stream(self._d2h):
            self._d2h.wait_event(self._out_ready[k])
            host_out.copy_(self._stage_out[k], non_blocking=True)
            self._out_free[k].record(self._d2h)
        return self._out_free[k]
