"""Serving-side plumbing around the engine: data-parallel batch sharding and a host-buffer pipeline.

torch is used here for what it is good at -- pinned host memory, CUDA streams/events and `torch.distributed`
(NCCL over NVLink) -- and for nothing else: all model compute stays in libvitb200 behind the C-ABI.

Data parallel: images are independent, so the global batch is split contiguously, rank r
owning images [r*B, (r+1)*B); every rank holds a full weight replica; the forward has no communication; the only
collective is ONE in-place all-gather of the fp32 logits (`[B, classes]` per rank), and the classifier-head kernel
writes its logits directly into this rank's slice of the gather buffer, so compute -> collective needs no copy.
"""
from __future__ import annotations

import os

from . import _lib


def shard_bounds(global_batch: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of rank `rank`; global_batch must divide evenly (static shapes, vit.py:161-163)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def init_distributed(backend: str = "nccl"):
    """One process per GPU, rendezvous from the torchrun environment.  Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def gpu_numa_cpus(local_rank: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs: /sys/bus/pci/devices/<bus id>/numa_node -> node cpulist), or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception:
        return None


def bind_to_gpu_numa(local_rank: int):
    """Pin this process (and therefore the pinned host buffers it allocates afterwards, first touch) to the CPUs of its GPU's
    NUMA node: with eight ranks each streaming 154 MB of images per step, host reads that cross the socket interconnect are
    what bends the end-to-end scaling curve (GPUs 4-7 sit on node 1).  Returns the node or None when the topology is not
    visible / the allowed CPU set does not intersect it."""
    info = gpu_numa_cpus(local_rank)
    if info is None:
        return None
    node, cpus = info
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def all_gather_logits(gathered, local, world: int):
    """In-place all-gather: `local` is this rank's slice of `gathered` ([world*B, classes])."""
    if world > 1:
        import torch.distributed as dist
        dist.all_gather_into_tensor(gathered, local)
    return gathered


class DataParallel:
    """Batch-sharded replica group around one engine model (one instance per process / GPU)."""

    def __init__(self, model, per_rank_batch: int, image_hw, rank: int = 0, world: int = 1):
        import torch
        self.model, self.B, (self.h, self.w) = model, per_rank_batch, image_hw
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", model.device)
        self.gathered = torch.empty((world * per_rank_batch, model.num_classes), dtype=torch.float32, device=self.dev)
        self.local = self.gathered[rank * per_rank_batch:(rank + 1) * per_rank_batch]

    def forward_device(self, img_dev):
        """img_dev: this rank's shard, float32 NHWC CUDA tensor.  Returns the gathered logits (device tensor).
        Asynchronous on torch's current stream."""
        import torch
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        self.model.forward_raw(img_dev.data_ptr(), _lib.MEM_DEVICE, self.B, self.h, self.w, self.local.data_ptr(),
                               _lib.MEM_DEVICE, stream)
        return all_gather_logits(self.gathered, self.local, self.world)


def default_nccl_library():
    """The NCCL the engine should dlopen (`VB_NCCL_LIB`): the one bundled with torch when present, so that a process that
    also uses torch.distributed shares a single libnccl."""
    if os.environ.get("VB_NCCL_LIB"):
        return os.environ["VB_NCCL_LIB"]
    try:
        import nvidia.nccl  # noqa: F401  (namespace package of the torch wheels)
        cand = os.path.join(os.path.dirname(list(nvidia.nccl.__path__)[0] + "/"), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


class NativeDataParallel:
    """The same data-parallel group as `DataParallel`, but with the collective inside the C-ABI: `vb_dp_init` joins the
    handle to an NCCL communicator and `vb_forward_allgather` runs the forward and the in-place all-gather of the logits on one
    stream -- no torch.distributed on the data path.  The 128-byte communicator id travels by whatever the launcher has
    (here: an already initialised torch.distributed group of any backend, or `id_bytes` passed in by the caller).
    This is the path `bench.py` measures for N > 1; tools/dp_check.py checks it bit for bit against `DataParallel` and
    against the single-GPU forward of the whole batch."""

    def __init__(self, model, per_rank_batch: int, image_hw, rank: int = 0, world: int = 1, id_bytes: bytes | None = None):
        import ctypes as C
        import torch
        os.environ.setdefault("VB_NCCL_LIB", default_nccl_library())
        self.model, self.B, (self.h, self.w) = model, per_rank_batch, image_hw
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", model.device)
        lib = model._lib
        if id_bytes is None:
            import torch.distributed as dist
            buf = (C.c_char * 128)()
            if rank == 0:
                _lib.check(lib.vb_dp_unique_id(buf))
            box = [bytes(buf)]
            if world > 1:
                dist.broadcast_object_list(box, src=0)
            id_bytes = box[0]
        assert len(id_bytes) == 128
        model._finalize()
        _lib.check(lib.vb_dp_init(model._h, C.c_char_p(id_bytes), rank, world), model._h)
        self.gathered = torch.empty((world * per_rank_batch, model.num_classes), dtype=torch.float32, device=self.dev)
        self.local = self.gathered[rank * per_rank_batch:(rank + 1) * per_rank_batch]

    def forward_device(self, img_dev):
        import ctypes as C
        import torch
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.model._lib.vb_forward_allgather(self.model._h, C.c_void_p(img_dev.data_ptr()), _lib.MEM_DEVICE, self.B, self.h,
                                                        self.w, C.c_void_p(self.gathered.data_ptr()), C.c_void_p(stream)), self.model._h)
        return self.gathered


class HostPipeline:
    """Host batches in -> host logits out, the call a serving user makes.

    Step i copies its pinned-host image shard to the device on a copy stream while step i-1 computes
    (double-buffered device inputs), runs the forward (+ all-gather) on the compute stream and copies the
    gathered logits back to pinned host memory; `submit` returns the previous step's host logits."""

    def __init__(self, dp: DataParallel):
        import torch
        self.dp = dp
        dev = dp.dev
        self.copy_stream = torch.cuda.Stream(dev)
        self.compute_stream = torch.cuda.Stream(dev)
        self.in_dev = [torch.empty((dp.B, dp.h, dp.w, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.out_host = [torch.empty(dp.gathered.shape, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.step_done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.h2d_bytes = self.in_dev[0].numel() * 4
        self.d2h_bytes = self.out_host[0].numel() * 4

    def submit(self, img_host):
        """img_host: pinned float32 NHWC CPU tensor [B,h,w,3].  Returns host logits of the previous submit (or None).

        Buffer lifetime: the returned tensor is one of TWO pinned staging buffers; the NEXT call to `submit` queues the
        device-to-host copy that overwrites it (asynchronously), so consume or copy it before calling `submit` again.
        `img_host` may be refilled once the following `submit` has returned (its H2D copy has completed by then)."""
        import torch
        k = self.i & 1
        with torch.cuda.stream(self.copy_stream):
            if self.i >= 2:
                self.copy_stream.wait_event(self.step_done[k])      # compute(i-2) no longer reads in_dev[k]
            self.in_dev[k].copy_(img_host, non_blocking=True)
            self.h2d_done[k].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[k])
            g = self.dp.forward_device(self.in_dev[k])
            self.out_host[k].copy_(g, non_blocking=True)
            self.step_done[k].record(self.compute_stream)
        prev = None
        if self.i >= 1:
            self.step_done[k ^ 1].synchronize()
            prev = self.out_host[k ^ 1]
        self.i += 1
        return prev

    def flush(self):
        """Wait for and return the last submitted step's host logits."""
        if self.i == 0:
            return None
        k = (self.i - 1) & 1
        self.step_done[k].synchronize()
        return self.out_host[k]
