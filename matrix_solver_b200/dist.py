"""torch.distributed plumbing for the one-process-per-GPU launch (torchrun): rendezvous, the NCCL
unique-id hand-off for the C library, and max-over-ranks timing.  No data-path work here: the
panel broadcast and the solve exchanges run inside the CUDA library (csrc/ms_nccl.cu)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(
        os.environ.get("WORLD_SIZE", "1"))


_DEVICE = "cpu"


def init(backend=None):
    """Join the process group torchrun described (no-op at world size 1). Returns (rank, local, world)."""
    rank, local, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        global _DEVICE
        backend = backend or "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            _DEVICE = f"cuda:{local}"
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def broadcast_bytes(payload, nbytes, src=0):
    """Rank `src` passes `payload` (bytes of length nbytes); everyone returns the same bytes."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(payload)
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    t = t.to(_DEVICE)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def max_over_ranks(value):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_DEVICE)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_DEVICE)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
