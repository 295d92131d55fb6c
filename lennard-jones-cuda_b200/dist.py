"""One-process-per-GPU plumbing over torch.distributed (NCCL on GPUs, gloo in CPU tests).

The data path never goes through torch: each rank's C-ABI handle owns its own NCCL communicator
(ljmd_create_distributed) and issues the per-step all-gather of positions and all-reduce of the
energy / virial / kinetic sums itself.  torch.distributed only carries the 128-byte ncclUniqueId from
rank 0 to the others, the barriers around timed regions, and the max-over-ranks of timings.
"""
import os


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


BLOCK = 512   # particles per block of the force kernels (kITile in csrc/ljmd_core.cu)


def shard_bounds(N, rank, world):
    """Contiguous i-shard of `rank`, in whole 512-particle blocks: the same rule as the library (ljmd_plan)."""
    nblk = (N + BLOCK - 1) // BLOCK
    cnt = ((nblk + world - 1) // world) * BLOCK
    return min(N, rank * cnt), min(N, (rank + 1) * cnt)


def init(backend=None):
    """Initialise the default process group from the torchrun environment.  Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    rank, world, local_rank = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def share_unique_id(make_id):
    """Rank 0 calls make_id() (-> 128 bytes); every rank returns the same bytes."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return make_id()
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = bytes(box[0])
    if len(uid) != 128:
        raise ValueError(f"ncclUniqueId must be 128 bytes, got {len(uid)}")
    return uid


def all_gather_bytes(b):
    """Every rank contributes a bytes object; every rank gets the list in rank order."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(b)]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(b))
    return [bytes(x) for x in out]


def connect_fabric(system):
    """Map every rank's window into every other rank (CUDA IPC) so the per-step exchange uses NVLink peer
    memory.  Returns True when the fabric is on; False (NCCL stays the transport) when LJMD_COMM=nccl or the
    mapping fails on any rank."""
    import torch
    import torch.distributed as dist

    if system.world == 1 or os.environ.get("LJMD_COMM", "p2p") == "nccl":
        return False
    handles = all_gather_bytes(system.fabric_handle())
    ok = 1
    try:
        system.fabric_connect(handles)
    except Exception:
        ok = 0
    t = torch.tensor([ok], device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 0:
        raise RuntimeError("fabric connect failed on some rank; set LJMD_COMM=nccl to run over NCCL only")
    return True


def barrier():
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(x):
    """Timings are reported as the max over ranks (the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def finalize():
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()
