"""Host-side helpers for one-process-per-GPU launches (torchrun).  The block-compression path has
no data-path collective: one stream is sharded over the GPUs by the process that owns the stream
(bnz_ctx_create(n_gpus=N)), independent objects are encoded by independent ranks.  Ranks share
only a timing barrier and a max-over-ranks reduction (NCCL on GPUs, gloo in the CPU tests)."""
import os


def env():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def object_seed(base_seed, rank):
    """every rank encodes its own object: distinct, reproducible corpus seed per rank"""
    return (base_seed + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF if rank else base_seed


class Group:
    """thin wrapper over torch.distributed that also works for world_size == 1 without torch"""

    def __init__(self, backend=None, device=None):
        self.rank, self.world, self.local = env()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch = torch
            self.dist = dist
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend or "gloo", **kw)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64,
                              device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64,
                              device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()


class CpuGate:
    """A barrier the waiting ranks sit in on the CPU (gloo).  While rank 0 drives every GPU of the
    box for the one-stream measurement the other ranks must not wait inside an NCCL barrier: its
    kernel would spin on the very SMs being measured."""

    def __init__(self, group):
        self.dist = group.dist
        self.pg = None
        if self.dist is not None:
            self.pg = self.dist.new_group(backend="gloo")

    def wait(self):
        if self.pg is not None:
            self.dist.barrier(group=self.pg)

    def close(self):
        self.pg = None


def aggregate_throughput(bytes_per_rank, world, elapsed_max_s, steps):
    """whole-job MB/s: every rank processed `bytes_per_rank` per step; time = max over ranks"""
    return world * bytes_per_rank * steps / elapsed_max_s / 1e6
