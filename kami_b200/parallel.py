"""Multi-GPU plumbing.  Training (BASELINE config 5): data-parallel over replay batches -- every rank
runs kb_trainer_forward_backward on its own mini-batch, the flat fp32 gradient vector is all-reduced over
NCCL / NVLink as ONE bucket (sum), and every replica applies the identical SGD step scaled by 1 / world.

Self-play path: games never interact (selfplay.cpp:97-200), so the
path shards by game with NO data-path collective.  torch.distributed is used only to launch one
process per GPU, to barrier around the timed region and to combine per-rank counters / device
times (max over ranks).  Backend-agnostic so the CPU tests can run it on gloo."""


def game_to_rank(game, world):
    """game g -> GPU g mod G."""
    return game % world


def local_games(total_games, world, rank):
    """Ids of the games rank `rank` owns; a partition of range(total_games)."""
    return list(range(rank, total_games, world))


def per_rank_seed(base_seed, rank):
    """Distinct RNG streams per shard (fixed per-tree seeds = seed0 + tree id inside a shard)."""
    return base_seed + 1000003 * rank


class Reducer:
    """max / sum of python floats over ranks; identity when not distributed."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist
        self.device = device

    def _all(self, x, op):
        if self.dist is None or not self.dist.is_initialized() or self.dist.get_world_size() == 1:
            return float(x)
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._all(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x):
        return self._all(x, self.dist.ReduceOp.SUM if self.dist else None)

    def barrier(self):
        if self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size() > 1:
            self.dist.barrier()


def whole_job_throughput(reducer, local_units, local_ms):
    """value = units all ranks processed / max-over-ranks device time (bench.py contract)."""
    units = reducer.sum(local_units)
    ms = reducer.max(local_ms)
    return units / (ms * 1e-3), units, ms


# ---- data-parallel training: the one collective of the system ------------------------------------------
class _DevArray:
    """A raw device pointer dressed as a CUDA array so torch can view it without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


def gradient_tensor(trainer, device):
    """torch fp32 view (no copy) of the trainer's flat gradient vector on `device`."""
    import torch

    ptr, n = trainer.grad_buffer()
    return torch.as_tensor(_DevArray(ptr, n), device=device)


def average_host_gradients(dist, grads):
    """The same reduction on a host array (gloo) -- what the CPU test of the N > 1 logic runs."""
    import torch

    t = torch.from_numpy(grads)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= dist.get_world_size()
    return t.numpy()


def data_parallel_step(trainer, dist, device, lib, obs_dev, pi_dev, z_dev, batch, lr, grad_view=None):
    """One data-parallel training step: local forward/backward, NCCL all-reduce(sum) of the gradient bucket,
    SGD with the 1/world scale.  The library runs on its own stream, torch's NCCL on torch's: both sides are
    synchronised around the collective.  Returns the gradient view for reuse."""
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    trainer.forward_backward_dev(obs_dev, pi_dev, z_dev, batch, want_loss=False)
    if world > 1:
        import torch

        if grad_view is None:
            grad_view = gradient_tensor(trainer, device)
        lib.kb_dev_sync()
        dist.all_reduce(grad_view, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    trainer.apply_sgd(lr, 1.0 / world)
    return grad_view
