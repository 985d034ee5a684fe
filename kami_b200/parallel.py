"""Multi-GPU plumbing of the self-play path: games never interact (selfplay.cpp:97-200), so the
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
