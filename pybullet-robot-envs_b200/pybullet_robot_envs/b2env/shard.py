"""Multi-GPU plumbing: the environment batch shards embarrassingly (environments never interact —
one physics client per env in the reference, panda_push_gym_env.py:62), so each rank owns a
contiguous block of environments and its own simulation.  The only traffic of the path is the
gather of per-env episode returns / lengths at report time."""
import torch
import torch.distributed as dist


def shard_range(global_envs, rank, world):
    """Contiguous block [lo, hi) of environment indices owned by ``rank``."""
    per = (global_envs + world - 1) // world
    lo = min(rank * per, global_envs)
    hi = min(lo + per, global_envs)
    return lo, hi


def env_seed(base_seed, global_env_index):
    """Seed of one environment: independent of how the batch is sharded."""
    return int(base_seed) + int(global_env_index)


def gather_returns(local_returns):
    """all_gather of the per-env episode returns (float32 [B_local]) -> [B_global] on every rank.
    NCCL over NVLink on GPUs; gloo in the CPU tests."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_returns
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_returns.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_returns.numel()], dtype=torch.int64, device=local_returns.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(mx, dtype=local_returns.dtype, device=local_returns.device)
    pad[: local_returns.numel()] = local_returns
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(out, sizes)])


def max_over_ranks(value, device):
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
