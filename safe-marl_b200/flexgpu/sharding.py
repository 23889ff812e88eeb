"""Multi-GPU host logic of the env path: one process per GPU, environments sharded by index range,
no collective on the step path, ONE all-reduce (sum) of the 16-double statistics vector per flush
(the batched form of the per-episode means of madrl/models/model.py:247-265).

Kept free of CUDA so that the logic runs under the `gloo` backend in the CPU test-suite
(tests/test_sharding_gloo.py); on the GPU the same functions run over NCCL.
"""
import torch

from . import _lib


def shard(total_envs, rank, world):
    """Env index range [offset, offset + count) owned by `rank`: contiguous, balanced to within one
    32-env tile (tile boundaries keep every rank's batch aligned to the kernels' tile size).  The
    offset is also the Philox key offset (`env_offset`), so per-env results do not depend on `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    tiles = (total_envs + 31) // 32
    t0, t1 = tiles * rank // world, tiles * (rank + 1) // world
    lo, hi = min(t0 * 32, total_envs), min(t1 * 32, total_envs)
    return lo, hi - lo


def reduce_stats(vec, group=None):
    """Sum the statistics vector (len FP_NSTATS, fp64) over all ranks, in place; no-op without an
    initialised process group.  NCCL on device tensors, gloo on CPU tensors."""
    if vec.dtype != torch.float64 or vec.numel() != _lib.FP_NSTATS:
        raise ValueError("statistics vector must be %d float64 values" % _lib.FP_NSTATS)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(vec, op=torch.distributed.ReduceOp.SUM, group=group)
    return vec


def stats_dict(vec):
    return {k: vec[i] for i, k in enumerate(_lib.STAT_KEYS)}


def episode_means(stats):
    """Per-step means of every info key, as the reference reports them per episode (model.py:261-264),
    from the summed statistics of all envs and ranks."""
    n = float(stats["env_steps"])
    out = {("mean_train_" + k): (float(stats[k]) / n if n > 0 else 0.0)
           for k in _lib.INFO_KEYS + ("violation_count", "line_violation_count")}
    # The reference's own 'mean_train_reward' counts the reward twice: info carries a 'reward' key (the pre-penalty
    # reward, flexibility_provision_env.py:697) that is accumulated into the same slot as the step's reward
    # (model.py:247-252: sum(info['reward']) + sum(reward), the latter incl. -200 per failed step) before the
    # division by the step count -- reported separately so that curves can be compared with the reference's logs.
    fail_penalty = 200.0
    out["mean_train_reward_as_reference_logs_it"] = ((2.0 * float(stats["reward"]) - fail_penalty * float(stats["solver_failed"])) / n
                                                     if n > 0 else 0.0)
    return out
