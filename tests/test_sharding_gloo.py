"""CPU, world_size 2 over gloo: the N > 1 host logic of the env path -- index-range sharding with the
global env id as Philox key offset, no exchange on the step path, one all-reduce of the statistics
vector -- with the C mirror standing in for the kernels (same per-env arithmetic, same Philox draws)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, STEPS, SEED = 200, 6, 9          # ragged: 7 tiles, shards of 96 and 104 envs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_shard(rank, world, offset, count, acts):
    from oracle import c_mirror, env_ref, ieee33
    from flexgpu import DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
    prof = synthetic_profiles(Network(create_network(DEFAULT_ENV_ARGS)), 5, T=1500, seed=0)
    fonet = c_mirror.make_net(ieee33.tree_arrays(ieee33.create_network()), env_ref.DEFAULT_ARGS, env_ref.DEFAULT_ARGS["buildings"])
    mb = c_mirror.MirrorBatch(fonet, prof.as_dict(), count)
    mb.reset_random(SEED, env_offset=offset)
    stats = np.zeros(16)
    rewards = []
    for t in range(STEPS):
        r, d, info = mb.step(acts[t, offset:offset + count])
        rewards.append(r.copy())
        stats[:8] += info.sum(0); stats[8] += mb.vcount.sum(); stats[9] += count; stats[10] += d.sum()
    return np.array(rewards), mb.V.copy(), stats


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from flexgpu import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acts = np.random.default_rng(1).uniform(0, 1, (STEPS, N_TOTAL, 20)).astype(np.float32)
    offset, count = sharding.shard(N_TOTAL, rank, world)
    rewards, V, stats = _run_shard(rank, world, offset, count, acts)
    vec = torch.from_numpy(stats.copy())
    sharding.reduce_stats(vec)                       # the one collective
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), offset=offset, count=count, rewards=rewards, V=V,
             local=stats, reduced=vec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_align():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
    from flexgpu import sharding
    for total in (1, 31, 32, 200, 65536, 2 ** 20, 2 ** 20 + 5):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1 and o1 % 32 == 0 or o1 == total          # contiguous, tile aligned
    assert sharding.shard(2 ** 20, 3, 8) == (3 * 131072, 131072)             # BASELINE config 5
    with pytest.raises(ValueError):
        sharding.shard(10, 2, 2)


def test_two_ranks_equal_one_batch_and_stats_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    # the same batch on one "GPU"
    sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
    acts = np.random.default_rng(1).uniform(0, 1, (STEPS, N_TOTAL, 20)).astype(np.float32)
    rewards, V, stats = _run_shard(0, 1, 0, N_TOTAL, acts)
    assert [int(p["offset"]) for p in parts] == [0, 96] and [int(p["count"]) for p in parts] == [96, 104]
    assert np.array_equal(np.concatenate([p["rewards"] for p in parts], axis=1), rewards)   # per-env results do not depend on world
    assert np.array_equal(np.concatenate([p["V"] for p in parts], axis=0), V)
    for p in parts:                                                                         # every rank holds the global sums
        assert np.array_equal(p["reduced"], parts[0]["local"] + parts[1]["local"])
    assert np.allclose(parts[0]["reduced"], stats, rtol=1e-13, atol=0)                      # == one-batch totals (summation order differs)
    assert parts[0]["reduced"][9] == N_TOTAL * STEPS
    from flexgpu import sharding
    means = sharding.episode_means(sharding.stats_dict(torch.from_numpy(parts[0]["reduced"])))
    assert abs(means["mean_train_reward"] - rewards.mean()) < 1e-12 * max(1.0, abs(rewards.mean())) + 1e-15
