"""Batched rollout in the shape of the reference's train_process (madrl/models/model.py:198-267), with the
four host<->device crossings per step removed: observations, actions, rewards and the transition sink stay
on the GPU; only the episode statistics leave it (one all-reduce when several ranks run).

    python examples/rollout.py --envs 65536 --steps 95
    torchrun --nproc-per-node 8 examples/rollout.py --envs 1048576      # 2^20 envs sharded over 8 GPUs

The policy here is a stand-in (a fixed random affine map of the observation squashed by tanh): the point is
the env-side API a learner plugs into.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="total environments over all ranks")
    ap.add_argument("--steps", type=int, default=95)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    offset, count = sharding.shard(args.envs, rank, world)             # env index range of this rank
    env = BatchedFlexProvisionEnv(None, n_envs=count, device=dev, seed=0, env_offset=offset)
    obs, state = env.reset()                                           # [N,5,144] f32, [N,110] f32
    g = torch.Generator(device=dev).manual_seed(1)
    Wp = torch.randn(env.obs_size, env.n_actions, device=dev, generator=g) * 0.05
    noise = 0.1 * torch.randn(8, count * env.n_agents, env.n_actions, device=dev, generator=g)
    torch.tanh(obs.view(-1, env.obs_size) @ Wp)                        # warm-up: cuBLAS handle, module load
    env.episode_stats(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(args.steps):
        action = torch.tanh(obs.view(-1, env.obs_size) @ Wp) + noise[t % 8]
        # translate_action (util.py:121-129) and the pushing get_obs (quirk Q7) are fused into the step kernel
        reward, done, info, obs = env.step(action, translate=True, want_info=False, return_obs=True)
        if (t + 1) % (env.episode_limit - 1) == 0:                     # every env of the batch ends after 95 steps (Q1):
            env.reset(mask=done.to(torch.uint8), return_obs=False)     # auto-reset the finished ones (Philox streams)
            obs = env.get_obs()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    stats = env.episode_stats()                                        # the one collective (NCCL all-reduce of 16 doubles)
    if rank == 0:
        means = sharding.episode_means(stats)
        print(f"{args.envs} envs x {args.steps} steps on {world} GPU(s): {args.envs * args.steps / dt:.3e} env-steps/s "
              f"(policy + step + get_obs)")
        for k in ("mean_train_reward", "mean_train_revenue", "mean_train_voltage_penalty", "mean_train_solver_failed"):
            print(f"  {k:32s} {means[k]: .6f}")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
