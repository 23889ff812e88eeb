"""Boundary glue of the rollout loop, mirroring utils/util.py of the reference (same names and
argument meaning) but staying on the device: nothing here copies to the host.

  translate_action(args, action, env)   utils/util.py:121-133
  prep_obs(state)                       utils/util.py:135-145

In the fast path neither is needed: `BatchedFlexProvisionEnv.step(actions, translate=True)` applies
translate_action inside the step kernel and `get_obs()` already returns the fp32 tensor prep_obs
would build.
"""
import numpy as np
import torch


def translate_action(args, action, env=None):
    """Continuous branch of utils/util.py:121-130: returns (actions, cp_actions) with cp_actions the
    clamped + rescaled fp32 tensor the env consumes (quirk Q5: [0, 1] maps to [0.5, 1]).  Unlike the
    reference the result stays a device tensor (the reference ends with .cpu().numpy())."""
    if not getattr(args, "continuous", True):
        raise NotImplementedError("the flex_provision env has continuous actions only")
    actions = action.detach().squeeze()
    low, high = args.action_low, args.action_high
    cp = torch.clamp(actions, min=low, max=high)
    cp = 0.5 * (cp + 1.0) * (high - low) + low
    return actions, cp


def prep_obs(state):
    """utils/util.py:135-145: list of per-agent arrays (one transition) or a tensor -> float tensor
    [batch, n_agents, obs_dim].  A tensor that already has that layout (what the batched env's
    get_obs returns) passes through without a copy."""
    if isinstance(state, torch.Tensor):
        if state.dim() == 2:
            state = state.unsqueeze(0)
        if state.dim() != 3:
            raise RuntimeError('The shape of the observation is incorrect.')
        return state.float()
    state = np.array(state)
    if len(state.shape) == 2:
        state = np.stack(state, axis=0)
    elif len(state.shape) == 4:
        state = np.concatenate(state, axis=0)
    else:
        raise RuntimeError('The shape of the observation is incorrect.')
    return torch.tensor(state).float()
