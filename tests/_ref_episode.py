"""Shared by the CPU (oracle) and GPU (CUDA path) parity tests: replay of an episode recorded from the
reference's own FlexibilityProvisionEnv (tests/golden/ref_env_*.npz, made by tests/golden/make_ref_golden.py)."""
import numpy as np

GUARD = 1e-8        # a reference voltage closer than this to a limit is left out of the bit-exact mask comparison


def ref_masks(V, v_min=0.9, v_max=1.1):
    """(mask, decided): violation mask derived from the reference's voltages; `decided` clears the bits of
    buses whose voltage is within GUARD of a limit (the sweep and the reference's root finder agree to 1e-8)."""
    viol = (V > v_max) | (V < v_min)
    near = np.minimum(np.abs(V - v_min), np.abs(V - v_max)) < GUARD
    bits = (1 << np.arange(V.shape[-1], dtype=np.uint64))
    return (viol * bits).sum(-1).astype(np.uint64), (~near * bits).sum(-1).astype(np.uint64)


def replay_reference_episode(g, reset, step, state, tol_v=1e-8):
    """Drives `step` (an adapter around the oracle mirror or the CUDA env) through a stored reference
    episode and compares every output; returns the number of violation bits that were checked."""
    reset(g['e0'][None], g['a0'][None])
    s = state()
    assert np.max(np.abs(s['V'] - g['V0'])) < tol_v and np.max(np.abs(s['E'] - g['E0'])) < 1e-15
    acts = g['actions'].astype(np.float32) if g['actions_f32'][0] else g['actions']
    checked = 0
    for t in range(len(g['reward'])):
        r, d, info = step(acts[t][None])
        s = state()
        rt = g['reward'][t]
        assert abs(r - rt) <= 1e-6 * abs(rt) + 1e-12, (t, r, rt)                            # north_star: 1e-6 relative
        assert bool(d) == bool(g['done'][t]), t
        assert np.max(np.abs(info[:7] - g['info'][t]) / np.maximum(1.0, np.abs(g['info'][t]))) < 1e-8, t
        assert (info[7] == 1.0) == bool(g['failed'][t]), t                                  # info['solver_failed'] (:337)
        assert np.max(np.abs(s['V'] - g['V'][t])) < tol_v, t
        assert np.max(np.abs(s['E'] - g['E'][t])) < 1e-15 and np.max(np.abs(s['setp'] - g['setp'][t])) < 1e-15, t
        assert np.max(np.abs(s['state'] - g['state'][t])) < tol_v, t
        want, decided = ref_masks(g['V'][t])
        assert (np.uint64(s['vmask']) & decided) == (want & decided), t                     # masks bit-exact
        if decided == np.uint64((1 << 33) - 1):
            assert s['vcount'] == bin(int(want)).count("1"), t
        checked += bin(int(want & decided)).count("1")
    return checked


