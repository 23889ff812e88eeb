"""flexgpu -- B200-native batched flex_provision environment (Safe-MARL hot path).

Public surface:
    BatchedFlexProvisionEnv   N envs on one GPU, device tensors in/out (the fast path)
    FlexibilityProvisionEnv   N=1 drop-in with the reference's exact Python types
    create_network, Network   feeder topology (utils/create_net.py)
    Profiles, synthetic_profiles, load_csv_profiles
"""
from .config import DEFAULT_ENV_ARGS, load_env_yaml, normalize_args
from .network import Network, create_network
from .profiles import Profiles, load_csv_profiles, synthetic_profiles
from ._lib import FlexGpuError, INFO_KEYS, STAT_KEYS
from .util import prep_obs, translate_action
from .env import ObsRing
from . import sharding


def __getattr__(name):
    # torch is imported lazily so that `import flexgpu` stays cheap for host-only users
    if name == "BatchedFlexProvisionEnv":
        from .env import BatchedFlexProvisionEnv
        return BatchedFlexProvisionEnv
    if name == "FlexibilityProvisionEnv":
        from .compat import FlexibilityProvisionEnv
        return FlexibilityProvisionEnv
    if name in ("VoltagePredictor", "DeviceReplayBuffer"):
        from . import predictor
        return getattr(predictor, name)
    raise AttributeError(name)
