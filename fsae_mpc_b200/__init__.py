"""fsae_mpc_b200 -- B200-native batched LTV-MPC step (drop-in for kerry-he/fsae-mpc's
mpc/ltv path).  CUDA-only: importing the API needs the built in-tree library."""
from .api import FsaeMpc, FsaePool, FsaeError, MpcResult, default_params, KINEMATIC, DYNAMIC  # noqa: F401
from ._lib import Params, LIN_EULER, LIN_RK2, LIN_RK4  # noqa: F401
