"""Module constants of the reference's cppflow/config.py (:8-25), restated for the CUDA host.

The reference takes DEVICE from jrl.config; this package only runs on CUDA, so DEVICE is the current CUDA device
when one exists (import still works on a CPU-only box so that host logic and the C-ABI symbol table can be tested)."""
import torch

DEVICE = "cuda:0" if torch.cuda.is_available() else "cpu"
DEFAULT_TORCH_DTYPE = torch.float32

VERBOSITY = 2

SUCCESS_THRESHOLD_initial_q_norm_dist = 0.2  # config.py:15

DEFAULT_RERUN_MJAC_THRESHOLD_DEG = 13.0  # config.py:17
DEFAULT_RERUN_MJAC_THRESHOLD_CM = 3.42  # config.py:18
OPTIMIZATION_CONVERGENCE_THRESHOLD = 0.005  # config.py:19

# LM optimization (config.py:23-25)
SELF_COLLISIONS_IGNORED = False
ENV_COLLISIONS_IGNORED = False
DEBUG_MODE_ENABLED = False
