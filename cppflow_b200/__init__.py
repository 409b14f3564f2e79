"""cppflow_b200: B200-native (sm_100a) path-refinement hot path of jstmn/cppflow behind the reference's
search.py / collision_detection.py / optimization.py / optimization_utils.py call surface."""
from . import config  # noqa: F401
from . import _lib  # noqa: F401

__version__ = "0.1.0"
