"""koala_b200 -- B200-native drop-in for Picovoice Koala's per-frame noise-suppression path.

Same public names as the reference Python package (/root/reference/binding/python/__init__.py:12-14): `create`,
`available_devices`, `Koala`, the Koala*Error classes, `default_library_path`, `default_model_path`; plus `BatchKoala`,
the batched throughput entry point.  Everything computes in libpv_koala_b200.so (hand-written sm_100a CUDA).
"""
from ._batch import *      # noqa: F401,F403
from ._factory import *    # noqa: F401,F403
from ._koala import *      # noqa: F401,F403
from ._util import *       # noqa: F401,F403
from .sharding import *    # noqa: F401,F403
from .files import *       # noqa: F401,F403

__version__ = "1.0.0"
