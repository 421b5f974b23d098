"""B200-native path-tracing backend: Python bindings over the C ABI (include/rt_b200.h, rt_host.h)."""
from . import abi  # noqa: F401
from .abi import RtError  # noqa: F401
