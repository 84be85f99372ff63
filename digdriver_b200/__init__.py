"""dig-b200: B200-native (sm_100a) implementation of DIGDriver's genome-scan, element-transfer and
burden-test hot path.  See DESIGN.md; the C ABI is include/dig_b200.h."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
