"""Minimal stand-in for flax (see ../README.md): `flax.linen` modules with Flax's
auto-naming, `flax.struct.dataclass`."""
from . import linen  # noqa: F401
from . import struct  # noqa: F401
