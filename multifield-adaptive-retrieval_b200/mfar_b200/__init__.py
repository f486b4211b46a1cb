"""mfar_b200 - B200-native multi-field scoring + top-k retrieval for mFAR.

Host-side mirror of the reference's ``mfar.data.*`` / ``mfar.modeling.*`` scorer API; every
numeric operation runs in ``libmfar_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/mfar_b200.h``).  There is no CPU fallback: without the library or without an sm_100
device the compute entry points raise.
"""
from .data.typedef import Field, FieldType            # noqa: F401
from .data.schema import resolve_fields               # noqa: F401
from .data.util import MemoryMapDict                  # noqa: F401

__all__ = ["Field", "FieldType", "resolve_fields", "MemoryMapDict"]
