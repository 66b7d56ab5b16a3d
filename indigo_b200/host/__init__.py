"""Host-side mirror of the reference interface for the hot path (see hostbackend.py)."""
from .hostbackend import HostBackend, DeviceArrayBase          # noqa: F401
from . import optree as operators                              # noqa: F401
from . import rewrites as transforms                           # noqa: F401
from . import treeinfo as analyses                             # noqa: F401
