"""Stand-in for gymnasium (see ../README.md): just enough surface to import the reference."""
from . import core, spaces, envs  # noqa: F401
from .core import Env  # noqa: F401
