from . import registration  # noqa: F401
from .registration import registry  # noqa: F401
