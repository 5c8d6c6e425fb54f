from . import modeling  # noqa: F401
from .modeling import *  # noqa: F401,F403
