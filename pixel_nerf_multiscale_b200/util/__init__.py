from .util import *  # noqa: F401,F403
from . import args, conf  # noqa: F401
