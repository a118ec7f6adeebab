from .util import *  # noqa: F401,F403
from . import conf  # noqa: F401
