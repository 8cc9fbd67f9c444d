from . import linen, core
