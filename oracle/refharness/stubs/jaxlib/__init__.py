from . import version
