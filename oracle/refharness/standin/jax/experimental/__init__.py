from . import multihost_utils  # noqa
