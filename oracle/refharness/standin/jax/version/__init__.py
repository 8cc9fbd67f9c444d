__version__ = "0.0.0+numpy-standin"
