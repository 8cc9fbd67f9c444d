from scipy import special, linalg, signal, interpolate  # noqa
