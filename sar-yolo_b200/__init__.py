from . import synth  # noqa: F401  (temporary minimal init; replaced below)
