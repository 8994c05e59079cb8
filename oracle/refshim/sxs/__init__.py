"""Placeholder so `import scri` succeeds (the frame-fixing drivers import sxs at module level); not on the hot path."""
from . import metadata, utilities, waveforms  # noqa: F401

__version__ = "0+oracle.shim"


class WaveformModes:  # pragma: no cover
    def __init__(self, *a, **k):
        raise ImportError("sxs is not available in this image (oracle/refshim placeholder)")
