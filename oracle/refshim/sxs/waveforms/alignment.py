def align2d(*a, **k):  # pragma: no cover
    raise ImportError("sxs is not available in this image (oracle/refshim placeholder)")
