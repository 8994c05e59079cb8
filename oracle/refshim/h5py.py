"""Placeholder so `import scri` succeeds (scri/SpEC/file_io imports h5py at module level); file I/O is out of scope."""


class File:  # pragma: no cover
    def __init__(self, *a, **k):
        raise ImportError("h5py is not available in this image (oracle/refshim placeholder)")
