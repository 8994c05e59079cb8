"""Import the UNMODIFIED reference (`/root/reference/scri`) in this container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python + numba; what keeps it from importing here
is only its third-party dependencies (quaternion, spherical_functions, spinsfast, h5py), none of which is installable.
`load()` puts stand-ins for those (oracle/refshim/, built on the oracle's restatements of their published arithmetic)
on sys.path and imports the real `scri` package from /root/reference, so that the reference's OWN code - its transform
flow, its numba loops, its frame logic, its codec - runs and can (1) validate oracle/scri_ref.py and (2) generate the
golden vectors under tests/golden/ (tests/golden/make_reference_vectors.py).  /root/reference does not exist on the
GPU box: nothing that runs there calls this module.
"""
import importlib
import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scri"))


def load():
    """Returns the reference's `scri` module (imported from /root/reference, third-party packages shimmed)."""
    if "scri" in sys.modules and getattr(sys.modules["scri"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["scri"]
    if not available():
        raise ImportError(f"{REFERENCE_ROOT}/scri is not present (the reference only exists in the build container)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_scri_reference")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, _SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib.metadata as md

    real_version = md.version

    def version(name):
        if name == "scri":
            try:
                return real_version(name)
            except md.PackageNotFoundError:
                return "2024.0.13"
        return real_version(name)

    md.version = version
    try:
        scri = importlib.import_module("scri")
    finally:
        md.version = real_version
    return scri
