class Metadata(dict):  # pragma: no cover - placeholder, see sxs/__init__.py
    pass
