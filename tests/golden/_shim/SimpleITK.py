"""Stub of SimpleITK (I/O only in the reference; never reached on the hot path)."""


class Image:
    pass
