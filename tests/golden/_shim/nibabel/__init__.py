"""Stub of nibabel: just enough surface for the reference package to import.

Test infrastructure only (used by tests/golden/generate.py to import the
read-only reference checkout and record golden vectors).  No file I/O works.
"""


class Nifti1Image:  # a real class: used in type unions and `match` patterns
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("nibabel stub")


def load(*args, **kwargs):
    raise NotImplementedError("nibabel stub")


def save(*args, **kwargs):
    raise NotImplementedError("nibabel stub")


from . import orientations, spatialimages  # noqa: E402,F401
