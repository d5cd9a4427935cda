"""Import stub (TEST INFRASTRUCTURE ONLY): reference frag/_delfi_gc_correct.py:11."""


def loess_1d(*a, **k):  # pragma: no cover
    raise NotImplementedError("loess is not available in this image")
