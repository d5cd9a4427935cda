"""Import stub for ``loess`` (TEST INFRASTRUCTURE ONLY; DELFI is out of scope)."""
