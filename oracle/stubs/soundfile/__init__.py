"""Import stub (wesep/cli/extractor.py imports soundfile). Test infrastructure only."""


def read(*a, **k):
    raise RuntimeError("stub")


def write(*a, **k):
    raise RuntimeError("stub")
