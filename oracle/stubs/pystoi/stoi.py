"""Import stub. Test infrastructure only."""


def stoi(*a, **k):
    raise RuntimeError("stub")
