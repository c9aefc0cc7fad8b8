"""Import stub (wesep/utils/score.py imports pesq). Test infrastructure only."""


def pesq(*a, **k):
    raise RuntimeError("stub")
