"""Import stub (wesep/utils/score.py imports pystoi.stoi). Test infrastructure only."""
