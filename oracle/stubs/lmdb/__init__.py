"""Import stub (wesep/dataset/lmdb_data.py imports lmdb). Test infrastructure only."""
