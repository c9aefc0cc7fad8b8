"""Import stub (wesep/dataset/processor.py imports librosa for resample). Test infrastructure only."""
