"""Import stub (wesep/cli/extractor.py imports silero_vad). Test infrastructure only."""


def load_silero_vad(*a, **k):
    raise RuntimeError("stub")


def get_speech_timestamps(*a, **k):
    raise RuntimeError("stub")
