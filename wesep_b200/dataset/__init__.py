"""Device-side data front end: online mixing and enrollment fbank on the GPU."""
from .processor import OnlineMixer, compute_fbank, snr_mixer  # noqa: F401
