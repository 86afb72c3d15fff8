"""Drop-in for the reference's ``freqencoder`` package (freqencoder/freq.py)."""
from .freq import FreqEncoder, freq_encode  # noqa: F401
