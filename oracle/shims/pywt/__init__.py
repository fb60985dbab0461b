"""pywt.Wavelet('haar') filter taps only (wave_modules.py:122-124,159-161)."""
import math


class Wavelet:
    def __init__(self, name):
        if name != "haar":
            raise ValueError("oracle shim only knows the haar wavelet")
        s = 1.0 / math.sqrt(2.0)
        self.dec_lo = [s, s]
        self.dec_hi = [-s, s]
        self.rec_lo = [s, s]
        self.rec_hi = [s, -s]
