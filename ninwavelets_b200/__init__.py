"""ninwavelets_b200 - the reference's WaveletBase API on a hand-written sm_100a CUDA backend.

Export list mirrors the reference's `ninwavelets/__init__.py:1-3`.
"""
from .base import WaveletBase, WaveletMode, plot_tf, Baseline
from .wavelets import Morse, MorseMNE, Morlet, Haar, MexicanHat, Shannon
from .mneutils import EpochsWavelet

__all__ = ["WaveletBase", "WaveletMode", "plot_tf", "Baseline", "Morse", "MorseMNE", "Morlet", "Haar",
           "MexicanHat", "Shannon", "EpochsWavelet"]
