"""
segmentalist_b200 -- B200-native (sm_100a) implementation of segmentalist's hot path:
score every candidate segment embedding against every mixture component, then
segment with dynamic programming.  Host code mirrors the reference's classes;
all compute goes through libsegb200.so (C ABI in include/segb200.h).
"""
from . import _lib  # noqa: F401

__all__ = ["fbgmm", "gaussian_components_fixedvar", "kmeans", "kmeans_components", "utterances",
           "unigram_acoustic_wordseg", "kmeans_acoustic_wordseg", "batch", "synth"]
