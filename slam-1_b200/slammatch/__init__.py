"""slammatch -- B200-native exact Hamming kNN-2 matching of 256-bit ORB descriptors.

Drop-in for the matcher object of the DavidHan008/SLAM-1 pipeline (tracking.py:12-34,
keypoint.py:35-57, Point3D.py:33-52): ``slammatch.Matcher`` mirrors the cv2 ``DescriptorMatcher``
members the reference uses, ``slammatch.install()`` rebinds ``cv2.FlannBasedMatcher`` so the unmodified
reference runs on the GPU, ``slammatch.knn2`` is the array fast path.  All arithmetic lives in
libslammatch.so (hand-written CUDA for sm_100a behind the C-ABI of include/slammatch.h).
"""
from . import synth  # noqa: F401  (pure numpy; safe without a GPU)
from ._lib import SlamMatchError, Context, context, load, LIB_PATH, SYMBOLS, VARIANTS  # noqa: F401
from .matcher import (DMatch, Matcher, REFERENCE_RATIO, find_2d_3d_device, get_matches, get_matches_device,  # noqa: F401
                      good_matches, install, knn2, knn2_batched, stereo_matches_device, uninstall)

from .keyframe_db import KeyframeDB, ShardedKeyframeDB  # noqa: F401,E402

__all__ = ["KeyframeDB", "ShardedKeyframeDB", "Matcher", "DMatch", "knn2", "knn2_batched", "install", "uninstall", "get_matches",
           "get_matches_device", "find_2d_3d_device", "stereo_matches_device", "good_matches", "context",
           "Context", "SlamMatchError", "load", "synth", "REFERENCE_RATIO"]
