"""The configuration values the hot path reads.

/root/reference/lib/config/default.py:16-25 defines a yacs node; the post-processing path reads
only MODEL.NUM_KEYPOINTS, MODEL.DOWNSAMPLE and TEST.THRESH_HEATMAP (paf_to_pose.py:94, 96, 348,
357).  TEST.THRESH_PAF and TEST.NUM_INTERMED_PTS_BETWEEN_KEYPOINTS exist in the reference config
but its C++ ignores them (compile-time constants, pafprocess.h:7, 13); they are kept for
attribute compatibility.  Any object with the same attributes (e.g. the reference's own ``cfg``)
can be passed instead.
"""
from types import SimpleNamespace

cfg = SimpleNamespace(
    MODEL=SimpleNamespace(NUM_KEYPOINTS=18, DOWNSAMPLE=8),
    TEST=SimpleNamespace(THRESH_HEATMAP=0.15, THRESH_PAF=0.05, NUM_INTERMED_PTS_BETWEEN_KEYPOINTS=10),
)
