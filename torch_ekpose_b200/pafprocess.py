"""Drop-in for the reference's SWIG module ``lib.pafprocess.pafprocess``.

Same seven names, argument order and last-call-global semantics as
/root/reference/lib/pafprocess/pafprocess.h:53-59 (bound by pafprocess.i:14-15); the work runs
in libekpose_b200.so on the GPU.  ``sys.modules['lib.pafprocess.pafprocess'] = this module`` (or
``from torch_ekpose_b200 import pafprocess``) is all the reference's paf_to_pose.py needs.

Argument conversion follows the numpy.i typemap the reference uses (numpy.i:316-338, 1097-1127):
any array-like is converted to a C-contiguous float32 array (copying if needed) and must be 3-D.
Differences from the reference, all on paths where it has undefined behaviour: invalid input
raises ValueError and a missing GPU raises RuntimeError instead of returning garbage.
"""
from __future__ import annotations

import numpy as np

from . import _lib

# constants the SWIG module also exports through %include "pafprocess.h" (pafprocess.h:6-24)
THRESH_HEAT = 0.05
THRESH_VECTOR_SCORE = 0.05
THRESH_VECTOR_CNT1 = 6
THRESH_PART_CNT = 4
THRESH_HUMAN_SCORE = 0.3
NUM_PART = 18
STEP_PAF = 10
COCOPAIRS_SIZE = 19


def _as_f32_3d(a, name):
    try:
        arr = np.ascontiguousarray(a, dtype=np.float32)
    except (TypeError, ValueError) as e:
        raise TypeError(f"process_paf: {name} is not convertible to a float32 array") from e
    if arr.ndim != 3:
        raise TypeError(f"process_paf: {name} must have 3 dimensions, got {arr.ndim}")  # numpy.i require_dimensions
    return arr


def process_paf(peaks, heat_mat, paf_mat) -> int:
    """pafprocess.cpp:22-194.  heat_mat is used only through its shape[0] (:83)."""
    peaks = _as_f32_3d(peaks, "peaks")
    paf_mat = _as_f32_3d(paf_mat, "paf_mat")
    if isinstance(heat_mat, np.ndarray):
        if heat_mat.ndim != 3:
            raise TypeError(f"process_paf: heat_mat must have 3 dimensions, got {heat_mat.ndim}")
        hs = heat_mat.shape
    else:
        hs = _as_f32_3d(heat_mat, "heat_mat").shape
    rc = _lib.lib.process_paf(peaks.shape[0], peaks.shape[1], peaks.shape[2], peaks.ctypes.data, hs[0], hs[1], hs[2],
                              None, paf_mat.shape[0], paf_mat.shape[1], paf_mat.shape[2], paf_mat.ctypes.data)
    if rc == _lib.ERR_ARG:
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    return 0


def get_num_humans() -> int:
    return _lib.lib.get_num_humans()


def get_part_cid(human_id: int, part_id: int) -> int:
    return _lib.lib.get_part_cid(int(human_id), int(part_id))


def get_score(human_id: int) -> float:
    return _lib.lib.get_score(int(human_id))


def get_part_x(cid: int) -> int:
    return _lib.lib.get_part_x(int(cid))


def get_part_y(cid: int) -> int:
    return _lib.lib.get_part_y(int(cid))


def get_part_score(cid: int) -> float:
    return _lib.lib.get_part_score(int(cid))
