"""COCO keypoint-result conversion.

Mirrors /root/reference/eval.py:93-125 (`append_result`): every human becomes one COCO result
with 17 keypoints in COCO order (`ORDER_COCO`, eval.py:35), coordinates
``x_norm * upsample_keypoints[1] + 0.5`` / ``y_norm * upsample_keypoints[0] + 0.5`` (eval.py:113),
visibility 1 for present parts, 0-triples for absent ones, and the hard-coded result score 1.0
(eval.py:122).  ``coco_results`` does it for a whole batch straight from the library's result
tables (no per-part Python objects); ``append_result`` is the drop-in on ``Human`` objects.
All arithmetic is float64 in the reference's operation order, so both forms are bit-identical.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

ORDER_COCO = [0, 15, 14, 17, 16, 5, 2, 6, 3, 7, 4, 11, 8, 12, 9, 13, 10]  # eval.py:35


def append_result(image_id, humans, upsample_keypoints, outputs) -> None:
    """Drop-in for eval.py:93-125 (works on torch_ekpose_b200.Human or the reference's Human)."""
    for human in humans:
        keypoints = np.zeros((18, 3))
        for i in range(18):
            if i in human.body_parts:
                bp = human.body_parts[i]
                keypoints[i, 0] = bp.x * upsample_keypoints[1] + 0.5
                keypoints[i, 1] = bp.y * upsample_keypoints[0] + 0.5
                keypoints[i, 2] = 1
        outputs.append({"image_id": image_id, "category_id": 1, "keypoints": list(keypoints[ORDER_COCO, :].reshape(51)),
                        "score": 1.})


def coco_keypoints(num_humans: np.ndarray, parts: np.ndarray, hw_full, upsample_keypoints) -> List[np.ndarray]:
    """Vectorised: per image an array [num_humans[i], 51] of COCO keypoints.

    num_humans [n]; parts [n, max_humans, 18] structured (x, y, score, id) as returned by
    ``PostProcessor.human_tables()``; hw_full = (H, W) of the full-resolution map the coordinates
    refer to (BodyPart.x = x / W, paf_to_pose.py:369-372); upsample_keypoints = one (up_h, up_w)
    pair or one per image (eval.py:166).
    """
    H, W = hw_full
    n = len(num_humans)
    ups = np.asarray(upsample_keypoints, np.float64)
    if ups.ndim == 1:
        ups = np.broadcast_to(ups, (n, 2))
    out = []
    for i in range(n):
        p = parts[i, :int(num_humans[i])]
        present = p["id"] >= 0
        kp = np.zeros(p.shape + (3,), np.float64)
        kp[..., 0] = np.where(present, p["x"].astype(np.float64) / W * ups[i, 1] + 0.5, 0.0)
        kp[..., 1] = np.where(present, p["y"].astype(np.float64) / H * ups[i, 0] + 0.5, 0.0)
        kp[..., 2] = present
        out.append(kp[:, ORDER_COCO, :].reshape(len(p), 51))
    return out


def coco_results(image_ids: Sequence, num_humans, parts, hw_full, upsample_keypoints) -> list:
    """The list of result dicts eval.py feeds to COCO.loadRes (eval.py:75-78), for a whole batch."""
    outputs = []
    for image_id, kps in zip(image_ids, coco_keypoints(num_humans, parts, hw_full, upsample_keypoints)):
        for row in kps:
            outputs.append({"image_id": image_id, "category_id": 1, "keypoints": list(row), "score": 1.})
    return outputs
