"""Batched, on-device hand-off from the network to the post-processing.

The reference's ``get_outputs`` (/root/reference/lib/evaluate/estimator.py:71-87) runs ONE image
(batch 1, :80), copies both outputs to the host and hands NumPy HWC views to the post-processing
(:85-86).  ``get_outputs_batched`` keeps that function's geometry -- long side scaled to 368,
zero-padded to a multiple of 8 (``padding`` :52-68), the same normalisation
(lib/datasets/preprocessing.py:16-43) -- but stacks a list of equally sized frames into one
forward pass and returns the network outputs as CUDA tensors in the layout the model emits
(NCHW): no device-to-host copy, no transpose.  ``infer_humans`` feeds them straight to the CUDA
post-processing.  The model itself is the caller's (cuDNN through PyTorch; out of scope here).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .paf_to_pose import PostProcessor


def _factor_closest(num, factor, is_ceil=True):   # estimator.py:45-49
    num = np.ceil(float(num) / factor) if is_ceil else np.floor(float(num) / factor)
    return int(num) * factor


def padding(im, dest_size, factor=8, is_ceil=True):
    """estimator.py:52-68: scale the LONG side to dest_size (cv2 bilinear), zero-pad to a multiple of factor."""
    import cv2
    im_scale = float(dest_size) / np.max(im.shape[0:2])
    im = cv2.resize(im, None, fx=im_scale, fy=im_scale)
    h, w, c = im.shape
    new_h, new_w = _factor_closest(h, factor, is_ceil), _factor_closest(w, factor, is_ceil)
    im_pad = np.zeros([new_h, new_w, c], dtype=im.dtype)
    im_pad[0:h, 0:w, :] = im
    return im_pad, im_scale, im.shape


def vgg_preprocess(image):
    """preprocessing.py:32-43: /255, BGR->RGB, (x - mean) / std, CHW float32."""
    image = image.astype(np.float32) / 255.
    means, stds = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out = image.copy()[:, :, ::-1]
    for i in range(3):
        out[:, :, i] = out[:, :, i] - means[i]
        out[:, :, i] = out[:, :, i] / stds[i]
    return out.transpose((2, 0, 1)).astype(np.float32)


def rtpose_preprocess(image):
    """preprocessing.py:16-21."""
    image = image.astype(np.float32) / 256. - 0.5
    return image.transpose((2, 0, 1)).astype(np.float32)


_prep_ctx = {}


def get_outputs_batched(images: Sequence[np.ndarray], model, preprocess: str, device, gpu_preprocess: bool = False):
    """Batched ``get_outputs``: returns (pafs [n,38,h,w], heatmaps [n,19,h,w], im_scale) with the
    two tensors left on ``device``.  All images must have the same shape (frames of one stream).
    gpu_preprocess=True uploads the raw uint8 frames and runs padding + normalisation as one CUDA
    kernel (row f4, bit-identical to the host path) instead of cv2 + NumPy on the host."""
    import torch
    if len({im.shape for im in images}) != 1:
        raise ValueError("get_outputs_batched needs equally sized images (batch them per resolution)")
    if gpu_preprocess:
        dev = torch.device(device)
        idx = dev.index or 0
        pp = _prep_ctx.get(idx)
        if pp is None:
            pp = _prep_ctx[idx] = PostProcessor(device=idx, max_batch=1, max_h=5, max_w=5, max_peaks=16, max_humans=4)
        frames = torch.from_numpy(np.stack(images)).to(dev, non_blocking=True)
        batch_var, scale = pp.preprocess(frames, mode=preprocess)
    else:
        prep = {"vgg": vgg_preprocess, "rtpose": rtpose_preprocess}[preprocess]
        batch, scale = [], 1.0
        for im in images:
            im_pad, scale, _ = padding(im, 368, factor=8, is_ceil=True)
            batch.append(prep(im_pad))
        batch_var = torch.from_numpy(np.stack(batch)).float().to(device, non_blocking=True)
    with torch.no_grad():
        predicted, _ = model(batch_var)
    return predicted[-2], predicted[-1], scale


_pp_cache = {}


def infer_humans(images: Sequence[np.ndarray], model, preprocess: str, device, frontend: str = "reference",
                 thr: float = 0.15) -> List[list]:
    """images -> per-image list[Human]: one forward pass, post-processing on the same GPU."""
    pafs, heats, _ = get_outputs_batched(images, model, preprocess, device)
    n, _, h, w = heats.shape
    idx = heats.device.index or 0
    key = (idx, n, h, w)
    pp = _pp_cache.get(key)
    if pp is None:
        pp = _pp_cache[key] = PostProcessor(device=idx, max_batch=n, max_h=h, max_w=w, max_peaks=2048, max_humans=128)
    pp.run(heats, pafs, layout="nchw", frontend=frontend, thr=thr)
    return pp.humans()
