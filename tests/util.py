"""Shared helpers for the parity tests.  The oracle (oracle/) is the CHECKER here and nothing else."""
import functools
import glob
import os

import numpy as np

import oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = ["c1_46x54_p3", "c2_46x54_p6", "c3_46x82_p8", "c4_crowd_64x96_p24", "empty_46x54"]


@functools.lru_cache(maxsize=None)
def golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=None)
def port():
    return oracle.PortPaf()


@functools.lru_cache(maxsize=None)
def frontend():
    return oracle.Frontend()


@functools.lru_cache(maxsize=None)
def ref_or_none():
    return oracle.RefPaf() if oracle.have_ref() else None


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = bits(a) == bits(b)
    if not same.all():
        idx = np.argwhere(~same)[:5]
        raise AssertionError(f"{what}: {int((~same).sum())} of {same.size} float32 values differ bitwise, first at "
                             f"{idx.tolist()}: {a[tuple(idx[0])]!r} vs {b[tuple(idx[0])]!r}")


def peaks_table(res, i):
    """(x, y, score, id, part) float32 rows of image i from PostProcessor.results(with_peaks=True)."""
    n = int(res["n_peaks"][i])
    line = res["peaks"][i][:n]
    po = res["part_off"][i]
    part = np.zeros(n, np.float32)
    for k in range(18):
        part[po[k]:po[k + 1]] = k
    out = np.zeros((n, 5), np.float32)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3], out[:, 4] = line["x"], line["y"], line["score"], line["id"], part
    return out


def oracle_people(peaks_n5, H, W, paf_mat, impl=None):
    """subset rows for one image: from the COMPILED REFERENCE (oracle/_ref/libpaf_ref.so, lib/pafprocess/pafprocess.cpp
    unmodified) wherever it was built -- it travels to the GPU box -- else from its C restatement (pinned to it bit
    for bit by tests/test_oracle_pinning.py)."""
    impl = impl or ref_or_none() or port()
    sub, line = oracle.subset_of(impl, np.ascontiguousarray(peaks_n5, np.float32), H, W, paf_mat)
    return sub, line


def oracle_dense(heat_hwc, paf_hwc, thr=0.15):
    """Dense front-end + people, all from the oracle: (peaks[N,5], subset[n,20])."""
    fe = frontend()
    peaks = fe.dense_peaks(heat_hwc, np.float32(thr))
    paf_mat = fe.upsample_bilinear(paf_hwc)
    sub, _ = oracle_people(peaks, heat_hwc.shape[0] * 8, heat_hwc.shape[1] * 8, paf_mat)
    return peaks, sub


def oracle_reference(heat_hwc, paf_hwc, thr=0.15):
    """Reference front-end + people, all from the oracle: (peaks[N,5], subset[n,20])."""
    fe = frontend()
    peaks = fe.ref_nms(heat_hwc, np.float32(thr))
    paf_mat = fe.upsample_nearest(paf_hwc)
    sub, _ = oracle_people(peaks, heat_hwc.shape[0] * 8, heat_hwc.shape[1] * 8, paf_mat)
    return peaks, sub
