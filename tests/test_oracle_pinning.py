"""CPU: pin the oracle against the real thing wherever the real thing is available.

* the compiled reference C++ (oracle/_ref, built from /root/reference; travels as a prebuilt .so)
* OpenCV / SciPy, the libraries the reference's Python side calls (paf_to_pose.py:1-6)
* the reference's own Python (only where /root/reference exists)
"""
import os

import numpy as np
import pytest

import oracle
from tests import util
from tests.util import assert_bits_equal
from torch_ekpose_b200 import synthetic

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libpaf_ref.so not built")


@needs_ref
@pytest.mark.parametrize("h,w,people,seed", [(46, 54, 1, 1), (46, 54, 4, 2), (46, 82, 7, 3), (64, 96, 22, 4), (92, 164, 36, 5)])
def test_port_equals_compiled_reference_on_seeded_scenes(h, w, people, seed):
    heat, paf = synthetic.make_scene(h, w, people, seed)
    fe = util.frontend()
    peaks, paf_up = fe.ref_nms(heat), fe.upsample_nearest(paf)
    a, la = util.oracle_people(peaks, 8 * h, 8 * w, paf_up, impl=util.ref_or_none())
    b, lb = util.oracle_people(peaks, 8 * h, 8 * w, paf_up)
    assert_bits_equal(a, b, "subset")
    for x, y in zip(la, lb):
        assert np.array_equal(x, y)
    # shuffled input order exercises the id != table-index quirk (pafprocess.cpp:208-218)
    perm = np.random.default_rng(seed).permutation(len(peaks))
    a, la = util.oracle_people(peaks[perm], 8 * h, 8 * w, paf_up, impl=util.ref_or_none())
    b, lb = util.oracle_people(peaks[perm], 8 * h, 8 * w, paf_up)
    assert_bits_equal(a, b, "subset (shuffled peaks)")


@needs_ref
def test_sort_port_equals_libstdcxx_incl_heapsort_fallback():
    ref, port = util.ref_or_none(), util.port()
    rng = np.random.default_rng(0)

    def killer(n):  # median-of-3 killer: drives introsort into its heapsort fallback
        k = n // 2
        a = np.zeros(n)
        for i in range(k):
            a[i] = i + 1 if i % 2 == 0 else k + i + 1
            a[k + i] = 2 * (i + 1)
        return (-a).astype(np.float32)

    before = port.heapsort_hits()
    for n in list(range(0, 40)) + [64, 100, 500, 2048]:
        for v in (rng.random(n), rng.integers(0, 3, n), np.zeros(n), np.sort(rng.integers(0, 9, n)), killer(n) if n > 1 else np.zeros(n)):
            v = np.asarray(v, np.float32)
            a, b = ref.sort_scores(v), port.sort_scores(v)
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    assert port.heapsort_hits() > before   # the fallback really ran


def test_bicubic_restatement_equals_opencv_own_code():
    cv2 = pytest.importorskip("cv2")
    fe = util.frontend()
    rng = np.random.default_rng(3)
    was = cv2.ipp.useIPP()
    try:
        for _ in range(40):
            ph, pw = rng.integers(3, 6, 2)
            patch = rng.random((ph, pw)).astype(np.float32)
            mine = fe.resize_cubic(patch)
            cv2.ipp.setUseIPP(False)
            assert_bits_equal(mine, cv2.resize(patch, None, fx=8, fy=8, interpolation=cv2.INTER_CUBIC), "vs cv2 (IPP off)")
            cv2.ipp.setUseIPP(True)
            ipp = cv2.resize(patch, None, fx=8, fy=8, interpolation=cv2.INTER_CUBIC)
            assert np.abs(mine - ipp).max() <= 2.5e-7 and mine.argmax() == ipp.argmax()
    finally:
        cv2.ipp.setUseIPP(was)


def test_dense_definition_close_to_library_primitives():
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    fe = util.frontend()
    heat, paf = synthetic.make_scene(46, 54, 4, 21)
    up = cv2.resize(heat, None, fx=8, fy=8, interpolation=cv2.INTER_LINEAR)
    assert np.abs(up - fe.upsample_bilinear(heat)).max() <= 1e-6
    lib = np.stack([ndi.gaussian_filter(up[:, :, k], sigma=3) for k in range(18)], -1)
    S = fe.dense_smooth(heat)
    assert np.abs(S - lib).max() <= 1e-6
    want = set()
    for k in range(18):
        m = (ndi.maximum_filter(lib[:, :, k], size=3, mode="constant", cval=-np.inf) == lib[:, :, k]) & (lib[:, :, k] > np.float32(0.15))
        want |= {(int(x), int(y), k) for y, x in zip(*np.nonzero(m))}
    got = {(int(r[0]), int(r[1]), int(r[4])) for r in fe.dense_nms(S)}
    assert got == want


@pytest.mark.skipif(not os.path.isdir(oracle.REF_ROOT), reason="/root/reference only exists in the authoring container")
def test_restatements_equal_reference_python_live():
    cv2 = pytest.importorskip("cv2")
    p2p, cfg, ref = oracle.reference_python()
    fe = util.frontend()
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for seed, (h, w, people) in enumerate([(46, 54, 2), (46, 82, 5), (64, 96, 15)]):
            heat, paf = synthetic.make_scene(h, w, people, 900 + seed)
            jl = p2p.NMS(heat, upsampFactor=8, config=cfg)
            want = np.array([tuple(pk) + (jt,) for jt, jp in enumerate(jl) for pk in jp], np.float32).reshape(-1, 5)
            assert_bits_equal(fe.ref_nms(heat), want, "NMS()")
            humans = p2p.paf_to_pose_cpp(heat, paf, cfg)
            sub, _ = util.oracle_people(want, 8 * h, 8 * w, fe.upsample_nearest(paf))
            assert len(humans) == len(sub)
            assert_bits_equal(sub, ref.subset(), "subset vs the reference run")
    finally:
        cv2.ipp.setUseIPP(was)


def test_input_side_restatement_equals_cv2_and_reference_python():
    """Row f4 oracle: padding() + vgg/rtpose_preprocess restated in C == cv2 (8-bit INTER_LINEAR is
    OpenCV's own fixed-point code, no IPP for 8UC3) and == the reference's Python where it exists."""
    cv2 = pytest.importorskip("cv2")
    from torch_ekpose_b200 import estimator
    fe = util.frontend()
    rng = np.random.default_rng(1)
    for (h, w) in [(480, 640), (720, 1280), (300, 500), (368, 432), (101, 77), (640, 480), (50, 50)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rh, rw, ph, pw, scale = fe.preprocess_dims(h, w)
        assert np.array_equal(fe.resize_linear_u8(img, scale), cv2.resize(img, None, fx=scale, fy=scale))
        pad, s2, shp = estimator.padding(img, 368)
        assert s2 == scale and pad.shape == (ph, pw, 3) and shp == (rh, rw, 3)
        assert_bits_equal(fe.preprocess(img, "vgg"), estimator.vgg_preprocess(pad), "vgg")
        assert_bits_equal(fe.preprocess(img, "rtpose"), estimator.rtpose_preprocess(pad), "rtpose")
    if os.path.isdir(oracle.REF_ROOT):
        import sys
        sys.path.insert(0, oracle.REF_ROOT)
        from lib.datasets import preprocessing as ref_prep
        img = rng.integers(0, 256, (333, 555, 3), dtype=np.uint8)
        pad, _, _ = estimator.padding(img, 368)
        assert_bits_equal(fe.preprocess(img, "vgg"), ref_prep.vgg_preprocess(pad), "vs reference vgg_preprocess")
