"""CPU: host-side logic that needs no GPU (synthetic workload, shape checks, sharding arithmetic)."""
import numpy as np
import pytest

from torch_ekpose_b200 import synthetic
from torch_ekpose_b200.common import BodyPart, CocoPairs, CocoPart, Human
from torch_ekpose_b200.sharding import shard_bounds


def test_synthetic_scene_shapes_and_determinism():
    heat, paf = synthetic.make_scene(46, 54, 3, 7)
    assert heat.shape == (46, 54, 19) and paf.shape == (46, 54, 38)
    assert heat.dtype == np.float32 and paf.dtype == np.float32
    h2, p2 = synthetic.make_scene(46, 54, 3, 7)
    assert np.array_equal(heat, h2) and np.array_equal(paf, p2)
    assert heat[:, :, :18].max() > 0.5 and np.all(heat[:, :, 18] >= 0)
    hb, pb = synthetic.make_batch(5, 46, 54, (1, 3), seed=1, layout="nchw")
    assert hb.shape == (5, 19, 46, 54) and pb.shape == (5, 38, 46, 54) and hb.flags.c_contiguous
    hb2, _ = synthetic.make_batch(5, 46, 54, (1, 3), seed=1, layout="nhwc")
    assert np.array_equal(hb2.transpose(0, 3, 1, 2), hb)
    assert synthetic.SHAPES["368x432"] == (46, 54)


def test_tables_match_reference_header():
    # lib/pafprocess/pafprocess.h:16-24 and lib/utils/common.py:27-30 agree
    assert list(synthetic.COCOPAIRS) == CocoPairs
    assert len(synthetic.COCOPAIRS_NET) == 19 and all(b == a + 1 for a, b in synthetic.COCOPAIRS_NET)
    assert sorted(c for pair in synthetic.COCOPAIRS_NET for c in pair) == list(range(38))
    assert CocoPart.Background.value == 18


def test_human_bodypart_contract():
    hm = Human([])
    hm.body_parts[1] = BodyPart("0-1", 1, 0.25, 0.5, 0.9)
    assert hm.part_count() == 1 and hm.get_max_score() == 0.9 and hm.score == 0.0
    assert hm.body_parts[1].get_part_name() is CocoPart.Neck
    assert "BodyPart:1-(0.25, 0.50) score=0.90" in repr(hm)


@pytest.mark.parametrize("n,world", [(0, 1), (1, 1), (64, 1), (64, 2), (64, 8), (65, 8), (7, 8), (256, 3)])
def test_shard_bounds_partition(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_dims_validation_without_gpu():
    from torch_ekpose_b200.paf_to_pose import PostProcessor
    pp = PostProcessor.__new__(PostProcessor)   # no context: only the shape logic is exercised
    assert pp._dims((2, 19, 46, 54), (2, 38, 46, 54), "nchw") == (2, 46, 54)
    assert pp._dims((2, 46, 54, 19), (2, 46, 54, 38), "nhwc") == (2, 46, 54)
    for bad in (((2, 18, 46, 54), (2, 38, 46, 54)), ((2, 19, 46, 54), (2, 38, 46, 55)), ((19, 46, 54), (38, 46, 54))):
        with pytest.raises(ValueError):
            pp._dims(bad[0], bad[1], "nchw")
    with pytest.raises(ValueError):
        pp._dims((2, 19, 46, 54), (2, 38, 46, 54), "chwn")


def test_coco_conversion_vectorised_equals_object_path():
    """eval.py:93-125: the vectorised table path must equal the per-Human path bit for bit."""
    from torch_ekpose_b200 import coco
    rng = np.random.default_rng(0)
    dt = np.dtype([("x", np.int32), ("y", np.int32), ("score", np.float32), ("id", np.int32)])
    n, mh, H, W = 3, 5, 368, 432
    num = np.array([2, 0, 5], np.int32)
    parts = np.zeros((n, mh, 18), dt)
    parts["x"], parts["y"] = rng.integers(0, W, (n, mh, 18)), rng.integers(0, H, (n, mh, 18))
    parts["score"] = rng.random((n, mh, 18))
    parts["id"] = np.where(rng.random((n, mh, 18)) < 0.7, rng.integers(0, 90, (n, mh, 18)), -1)
    ups = [(368 / 0.77, 432 / 0.77), (368.0, 432.0), (368 / 1.3, 432 / 1.3)]
    got = coco.coco_results([11, 12, 13], num, parts, (H, W), ups)
    want = []
    for i in range(n):
        humans = []
        for k in range(num[i]):
            hm = Human([])
            for p in range(18):
                if parts[i, k, p]["id"] >= 0:
                    hm.body_parts[p] = BodyPart("%d-%d" % (k, p), p, float(parts[i, k, p]["x"]) / W,
                                                float(parts[i, k, p]["y"]) / H, float(parts[i, k, p]["score"]))
            humans.append(hm)
        coco.append_result(11 + i, humans, ups[i], want)
    assert len(got) == len(want) == 7
    for a, b in zip(got, want):
        assert a["image_id"] == b["image_id"] and a["category_id"] == 1 and a["score"] == 1.0
        assert len(a["keypoints"]) == 51 and a["keypoints"] == b["keypoints"]
    assert coco.ORDER_COCO == [0, 15, 14, 17, 16, 5, 2, 6, 3, 7, 4, 11, 8, 12, 9, 13, 10]


def test_estimator_geometry_matches_reference_helpers():
    """padding / preprocessing of the batched hand-off keep the reference's geometry (estimator.py:52-68)."""
    cv2 = pytest.importorskip("cv2")
    from torch_ekpose_b200 import estimator
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    pad, scale, shp = estimator.padding(img, 368)
    assert scale == 368 / 640 and pad.shape == (280, 368, 3) and shp == (276, 368, 3)
    assert not pad[276:].any()
    x = estimator.vgg_preprocess(pad)
    assert x.shape == (3, 280, 368) and x.dtype == np.float32
    want_r = (np.float32(pad[10, 20, 2]) / np.float32(255.) - np.float32(0.485)) / np.float32(0.229)
    assert abs(x[0, 10, 20] - want_r) < 1e-6
    import os, sys
    if os.path.isdir("/root/reference"):   # and bit-for-bit against the reference's own functions
        sys.path.insert(0, "/root/reference")
        from lib.datasets import preprocessing as ref_prep
        assert np.array_equal(ref_prep.vgg_preprocess(pad), x)
        assert np.array_equal(ref_prep.rtpose_preprocess(pad), estimator.rtpose_preprocess(pad))
