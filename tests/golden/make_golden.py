"""Generate the committed golden fixtures tests/golden/*.npz.

Run HERE (the authoring container), where /root/reference exists:

    python tests/golden/make_golden.py

Every expected value in the fixtures comes from the REFERENCE ITSELF:
  * peaks      <- the reference's own Python NMS() (lib/utils/paf_to_pose.py:60-133) imported
                  from /root/reference, run twice: with OpenCV's own bicubic code
                  (cv2.ipp.setUseIPP(False); the bit-exact target) and with cv2's default IPP
                  dispatch (the tolerance target, CPU-specific 1-ulp differences);
  * subset / peak table / humans <- the UNMODIFIED reference C++ process_paf compiled into
                  oracle/_ref/libpaf_ref.so, driven by the reference's paf_to_pose_cpp
                  (paf_to_pose.py:346-380).
The dense (north_star) front-end has no reference implementation; its expected peaks come from
oracle/frontend_oracle.c (arithmetic defined there, "parity unpinned by the reference") and the
people for those peaks again from the compiled reference process_paf.

The reference cannot travel to the GPU box, the fixtures can.  Inputs are stored in the files,
so nothing depends on NumPy's RNG stream staying stable.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

import oracle  # noqa: E402
from torch_ekpose_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def humans_to_array(humans):
    """list[Human] -> (parts[n,18,4] = present,x,y,score ; scores[n])"""
    parts = np.zeros((len(humans), 18, 4), np.float64)
    scores = np.zeros(len(humans), np.float64)
    for i, hm in enumerate(humans):
        scores[i] = hm.score
        for k, bp in hm.body_parts.items():
            parts[i, k] = (1.0, bp.x, bp.y, bp.score)
    return parts, scores


def scene_fixture(name, heat, paf, p2p, cfg, ref, fe):
    h, w = heat.shape[:2]
    H, W = 8 * h, 8 * w
    d = dict(heat=heat, paf=paf)

    def ref_peaks():
        jl = p2p.NMS(heat, upsampFactor=cfg.MODEL.DOWNSAMPLE, config=cfg)
        rows = [tuple(pk) + (jt,) for jt, jp in enumerate(jl) for pk in jp]
        return np.array(rows, np.float32).reshape(-1, 5)

    cv2.ipp.setUseIPP(False)
    d["ref_peaks"] = pk_off = ref_peaks()
    humans_off = p2p.paf_to_pose_cpp(heat, paf, cfg)
    d["ref_subset"] = ref.subset() if len(pk_off) else np.zeros((0, 20), np.float32)
    x, y, s, i = ref.peaks_line() if len(pk_off) else [np.zeros(0)] * 4
    d["ref_line_x"], d["ref_line_y"], d["ref_line_score"], d["ref_line_id"] = (
        np.asarray(x, np.int32), np.asarray(y, np.int32), np.asarray(s, np.float32), np.asarray(i, np.int32))
    d["ref_humans_parts"], d["ref_humans_score"] = humans_to_array(humans_off)

    cv2.ipp.setUseIPP(True)
    d["ref_peaks_ipp"] = pk_on = ref_peaks()
    humans_on = p2p.paf_to_pose_cpp(heat, paf, cfg)
    d["ref_subset_ipp"] = ref.subset() if len(pk_on) else np.zeros((0, 20), np.float32)
    d["ref_humans_parts_ipp"], d["ref_humans_score_ipp"] = humans_to_array(humans_on)

    # dense front-end: oracle-defined peaks, reference-computed people
    S = fe.dense_smooth(heat)
    d["dense_peaks"] = dpk = fe.dense_nms(S, np.float32(cfg.TEST.THRESH_HEATMAP))
    paf_mat = fe.upsample_bilinear(paf)
    sub, line = oracle.subset_of(ref, dpk, H, W, paf_mat)
    d["dense_subset"] = sub
    # a few probes of the smoothed / upsampled tensors (full tensors would be tens of MB)
    rng = np.random.default_rng(12345)
    ys = rng.integers(0, H, 64); xs = rng.integers(0, W, 64)
    d["probe_yx"] = np.stack([ys, xs], 1).astype(np.int32)
    d["probe_smooth"] = S[ys, xs, :]
    d["probe_paf_mat"] = paf_mat[ys, xs, :]
    d["probe_heat_mat"] = fe.upsample_bilinear(heat)[ys, xs, :]

    same_xy = len(pk_off) == len(pk_on) and np.array_equal(pk_off[:, [0, 1, 3, 4]], pk_on[:, [0, 1, 3, 4]])
    ties = 0
    port = oracle.PortPaf()
    if len(pk_off):
        oracle.subset_of(port, pk_off, H, W, fe.upsample_nearest(paf))
        for limb in range(19):
            c = port.candidates(limb)["score"]
            if len(c) > 16 and len(np.unique(c)) < len(c):
                ties += 1
    print(f"{name}: {h}x{w} peaks={len(pk_off)} humans={len(humans_off)} dense_peaks={len(dpk)} "
          f"dense_humans={len(sub)} ipp_coords_same={same_xy} limbs_with_ties_over16={ties} "
          f"humans_same_ipp={np.array_equal(d['ref_subset'][:, :18], d['ref_subset_ipp'][:, :18])}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def kat_fixture(ref):
    """SURVEY.md A.6 hand-checkable cases, answers from the compiled reference."""
    d = {}
    paf = np.zeros((64, 64, 38), np.float32)
    paf[:, :, 12] = paf[:, :, 14] = paf[:, :, 16] = 1
    chain = np.array([(10, 10, .9, 0, 1), (20, 10, .8, 1, 2), (30, 10, .7, 2, 3), (40, 10, .6, 3, 4)], np.float32)
    cases = {
        "chain": chain,
        "chain_reversed": chain[::-1].copy(),           # A.1 quirk: ids follow input order
        "short": chain[:3].copy(),                      # 3 parts -> pruned
        "long_limb": np.array([(2, 2, .9, 0, 1), (62, 2, .8, 1, 2)], np.float32),  # norm > h1/2 -> penalty
        "same_pixel": np.array([(10, 10, .9, 0, 1), (10, 10, .8, 1, 2)], np.float32),  # norm == 0 skipped
    }
    d["paf"] = paf
    for k, pk in cases.items():
        sub, line = oracle.subset_of(ref, pk, 64, 64, paf)
        d[k + "_peaks"] = pk
        d[k + "_subset"] = sub
        d[k + "_line_x"] = line[0]
        d[k + "_line_id"] = line[3]
        print("kat", k, "humans", len(sub), sub[:, 18:] if len(sub) else "")
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **d)


def reference_append_result():
    """The reference's OWN append_result (eval.py:93-125) and ORDER_COCO (eval.py:35), taken from its source with `ast`
    (importing eval.py would pull in pycocotools, absent here) and executed unmodified."""
    import ast
    src = open(os.path.join(oracle.REF_ROOT, "eval.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body
            if (isinstance(n, ast.FunctionDef) and n.name == "append_result")
            or (isinstance(n, ast.Assign) and any(getattr(t, "id", None) == "ORDER_COCO" for t in n.targets))]
    assert len(keep) == 2, "eval.py no longer has append_result / ORDER_COCO at module level"
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "eval.py (append_result, ORDER_COCO)", "exec"), ns)
    return ns["append_result"], ns["ORDER_COCO"]


def coco_fixture():
    """Row f3: what the reference's append_result produces from the reference's paf_to_pose_cpp humans on the golden
    scenes (OpenCV's own bicubic code; the coordinates do not depend on IPP).  The inputs are the committed scene
    fixtures, so this can be regenerated without touching them."""
    append_result, order = reference_append_result()
    p2p, cfg, ref = oracle.reference_python()
    cv2.ipp.setUseIPP(False)
    d = {"ORDER_COCO": np.asarray(order, np.int32)}
    ups_by_scene = {"c1_46x54_p3": (368 / 0.77, 432 / 0.77), "c2_46x54_p6": (368.0, 432.0), "c3_46x82_p8": (480.0, 853.0),
                    "c4_crowd_64x96_p24": (512 / 1.3, 768 / 1.3), "empty_46x54": (368.0, 432.0)}
    for image_id, (name, ups) in enumerate(ups_by_scene.items()):
        with np.load(os.path.join(OUT, name + ".npz")) as z:
            heat, paf = z["heat"], z["paf"]
        humans = p2p.paf_to_pose_cpp(heat, paf, cfg)
        outputs = []
        append_result(100 + image_id, humans, ups, outputs)
        d[name + "_upsample_keypoints"] = np.asarray(ups, np.float64)
        d[name + "_image_id"] = np.asarray(100 + image_id)
        d[name + "_keypoints"] = np.asarray([o["keypoints"] for o in outputs], np.float64).reshape(len(outputs), 51)
        d[name + "_score"] = np.asarray([o["score"] for o in outputs], np.float64)
        d[name + "_category_id"] = np.asarray([o["category_id"] for o in outputs], np.int64)
        print("coco", name, "results", len(outputs))
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(os.path.join(OUT, "coco_append_result.npz"), **d)


def border_heat():
    """18 maps of 46x54 with blobs in the corners, on the edges and one cell away from them (patches of 3x3, 3x5, 4x5 ...
    cells: the clipped windows of paf_to_pose.py:100-102), plus interior ones."""
    rng = np.random.default_rng(77)
    h, w = 46, 54
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    heat = np.zeros((h, w, 19), np.float32)
    spots = [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (0, 20), (h - 1, 31), (17, 0), (29, w - 1), (1, 1), (h - 2, w - 2),
             (1, 40), (22, 1), (10, 10), (30, 25), (20, 44)]
    for k in range(18):
        m = np.zeros((h, w))
        for (cy, cx) in spots:
            if rng.random() < 0.6:
                jy, jx = rng.normal(0, 0.35, 2)
                m = np.maximum(m, rng.uniform(0.5, 1.0) * np.exp(-((yy - cy - jy) ** 2 + (xx - cx - jx) ** 2) / (2 * 0.875 ** 2)))
        heat[:, :, k] = (m + rng.normal(0, 0.005, (h, w))).astype(np.float32)
    return heat


def nms_gauss_fixture():
    """NMS(bool_gaussian_filt=True) (paf_to_pose.py:111-112: scipy.ndimage.gaussian_filter(sigma=3) on every upsampled
    patch before the arg-max) by the reference's own Python on the committed scenes and on a border-stress map, OpenCV's
    own bicubic code."""
    p2p, cfg, ref = oracle.reference_python()
    cv2.ipp.setUseIPP(False)
    d = {"border_heat": border_heat()}
    heats = {"border": d["border_heat"]}
    for name in ("c1_46x54_p3", "c2_46x54_p6", "c3_46x82_p8", "c4_crowd_64x96_p24", "empty_46x54"):
        with np.load(os.path.join(OUT, name + ".npz")) as z:
            heats[name] = z["heat"]
    for name, heat in heats.items():
        jl = p2p.NMS(heat, upsampFactor=cfg.MODEL.DOWNSAMPLE, bool_gaussian_filt=True, config=cfg)
        rows = [tuple(pk) + (jt,) for jt, jp in enumerate(jl) for pk in jp]
        d[name + "_peaks"] = np.array(rows, np.float64).reshape(-1, 5)
        jl0 = p2p.NMS(heat, upsampFactor=cfg.MODEL.DOWNSAMPLE, config=cfg)
        plain = sum(len(jp) for jp in jl0)
        if name == "border":   # the clipped windows through the plain refinement as well (the scenes' own fixtures hold theirs)
            d["border_peaks_plain"] = np.array([tuple(pk) + (jt,) for jt, jp in enumerate(jl0) for pk in jp], np.float64).reshape(-1, 5)
        print("nms_gauss", name, "peaks", len(rows), "(without the filter:", plain, ")")
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(os.path.join(OUT, "nms_gauss.npz"), **d)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "coco":   # only the f3 fixture (reads the committed scenes)
        oracle.build()
        coco_fixture()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gauss":  # only the NMS(bool_gaussian_filt=True) fixture (reads the committed scenes)
        oracle.build()
        nms_gauss_fixture()
        return
    oracle.build(force=True)
    p2p, cfg, ref = oracle.reference_python()
    fe = oracle.Frontend()
    scenes = {
        "c1_46x54_p3": synthetic.make_scene(46, 54, 3, 1003),
        "c2_46x54_p6": synthetic.make_scene(46, 54, 6, 2006),
        "c3_46x82_p8": synthetic.make_scene(46, 82, 8, 3008),
        "c4_crowd_64x96_p24": synthetic.make_scene(64, 96, 24, 4024),
    }
    empty_heat, empty_paf = synthetic.make_scene(46, 54, 0, 5000, noise=False)
    scenes["empty_46x54"] = (empty_heat, empty_paf)
    for name, (heat, paf) in scenes.items():
        scene_fixture(name, heat, paf, p2p, cfg, ref, fe)
    kat_fixture(ref)
    coco_fixture()
    nms_gauss_fixture()


if __name__ == "__main__":
    main()
