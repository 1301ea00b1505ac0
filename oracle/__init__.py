"""oracle -- CPU checkers for the PAF post-processing hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under torch_ekpose_b200/ imports it, and the product path has
no CPU fallback.

Three checkers:

* ``RefPaf``   -- the UNMODIFIED reference C++ (lib/pafprocess/pafprocess.cpp) compiled into
                  oracle/_ref/libpaf_ref.so by oracle/Makefile.  Present wherever the Makefile
                  ran with /root/reference available; the prebuilt .so travels to the GPU box.
* ``PortPaf``  -- oracle/paf_oracle.c, our plain-C restatement of the same code, pinned
                  bit-for-bit against RefPaf and against tests/golden/.
* front-ends   -- oracle/frontend_oracle.c: the reference's Python NMS() restated in C, and the
                  dense (north_star) front-end whose arithmetic is defined there.

``reference_python()`` imports the reference's own lib/utils/paf_to_pose.py with RefPaf injected as
``lib.pafprocess.pafprocess`` -- from /root/reference where that tree exists (this container), else
from oracle/_ref/py, the sourceless bytecode oracle/build_refpy.py compiled from it (a build output
like the .so; it travels to the GPU box).  tests/golden/make_golden.py uses it to generate the
committed fixtures, bench.py's reference arm / cpu_baseline to time the UNMODIFIED reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
PORT_SO = os.path.join(HERE, "libekp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpaf_ref.so")
REFPY_ROOT = os.path.join(HERE, "_ref", "py")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the checkers (gcc only).  Building the checker is not using it."""
    srcs = [os.path.join(HERE, f) for f in ("paf_oracle.c", "frontend_oracle.c", "Makefile")]
    stale = force or not os.path.exists(PORT_SO) or any(os.path.getmtime(s) > os.path.getmtime(PORT_SO) for s in srcs)
    need_ref = os.path.isdir(REF_ROOT) and (force or not os.path.exists(REF_SO) or not have_refpy())
    if stale or need_ref:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_refpy() -> bool:
    """The reference's own Python for the path is importable (source tree or its byte-compiled form)."""
    return os.path.isfile(os.path.join(REFPY_ROOT, "lib", "utils", "paf_to_pose.bin"))


def refpy_root() -> str:
    if os.path.isfile(os.path.join(REF_ROOT, "lib", "utils", "paf_to_pose.py")):
        return REF_ROOT
    if have_refpy():
        return REFPY_ROOT
    raise FileNotFoundError("neither /root/reference nor oracle/_ref/py (make -C oracle refpy) is available")


class _PafBase:
    """Common ctypes surface: process_paf + the seven getters + whole-state copies."""

    prefix = ""

    def __init__(self, so_path: str):
        self.lib = C.CDLL(so_path)
        p = self.prefix
        L = self.lib
        self._process = getattr(L, p + "process_paf")
        self._process.restype = C.c_int
        self._process.argtypes = [C.c_int] * 3 + [_f32p] + [C.c_int] * 3 + [C.c_void_p] + [C.c_int] * 3 + [_f32p]
        for name, res, args in (("get_num_humans", C.c_int, []), ("get_part_cid", C.c_int, [C.c_int, C.c_int]),
                                ("get_score", C.c_float, [C.c_int]), ("get_part_x", C.c_int, [C.c_int]),
                                ("get_part_y", C.c_int, [C.c_int]), ("get_part_score", C.c_float, [C.c_int]),
                                ("subset_rows", C.c_int, []), ("num_peaks", C.c_int, [])):
            f = getattr(L, p + name)
            f.restype = res
            f.argtypes = args
            setattr(self, name, f)
        self._subset_copy = getattr(L, p + "subset_copy")
        self._subset_copy.argtypes = [_f32p]
        self._subset_copy.restype = None
        self._peaks_copy = getattr(L, p + "peaks_copy")
        self._peaks_copy.argtypes = [_i32p, _i32p, _f32p, _i32p]
        self._peaks_copy.restype = None

    def process_paf(self, peaks, heat_mat, paf_mat) -> int:
        """Same call shape as the SWIG module (pafprocess.i:14): three float32 3-D arrays.

        heat_mat may be an array (only its shape is used, pafprocess.cpp:83) or a shape tuple.
        """
        peaks = np.ascontiguousarray(peaks, np.float32)
        paf_mat = np.ascontiguousarray(paf_mat, np.float32)
        assert peaks.ndim == 3 and paf_mat.ndim == 3
        hs = tuple(heat_mat.shape) if hasattr(heat_mat, "shape") else tuple(heat_mat)
        assert len(hs) == 3
        return self._process(peaks.shape[0], peaks.shape[1], peaks.shape[2], peaks, hs[0], hs[1], hs[2], None,
                             paf_mat.shape[0], paf_mat.shape[1], paf_mat.shape[2], paf_mat)

    def subset(self) -> np.ndarray:
        n = self.subset_rows()
        out = np.zeros((max(n, 1), 20), np.float32)
        if n:
            self._subset_copy(out)
        return out[:n]

    def peaks_line(self):
        n = self.num_peaks()
        x = np.zeros(max(n, 1), np.int32); y = np.zeros(max(n, 1), np.int32)
        s = np.zeros(max(n, 1), np.float32); i = np.zeros(max(n, 1), np.int32)
        if n:
            self._peaks_copy(x, y, s, i)
        return x[:n], y[:n], s[:n], i[:n]

    def as_module(self) -> types.ModuleType:
        """A module object with the reference SWIG module's seven names (pafprocess.h:53-59)."""
        m = types.ModuleType("lib.pafprocess.pafprocess")
        m.process_paf = self.process_paf
        for name in ("get_num_humans", "get_part_cid", "get_score", "get_part_x", "get_part_y", "get_part_score"):
            setattr(m, name, getattr(self, name))
        return m


class RefPaf(_PafBase):
    prefix = "ref_"

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle` where /root/reference exists")
        super().__init__(REF_SO)
        self.lib.ref_std_sort.argtypes = [C.c_int, _f32p, _i32p]

    def sort_scores(self, scores):
        """libstdc++ std::sort with the reference's comparator: returns (sorted scores, permutation)."""
        s = np.ascontiguousarray(scores, np.float32).copy()
        t = np.arange(len(s), dtype=np.int32)
        self.lib.ref_std_sort(len(s), s, t)
        return s, t


class PortPaf(_PafBase):
    prefix = "okp_"

    def __init__(self):
        build()
        super().__init__(PORT_SO)
        L = self.lib
        L.okp_num_connections.argtypes = [C.c_int]; L.okp_num_connections.restype = C.c_int
        L.okp_connections_copy.argtypes = [C.c_int, _i32p, _i32p, _f32p, _i32p, _i32p]
        L.okp_num_candidates.argtypes = [C.c_int]; L.okp_num_candidates.restype = C.c_int
        L.okp_candidates_copy.argtypes = [C.c_int, _i32p, _i32p, _f32p]
        L.okp_sort_scores.argtypes = [C.c_int, _f32p, _i32p]
        L.okp_heapsort_hits.restype = C.c_int

    def connections(self, limb: int):
        n = self.lib.okp_num_connections(limb)
        a = [np.zeros(max(n, 1), np.int32) for _ in range(4)]
        s = np.zeros(max(n, 1), np.float32)
        if n:
            self.lib.okp_connections_copy(limb, a[0], a[1], s, a[2], a[3])
        return dict(cid1=a[0][:n], cid2=a[1][:n], score=s[:n], peak_id1=a[2][:n], peak_id2=a[3][:n])

    def candidates(self, limb: int):
        n = self.lib.okp_num_candidates(limb)
        i1 = np.zeros(max(n, 1), np.int32); i2 = np.zeros(max(n, 1), np.int32); s = np.zeros(max(n, 1), np.float32)
        if n:
            self.lib.okp_candidates_copy(limb, i1, i2, s)
        return dict(idx1=i1[:n], idx2=i2[:n], score=s[:n])

    def sort_scores(self, scores):
        s = np.ascontiguousarray(scores, np.float32).copy()
        t = np.arange(len(s), dtype=np.int32)
        self.lib.okp_sort_scores(len(s), s, t)
        return s, t

    def heapsort_hits(self) -> int:
        return self.lib.okp_heapsort_hits()


class Frontend:
    """oracle/frontend_oracle.c through ctypes.  All tensors are HWC float32 (one image)."""

    DENSE_TAPS = 5

    def __init__(self):
        build()
        L = self.lib = C.CDLL(PORT_SO)
        L.okp_ref_nms.restype = C.c_int
        L.okp_ref_nms.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _f32p, C.c_int]
        L.okp_ref_nms_ex.restype = C.c_int
        L.okp_ref_nms_ex.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _f32p, C.c_int]
        L.okp_scipy_gauss3.argtypes = [_f32p, C.c_int, C.c_int]
        L.okp_scipy_gauss3_weights.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
        L.okp_upsample_nearest.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.okp_upsample_bilinear.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.okp_resize_cubic.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.okp_dense_tables.restype = C.c_int
        L.okp_dense_tables.argtypes = [C.c_int, _f32p, _i32p]
        L.okp_dense_smooth.restype = C.c_int
        L.okp_dense_smooth.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.okp_dense_smooth_sequential.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.okp_dense_nms.restype = C.c_int
        L.okp_dense_nms.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float, _f32p, C.c_int]
        _u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        L.okp_preprocess_dims.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_double)]
        L.okp_resize_linear_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _u8p]
        L.okp_preprocess.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]

    @staticmethod
    def _hwc(a):
        a = np.ascontiguousarray(a, np.float32)
        assert a.ndim == 3
        return a

    def ref_nms(self, heat, thr=0.15, up=8, nparts=18, cap=8192, gauss=False) -> np.ndarray:
        """NMS() + joint_list of paf_to_pose_cpp: float32 [N,5] rows (x, y, score, id, part); gauss=True is
        NMS(bool_gaussian_filt=True), paf_to_pose.py:111-112."""
        heat = self._hwc(heat)
        out = np.zeros((cap, 5), np.float32)
        n = self.lib.okp_ref_nms_ex(heat, heat.shape[0], heat.shape[1], heat.shape[2], nparts, thr, up, int(bool(gauss)), out, cap)
        assert n <= cap
        return out[:n].copy()

    def scipy_gauss3(self, patch) -> np.ndarray:
        """scipy.ndimage.gaussian_filter(patch, sigma=3) of a float32 2-D patch (both sides >= 13)."""
        out = np.ascontiguousarray(patch, np.float32).copy()
        assert out.ndim == 2 and min(out.shape) > 12
        self.lib.okp_scipy_gauss3(out, out.shape[0], out.shape[1])
        return out

    def scipy_gauss3_weights(self) -> np.ndarray:
        """the 13 distinct weights, w[-12] .. w[0]"""
        fw = np.zeros(13, np.float64)
        self.lib.okp_scipy_gauss3_weights(fw)
        return fw

    def resize_cubic(self, patch, up=8) -> np.ndarray:
        patch = np.ascontiguousarray(patch, np.float32)
        out = np.zeros((patch.shape[0] * up, patch.shape[1] * up), np.float32)
        self.lib.okp_resize_cubic(patch, patch.shape[0], patch.shape[1], patch.shape[1], 1, up, out)
        return out

    def upsample_nearest(self, lo, up=8) -> np.ndarray:
        lo = self._hwc(lo)
        out = np.zeros((lo.shape[0] * up, lo.shape[1] * up, lo.shape[2]), np.float32)
        self.lib.okp_upsample_nearest(lo, lo.shape[0], lo.shape[1], lo.shape[2], up, out)
        return out

    def upsample_bilinear(self, lo) -> np.ndarray:
        lo = self._hwc(lo)
        out = np.zeros((lo.shape[0] * 8, lo.shape[1] * 8, lo.shape[2]), np.float32)
        self.lib.okp_upsample_bilinear(lo, lo.shape[0], lo.shape[1], lo.shape[2], out)
        return out

    def dense_tables(self, n: int):
        taps = np.zeros((n * 8, self.DENSE_TAPS), np.float32)
        base = np.zeros(n * 8, np.int32)
        rc = self.lib.okp_dense_tables(n, taps, base)
        if rc:
            raise ValueError(f"okp_dense_tables({n}) -> {rc}")
        return taps, base

    def dense_smooth(self, heat, nparts=18, sequential=False) -> np.ndarray:
        heat = self._hwc(heat)
        out = np.zeros((heat.shape[0] * 8, heat.shape[1] * 8, nparts), np.float32)
        if sequential:
            self.lib.okp_dense_smooth_sequential(heat, heat.shape[0], heat.shape[1], heat.shape[2], nparts, out)
        else:
            rc = self.lib.okp_dense_smooth(heat, heat.shape[0], heat.shape[1], heat.shape[2], nparts, out)
            if rc:
                raise ValueError(f"okp_dense_smooth -> {rc}")
        return out

    def dense_nms(self, smooth, thr=0.15, cap=16384) -> np.ndarray:
        smooth = self._hwc(smooth)
        out = np.zeros((cap, 5), np.float32)
        n = self.lib.okp_dense_nms(smooth, smooth.shape[0], smooth.shape[1], smooth.shape[2], thr, out, cap)
        assert n <= cap
        return out[:n].copy()

    def preprocess_dims(self, sh, sw, dest_size=368, factor=8):
        v = [C.c_int() for _ in range(4)]
        sc = C.c_double()
        self.lib.okp_preprocess_dims(sh, sw, dest_size, factor, *[C.byref(x) for x in v], C.byref(sc))
        return tuple(x.value for x in v) + (sc.value,)   # (rh, rw, ph, pw, scale)

    def resize_linear_u8(self, img, scale) -> np.ndarray:
        img = np.ascontiguousarray(img, np.uint8)
        rh, rw, _, _, _ = self.preprocess_dims(img.shape[0], img.shape[1])
        dh, dw = int(np.rint(img.shape[0] * scale)), int(np.rint(img.shape[1] * scale))
        out = np.zeros((dh, dw, img.shape[2]), np.uint8)
        self.lib.okp_resize_linear_u8(img, img.shape[0], img.shape[1], img.shape[2], scale, dh, dw, out)
        return out

    def preprocess(self, bgr, mode="vgg", dest_size=368, factor=8) -> np.ndarray:
        """padding() + vgg_preprocess / rtpose_preprocess of one uint8 BGR image -> float32 [3, ph, pw]."""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        _, _, ph, pw, _ = self.preprocess_dims(bgr.shape[0], bgr.shape[1], dest_size, factor)
        out = np.zeros((3, ph, pw), np.float32)
        self.lib.okp_preprocess(bgr, bgr.shape[0], bgr.shape[1], dest_size, factor, {"vgg": 0, "rtpose": 1}[mode], out)
        return out

    def dense_peaks(self, heat, thr=0.15) -> np.ndarray:
        return self.dense_nms(self.dense_smooth(heat), thr)


def subset_of(paf_impl: _PafBase, peaks_n5: np.ndarray, H: int, W: int, paf_mat: np.ndarray):
    """Run process_paf on one image's peak list; returns (subset[n,20], peaks_line tuple)."""
    if len(peaks_n5) == 0:
        return np.zeros((0, 20), np.float32), (np.zeros(0, np.int32),) * 2 + (np.zeros(0, np.float32), np.zeros(0, np.int32))
    paf_impl.process_paf(peaks_n5[None], (H, W, 19), paf_mat)
    return paf_impl.subset(), paf_impl.peaks_line()


def reference_cfg():
    """lib.config needs yacs (absent), so cfg is a SimpleNamespace with the values of lib/config/default.py:16-25;
    cfg is passed as an argument by every caller (paf_to_pose.py:346)."""
    ns = types.SimpleNamespace
    return ns(MODEL=ns(NUM_KEYPOINTS=18, DOWNSAMPLE=8),
              TEST=ns(THRESH_HEATMAP=0.15, THRESH_PAF=0.05, NUM_INTERMED_PTS_BETWEEN_KEYPOINTS=10))


class _RefPyFinder:
    """Imports ``lib.*`` from the byte-compiled reference under oracle/_ref/py: packages are the directories there
    (the reference has no lib/__init__.py either), modules the ``<name>.bin`` files build_refpy.py wrote."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname != "lib" and not fullname.startswith("lib."):
            return None
        base = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(base + ".bin"):
            loader = importlib.machinery.SourcelessFileLoader(fullname, base + ".bin")
            return importlib.util.spec_from_file_location(fullname, base + ".bin", loader=loader)
        if os.path.isdir(base):
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        return None


def _reference_path(root=None) -> str:
    root = root or refpy_root()
    if os.path.isfile(os.path.join(root, "lib", "utils", "paf_to_pose.bin")):   # the byte-compiled form
        if not any(isinstance(f, _RefPyFinder) and f.root == root for f in sys.meta_path):
            sys.meta_path.insert(0, _RefPyFinder(root))
    elif root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    return root


def reference_python(use_ref: bool = True, root=None):
    """Import the reference's own lib/utils/paf_to_pose.py, unmodified.

    Returns (paf_to_pose_module, cfg_namespace, paf_impl).  The SWIG module the reference imports
    (paf_to_pose.py:7) is replaced by a ctypes view of the compiled reference C++ (RefPaf) -- SWIG is
    absent and adds no arithmetic.  ``root`` forces /root/reference or oracle/_ref/py.
    """
    _reference_path(root)
    impl = RefPaf() if use_ref else PortPaf()
    mod = impl.as_module()
    pkg = sys.modules.get("lib.pafprocess")
    if pkg is None:   # there is no importable lib/pafprocess package (no __init__, SWIG output absent): a stub holds the module
        pkg = types.ModuleType("lib.pafprocess")
        pkg.__path__ = []
        sys.modules["lib.pafprocess"] = pkg
    sys.modules["lib.pafprocess.pafprocess"] = mod
    pkg.pafprocess = mod
    from lib.utils import paf_to_pose  # noqa: E402
    paf_to_pose.pafprocess = mod
    return paf_to_pose, reference_cfg(), impl


def reference_module(name: str, root=None):
    """Any other byte-compiled reference module, unmodified: 'lib.network.vgg2016', 'lib.evaluate.estimator',
    'lib.datasets.preprocessing', 'lib.utils.common'."""
    import importlib
    _reference_path(root)
    return importlib.import_module(name)


def dense_restatement_libs(heat_hwc: np.ndarray, paf_hwc: np.ndarray, impl: _PafBase, thr: float = 0.15):
    """BASELINE.md section 4's "dense restatement": the north_star stages 1-3 with the library primitives the
    reference itself imports (paf_to_pose.py:1-6) -- cv2.resize(INTER_LINEAR) x8 of heat and PAF,
    scipy.ndimage.gaussian_filter(sigma=3) on the 18 part maps, maximum_filter(size=3) + threshold -- followed by
    ``impl.process_paf`` and the getter loop.  The like-for-like CPU line of the dense GPU arm (the GPU computes
    the same operator in float32 polyphase form; peak SETS agree, tests/test_oracle_pinning.py).
    Returns (peaks[N,5], number of humans)."""
    import cv2
    from scipy.ndimage import gaussian_filter, maximum_filter
    heat_up = cv2.resize(heat_hwc, None, fx=8, fy=8, interpolation=cv2.INTER_LINEAR)
    paf_up = cv2.resize(paf_hwc, None, fx=8, fy=8, interpolation=cv2.INTER_LINEAR)
    rows = []
    for k in range(18):
        g = gaussian_filter(heat_up[:, :, k], sigma=3)
        pk = (maximum_filter(g, size=3) == g) & (g > np.float32(thr))
        ys, xs = np.nonzero(pk)
        for y, x in zip(ys, xs):
            rows.append((x, y, g[y, x], len(rows), k))
    peaks = np.asarray(rows, np.float32).reshape(-1, 5)
    n = 0
    if len(peaks):
        impl.process_paf(peaks[None], heat_up, paf_up)
        n = impl.get_num_humans()
        for hid in range(n):
            for part in range(18):
                cid = impl.get_part_cid(hid, part)
                if cid >= 0:
                    impl.get_part_x(cid), impl.get_part_y(cid), impl.get_part_score(cid)
            impl.get_score(hid)
    return peaks, n
