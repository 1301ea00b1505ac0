"""oracle/build_refpy.py -- TEST INFRASTRUCTURE: byte-compile the reference's own Python for the hot path.

The reference's Python side of the path (lib/utils/paf_to_pose.py, lib/utils/common.py) and the two
modules configs[0] / configs[4] need around it (lib/network/vgg2016.py, lib/evaluate/estimator.py,
lib/datasets/preprocessing.py) are compiled -- from where they lie under /root/reference, unmodified --
into SOURCELESS bytecode under oracle/_ref/py/lib/.../<module>.bin (standard .pyc content; the .bin suffix because
snapshot tools commonly drop *.pyc), exactly as oracle/Makefile compiles
lib/pafprocess/pafprocess.cpp into oracle/_ref/libpaf_ref.so.  No reference source is copied into the
repository: oracle/_ref/ holds build outputs only, is git-ignored, and travels to the GPU box with the
snapshot (same image, same interpreter, so the bytecode loads there).  bench.py's reference arm and
cpu_baseline import the result through oracle.reference_python(); nothing under torch_ekpose_b200/ does.

usage: python oracle/build_refpy.py [REF_ROOT] [OUT_DIR]
"""
import os
import py_compile
import sys

FILES = (
    "lib/utils/paf_to_pose.py",        # NMS, paf_to_pose_cpp  (the Python side of the hot path)
    "lib/utils/common.py",             # Human, BodyPart
    "lib/network/vgg2016.py",          # configs[0] / configs[4]: the backbone, run unmodified
    "lib/evaluate/estimator.py",       # padding / get_outputs geometry (row f1 / f4 checks)
    "lib/datasets/preprocessing.py",   # vgg_preprocess / rtpose_preprocess
)


def main(ref_root: str, out_dir: str) -> int:
    if not os.path.isfile(os.path.join(ref_root, FILES[0])):
        print(f"reference tree {ref_root} absent: keeping prebuilt {out_dir} (if any)")
        return 0
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        dst = os.path.join(out_dir, rel[:-3] + ".bin")   # lib/utils/paf_to_pose.bin, imported by oracle._RefPyFinder
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(out_dir, "BUILT_FROM"), "w") as f:
        f.write(f"{ref_root} with python {sys.version.split()[0]}; files: {', '.join(FILES)}\n")
    print(f"built {out_dir} (sourceless bytecode of {len(FILES)} reference modules) from {ref_root}")
    return 0


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference",
                  sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "_ref", "py")))
