"""sha256 (first 16 hex digits) of every kernel source, printed as JSON: run on the GPU box next to an ncu capture so that
profiles/kernels.json records the sources the capture was made from (tools/make_kernels_json.py)."""
import glob, hashlib, json, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print(json.dumps({os.path.basename(f): hashlib.sha256(open(f, "rb").read()).hexdigest()[:16]
                  for f in sorted(glob.glob(os.path.join(root, "torch_ekpose_b200", "csrc", "*.cu*")))}, indent=1))
