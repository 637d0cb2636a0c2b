"""Runs the camera_converter CLI on the sample Kannala-Brandt camera (tests/golden/cameras.json):  python scripts/run_converter.py [num_points]
(under torchrun the work is sharded over the ranks)."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_camera_models_b200 import camera_converter as cc
p = json.load(open(os.path.join(ROOT, "tests", "golden", "cameras.json")))["kannala_brandt"]["params"]
d = tempfile.mkdtemp()
y = os.path.join(d, f"kb_{os.environ.get('RANK', '0')}.yaml")
open(y, "w").write(f"cam0:\n  camera_model: kannala_brandt\n  intrinsics: [{p[0]!r}, {p[1]!r}, {p[2]!r}, {p[3]!r}]\n"
                   f"  distortion: [{p[4]!r}, {p[5]!r}, {p[6]!r}, {p[7]!r}]\n  resolution: [512, 512]\n")
n = sys.argv[1] if len(sys.argv) > 1 else "500"
sys.exit(cc.main(["-i", "kb", "-p", y, "-n", n, "-o", os.path.join(d, "out")]))
