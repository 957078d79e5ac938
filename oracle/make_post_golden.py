"""Write tests/golden/post_processing.npz from the UNMODIFIED reference `post_processing` (src/utils.py:55-64), imported by
file path (cv2, which that module imports for an unrelated helper, is stubbed when absent).  Authoring container only.
    python oracle/make_post_golden.py"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import post_oracle as PO  # noqa: E402

REF = Path("/root/reference")
try:
    import cv2  # noqa: F401
except ImportError:
    sys.modules["cv2"] = types.ModuleType("cv2")
spec = importlib.util.spec_from_file_location("ref_utils", REF / "src/utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

params = dict(gauss_sigma=3.0, height=0.2, distance=15)          # src/ball_action/constants.py:39-43
out = {}
for n, seed in ((5000, 1), (997, 2), (40, 3)):
    x = PO.synthetic_raw_predictions(n, 2, seed)
    for c in range(2):
        idx, conf = ref.post_processing(list(range(15, 15 + n)), x[:, c], **params)
        assert (idx, conf) == PO.post_processing(list(range(15, 15 + n)), x[:, c], **params)
        out[f"idx_{n}_{seed}_{c}"] = np.asarray(idx, dtype=np.int64)
        out[f"conf_{n}_{seed}_{c}"] = np.asarray(conf, dtype=np.float32)
np.savez_compressed(ROOT / "tests" / "golden" / "post_processing.npz", **out)
print({k: v.shape for k, v in out.items()})
