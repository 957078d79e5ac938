#!/usr/bin/env python
"""End-to-end logit error of the 16 full-size random stacks (the inputs of test_parity_sweep_gpu.py) for every MMA issue order
(MDS_NUMERICS_VARIANT 0..7, csrc/common.cuh): all are the same sums in exact arithmetic; the spread of the results is the
build-to-build scatter of the 1e-3 gate.  tools/build_variants.sh builds tools/bin/libmds_v{k}.so (no GPU needed); on the GPU box:
    python tests/parity_variants.py            # reference once, then one subprocess per variant
Lives under tests/ because it uses the oracle as the checker."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/tmp/parity_variants_ref.pt")


def reference():
    from oracle import mds_oracle as O
    cfg = O.ModelConfig()
    sd = O.make_state_dict(cfg, seed=1234)
    u8 = torch.randint(0, 256, (16, 15, 720, 1280), dtype=torch.uint8, generator=torch.Generator().manual_seed(100))
    with torch.no_grad():
        ref = torch.cat([O.forward(sd, O.pad_normalize(u8[i:i + 1], (1280, 736)), cfg) for i in range(16)])
    torch.save({"sd": sd, "u8": u8, "ref": ref}, REF)


def one(lib_path):
    from ball_action_spotting_b200 import _lib
    if lib_path != "default":
        _lib._LIB_PATH = Path(lib_path)
    from ball_action_spotting_b200 import MultiDimStacker
    d = torch.load(REF)
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_3d_blocks=4, expansion_3d_ratio=3)
    net.load_state_dict(d["sd"])
    net.to("cuda:0").eval()
    got = net(d["u8"].to("cuda:0")).cpu()
    ref = d["ref"]
    err = (got - ref).abs() / ref.abs().max()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs()
    print(json.dumps({"lib": lib_path, "max": round(err.max().item(), 7), "rms": round(err.pow(2).mean().sqrt().item(), 7),
                      "prob_max": round(perr.max().item(), 7), "stacks_over_1e-3": int((err.max(1).values > 1e-3).sum())}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
    else:
        reference()
        libs = ["default"] + sorted(str(p) for p in (ROOT / "tools" / "bin").glob("libmds_v*.so"))
        for lib in libs:
            subprocess.run([sys.executable, __file__, lib], check=False)
