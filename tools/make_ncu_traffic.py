#!/usr/bin/env python
"""profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) from profiles/ncu_<tag>_traffic_by_capture.json:
DRAM bytes of one captured launch of each kernel kind next to its algorithmic bytes.  The encoder of tools/ncu_target.py 4 runs as
two halves of 10 images on two streams, so an encoder launch covers 10 images; the 3D part covers the 4 stacks.
    python tools/make_ncu_traffic.py r02"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import accounting as acc  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
cap = json.loads((ROOT / "profiles" / f"ncu_{tag}_traffic_by_capture.json").read_text())
enc = {l.name: l for l in acc.encoder_launches(736, 1280, 720, tail_mode=0)}
s3d = {l.name: l for l in acc.stack3d_launches(23, 40, 5, tail_mode=0)}
IMG = 10
PICK = [  # kind key, capture key, launch name, images, description
    ("gemm1x1", "ncu_r02_gemm_b41[0]", "b4.1.pw", IMG, "blocks.4.1 conv_pw (gemm_tc_kernel<ACT>, resident A)"),
    ("gemm1x1_pwl", "ncu_r02_gemm_b41[1]", "b4.1.pwl", IMG, "blocks.4.1 conv_pwl (gemm_tc_kernel<RES>, streamed, per-image gated weights)"),
    ("dwconv2d", "ncu_r02_dw_b41[0]", "b4.1.dw", IMG, "blocks.4.1 conv_dw (dwconv_tma_kernel<1,1>)"),
    ("stem", "ncu_r02_stem[0]", "stem", IMG, "stem_tc_kernel (TMA + tcgen05)"),
    ("conv3x3", "ncu_r02_conv_tc[2]", "b1.1.c3+pwl", IMG, "blocks.1.1 conv_exp + conv_pwl fused (conv_tc_kernel<32,128,32>)"),
    ("conv3x3_b00", "ncu_r02_conv_tc[0]", "b0.0.c3", IMG, "blocks.0.0 ConvBnAct (conv_tc_kernel, column taps folded into N)"),
    ("conv3x3_b10", "ncu_r02_conv_tc[1]", "b1.0.c3+pwl", IMG, "blocks.1.0 EdgeResidual stride 2 (conv_tc_kernel<16,64,32,2>)"),
    ("conv3x3_b20", "ncu_r02_conv_tc[3]", "b2.0.c3+pwl", IMG, "blocks.2.0 EdgeResidual stride 2 (conv_tc_kernel<32,128,48,2>)"),
    ("conv3x3_b21", "ncu_r02_conv_tc_ws[0]", "b2.1.c3+pwl", IMG, "blocks.2.1 (conv_tc_ws_kernel, 3x3 weights streamed from L2)"),
    ("se_fc", "ncu_r02_se_b51[0]", None, IMG, "blocks.5.1 SE (8-CTA cluster per image)"),
    ("dwconv3d", "ncu_r02_dw_3d[0]", "c3d.0.dw", 1, "conv3d_encoder.0 conv_dw 3x3x3 (dwconv_tma_kernel<3,1>), 4 stacks"),
    ("head", "ncu_r02_head[0]", None, 1, "gem_kernel, 4 stacks"),
]
out = {}
for key, ck, name, imgs, desc in PICK:
    if ck not in cap:
        continue
    c = cap[ck]
    algo = None
    if name in enc:
        algo = enc[name].bytes * imgs
    elif name in s3d:
        algo = s3d[name].bytes * 4
    e = {"launch": f"{desc}, {imgs} images ({tag} capture {ck})" if imgs > 1 else f"{desc} ({tag} capture {ck})",
         "traffic_bytes": int(c["traffic_bytes"]), "algorithmic_bytes": int(algo) if algo is not None else (0 if key == "se_fc" else None),
         "duration_us_under_ncu": round(c["duration_us"], 1)}
    if algo and c["traffic_bytes"] < 0.9 * algo:
        e["note"] = "below algorithmic: part of the input / output is still in the 126 MB L2 (write-back after the launch is not counted)"
    out[key] = e
(ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
