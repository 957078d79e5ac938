"""CPU: host-side logic of the product package (no compute calls): reference API mirror, state-dict contract,
BN folding, C-ABI symbol table."""
import re
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import mds_oracle as O

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib):
    hdr = (ROOT / "include" / "mds_b200.h").read_text()
    declared = set(re.findall(r"\b(mds_[a-z0-9_]+)\s*\(", hdr))
    from ball_action_spotting_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mds_last_error() is not None
    assert lib.mds_launch_count(0) >= 0


def test_state_dict_contract_matches_reference_keys(oracle_sd):
    from ball_action_spotting_b200 import MultiDimStacker
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_frames=15, stack_size=3, num_3d_blocks=4,
                          expansion_3d_ratio=3, se_reduce_3d_ratio=24, drop_rate=0.2, drop_path_rate=0.2)
    mine = net.state_dict()
    assert list(mine.keys()) == list(oracle_sd.keys())          # same keys, same order as the reference module
    for k in mine:
        assert mine[k].shape == oracle_sd[k].shape, k
    net.load_state_dict(oracle_sd, strict=True)
    assert net.num_stacks == 5 and net.num_features == 1280 and net.num_3d_features == 192
    n_enc = sum(p.numel() for p in net.conv2d_encoder.parameters())
    assert n_enc == 5_610_384


def test_constructor_rejects_what_the_reference_rejects():
    from ball_action_spotting_b200 import MultiDimStacker
    with pytest.raises(AssertionError):
        MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_frames=16, stack_size=3)      # multidim_stacker.py:155
    with pytest.raises(NotImplementedError):
        MultiDimStacker("resnet18", 2)


def test_no_cpu_fallback(oracle_sd):
    from ball_action_spotting_b200 import MultiDimStacker
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_3d_blocks=4, expansion_3d_ratio=3).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 15, 64, 64))


def test_bn_folding_reproduces_conv_bn(oracle_sd):
    from ball_action_spotting_b200.packer import pack_state_dict
    pk = pack_state_dict(oracle_sd, 4, bias_correction=False)
    g = torch.Generator().manual_seed(0)
    # dense 3x3 (blocks.1.1.conv_exp + bn1): packed [co][(r*3+s)*ci + c] fp16 / bias fp32
    x = torch.randn(1, 32, 12, 12, generator=g)
    p = "conv2d_encoder.blocks.1.1."
    ref = O._bn(oracle_sd, p + "bn1", F.conv2d(x, oracle_sd[p + "conv_exp.weight"], padding=1), O.ENC_BN_EPS)
    w = pk["b1.1.c3.w"].float().view(128, 3, 3, 32).permute(0, 3, 1, 2)
    got = F.conv2d(x, w, pk["b1.1.c3.b"], padding=1)
    assert (got - ref).abs().max() / ref.abs().max() < 1e-3      # fp16 weight rounding only
    # depthwise 3x3x3 (conv3d_encoder.0.conv_dw + bn2): packed [taps][C] fp32 — exact up to fp32 rounding
    x3 = torch.randn(1, 576, 3, 5, 5, generator=g)
    p = "conv3d_encoder.0."
    ref = O._bn(oracle_sd, p + "bn2.bn3d", F.conv3d(x3, oracle_sd[p + "conv_dw.weight"], padding=1, groups=576), O.REF_BN_EPS)
    w = pk["c3d.0.dw.w"].t().reshape(576, 1, 3, 3, 3)
    got = F.conv3d(x3, w, pk["c3d.0.dw.b"], padding=1, groups=576)
    assert (got - ref).abs().max() / ref.abs().max() < 1e-5
    # stem [27][32]
    xs = torch.rand(1, 3, 16, 16, generator=g)
    ref = O._bn(oracle_sd, "conv2d_encoder.bn1", O._conv_same(xs, oracle_sd["conv2d_encoder.conv_stem.weight"], 2), O.ENC_BN_EPS)
    w = (pk["stem.wh"][0].float() + pk["stem.wh"][1].float())[:, :27].reshape(32, 3, 3, 3)
    assert float(pk["stem.wh"][:, :, 27:].abs().max()) == 0.0
    got = F.conv2d(F.pad(xs, (0, 1, 0, 1)), w, pk["stem.b"], stride=2)
    assert (got - ref).abs().max() / ref.abs().max() < 1e-5
    assert pk["b5.7.se.w2t"].shape == (48, 1152) and pk["cls.w"].shape == (2, 1280)


def test_host_mirrors_equal_oracle_restatements():
    from ball_action_spotting_b200 import StackIndexesGenerator, get_frames_processor
    for size, step in [(15, 2), (33, 2), (5, 1)]:
        gen = StackIndexesGenerator(size, step)
        for i in (-2, 0, 40, 99):
            assert gen.make_stack_indexes(i) == O.make_stack_indexes(i, size, step)
            assert gen.clip_index(i, 100, 1) == O.clip_index(i, 100, size, step, 1)
    u8 = torch.randint(0, 256, (2, 3, 720, 1280), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    proc = get_frames_processor("pad_normalize", dict(size=(1280, 736), pad_mode="constant", fill_value=0))
    assert torch.equal(proc(u8), O.pad_normalize(u8, (1280, 736)))


def test_bias_correction_is_tiny_and_data_free(oracle_sd):
    """The fp16 rounding-bias correction only nudges biases (<< one fp16 ulp of the activations) and touches nothing else."""
    from ball_action_spotting_b200.packer import pack_state_dict
    a = pack_state_dict(oracle_sd, 4, bias_correction=False)
    b = pack_state_dict(oracle_sd, 4, bias_correction=True)
    assert a.keys() == b.keys()
    changed = [k for k in a if not torch.equal(a[k], b[k])]
    assert changed and all(k.endswith(".b") or k.endswith(".bm") for k in changed)
    for k in changed:
        if k.endswith(".b"):
            assert (a[k] - b[k]).abs().max() < 5e-3
    assert not any(".pwl" in k and k.startswith(("b3", "b4", "b5", "c3d")) for k in changed)     # SE-gated layers are untouched


# ---- sweep window assembly on a fake engine (CPU): the index arithmetic of SlidingSweep.predict_range --------------------
class _FakeDesc:
    def __init__(self, frames, img_stride, plane_stride, hflip, offset_elems):
        self.frames, self.img_stride, self.plane_stride, self.hflip, self.offset = frames, img_stride, plane_stride, hflip, offset_elems


class _FakeEngine:
    """Stands in for engine.Engine: an image's 'feature' is the per-frame mean of its three planes (flipped or not, which
    the mean ignores, so the flip branch is marked by a sign), the 3D stage is the identity and the head sums over T with
    position weights -- enough to tell whether every prediction sees exactly the triples the reference predictor would."""

    def __init__(self, hw):
        self.hw = hw

    def frames_desc(self, frames, H, W, img_stride, plane_stride, hflip=False, offset_elems=0):
        return _FakeDesc(frames, img_stride, plane_stride, hflip, offset_elems)

    def forward_2d(self, desc, n_img):
        import torch
        flat = desc.frames.reshape(-1).float()
        hw = self.hw
        out = torch.empty((n_img, 1, 1, 3))
        for i in range(n_img):
            for c in range(3):
                o = desc.offset + i * desc.img_stride + c * desc.plane_stride
                out[i, 0, 0, c] = flat[o:o + hw].mean() * (-1.0 if desc.hflip else 1.0)
        return out

    def gather_stacks(self, feats, first_image, hop, n_pred, T):
        import torch
        idx = first_image + torch.arange(n_pred)[:, None] + hop * torch.arange(T)[None, :]
        return feats[idx.reshape(-1)].view(n_pred, T, *feats.shape[1:])

    def forward_3d(self, x):
        return x

    def forward_head(self, x, sigmoid=False):
        import torch
        w = torch.arange(1, x.shape[1] * 3 + 1, dtype=torch.float32).view(1, x.shape[1], 3)
        s = (x.reshape(x.shape[0], x.shape[1], 3) * w).sum((1, 2))
        return torch.stack([s, -s], 1)

    def axpby_(self, y, x, a, b):
        y.mul_(a).add_(x, alpha=b)
        return y


class _FakeModule:
    stack_size = 3

    class _cfg:
        num_classes = 2

    def __init__(self, hw):
        self._eng = _FakeEngine(hw)

    def engine(self, device):
        return self._eng


@pytest.mark.parametrize("tta", [False, True])
def test_sweep_assembles_the_same_windows_as_the_streaming_predictor(tta):
    import torch
    from ball_action_spotting_b200.indexes import StackIndexesGenerator
    from ball_action_spotting_b200.sweep import SlidingSweep
    h, w, n = 2, 4, 70
    frames = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    sweep = SlidingSweep(_FakeModule(h * w), 15, 2, (w, h), tta=tta, max_stacks=7)        # several ragged chunks
    first_frame, a, b = 100, 120, 143                                                       # buffer = video frames 100..169
    got = sweep.predict_range(frames, first_frame, a, b)
    gen = StackIndexesGenerator(15, 2)
    means = frames.float().mean((1, 2))
    weights = torch.arange(1, 16, dtype=torch.float32)
    for j, p in enumerate(range(a, b)):
        idx = gen.make_stack_indexes(p)                                                      # predictors.py:56
        assert idx[0] == p - 14 and idx[-1] == p + 14
        s = (means[[i - first_frame for i in idx]] * weights).sum()
        want = torch.stack([s, -s]) * (0.0 if tta else 1.0)       # the fake flip branch negates: the TTA mean cancels exactly
        assert torch.allclose(got[j], want, rtol=1e-5, atol=1e-3), (p, got[j], want)
    with pytest.raises(RuntimeError, match="need frames"):
        sweep.predict_range(frames, first_frame, 110, 120)                                   # halo outside the buffer


def test_predictor_rings_follow_the_reference_dict_semantics():
    """predictor._Rings (device rings addressed by index % size + host tags) against the reference's two dicts
    (src/predictors.py:33-48, 53-67): same 'window complete' decisions and the same cache misses, frame by frame,
    including a reset and a jump in the frame index."""
    from types import SimpleNamespace
    from ball_action_spotting_b200.indexes import StackIndexesGenerator
    from ball_action_spotting_b200.predictor import _Rings, batched
    gen = StackIndexesGenerator(15, 2)
    fake = SimpleNamespace(device=torch.device("cpu"), image_size=(64, 32), indexes_generator=gen, model_stack_size=3,
                           frame_stack_step=2, tta=True, num_classes=2)
    rings = _Rings(fake, 30, 64)
    assert (rings.nf, rings.nt) == (29, 25) and rings.feats.shape == (2, 25, 1, 2, 192)
    frames_d, triples_d = {}, {}
    seq = list(range(0, 70)) + list(range(500, 560))          # a jump: everything cached becomes stale
    for index in seq:
        p = index - 14
        idx = gen.make_stack_indexes(p)
        frames_d[index] = True                                                    # reference: predictors.py:53-56
        for k in [k for k in frames_d if k < idx[0]]:
            del frames_d[k]
        for k in [k for k in triples_d if any(i < idx[0] for i in k)]:
            del triples_d[k]
        ref_ready = set(idx) <= set(frames_d)
        ref_miss = []
        if ref_ready:
            for tr in batched(idx, 3):
                if tr not in triples_d:
                    triples_d[tr] = True
                    ref_miss.append(tr[0])
        rings.clear_old(idx[0])                                                   # ours
        rings.put_frame(torch.full((30, 64), index % 251, dtype=torch.uint8), index)
        ready = all(rings.has_frame(i) for i in idx)
        miss = []
        if ready:
            for tr in batched(idx, 3):
                if not rings.has_triple(tr[0]):
                    rings.put_triple(tr[0], torch.full((2, 1, 2, 192), float(tr[0] % 100), dtype=torch.float16))
                    miss.append(tr[0])
            for tr in batched(idx, 3):                                            # slots hold what was stored for that triple
                assert float(rings.triple(tr[0])[0, 0, 0, 0]) == float(tr[0] % 100)
            assert int(rings.frame(idx[0])[0, 0]) == idx[0] % 251
        assert ready == ref_ready and miss == ref_miss, (index, miss, ref_miss)
    assert miss == [idx[12]]                                   # steady state: one new triple per frame
    rings.reset()
    assert not any(rings.has_frame(i) for i in range(560))
    o1, o2 = rings.emit(torch.tensor([0.25, 0.75])), rings.emit(torch.tensor([0.5, 0.5]))
    assert o1.tolist() == [0.25, 0.75] and o2.tolist() == [0.5, 0.5]              # rotating output slots do not alias


def test_accounting_is_consistent_across_tail_modes():
    """accounting.py feeds bench.py's roofline: the launch lists of the MBConv-tail schemes must describe the same work."""
    from ball_action_spotting_b200 import accounting as acc
    t3, t1 = acc.per_stack_totals(tail_mode=2), acc.per_stack_totals(tail_mode=1)
    t0 = acc.per_stack_totals(tail_mode=0)
    flops = lambda t: sum(v["flops"] for v in t.values())
    assert abs(flops(t3) - flops(t1)) < 1e-6 * flops(t3) and abs(flops(t3) - flops(t0)) < 1e-6 * flops(t3)
    assert abs(flops(t3) / 141.6e9 - 1) < 0.01                                  # 141.6 GFLOP per stack
    assert t0["se_fc"]["launches"] == 20 and "se_fc" not in t3 and t1["tail2d"]["launches"] == 16 and t1["tail3d"]["launches"] == 4
    dw = acc.depthwise_only_bytes()
    assert abs(dw["dwconv2d"] / (5 * 102.3e6) - 1) < 0.01 and abs(dw["dwconv3d"] / 42.4e6 - 1) < 0.01   # SURVEY 8(d) DW-only bytes
    # one launch of every encoder kernel per image chunk: 1 stem + 5 conv3x3 + 16 x (pw, dw, se, pwl) + proj = 71 (mode 0)
    assert len(acc.encoder_launches(736, 1280, 720, tail_mode=0)) == 71 and len(acc.encoder_launches(736, 1280, 720, tail_mode=1)) == 39


def test_stem_tc_operand_layout_reproduces_the_conv():
    """CPU emulation of stem_tc_kernel's operands (csrc/stem_tc.cuh): the K order k' = (ci*3 + r)*4 + s with a zero-weight fourth slot,
    the two-pixels-per-thread byte gather (pixel A = x0 x1 x2 [x3], pixel B = x2 x3 x4 [x5]), the mirrored tile of the hflip TTA
    (tile byte j = logical column 2*ox0 + 143 - j) and the weight tile built from packer.stem_weights.  The dot products must equal
    the TF-SAME stride-2 convolution of the zero-padded, /255-normalised frames (frames.py:7-31 + timm conv_stem)."""
    import torch.nn.functional as F
    from ball_action_spotting_b200.packer import stem_weights
    g = torch.Generator().manual_seed(0)
    W_, stored_h, H = 160, 20, 24
    pad_top = (H - stored_h) // 2
    raw = torch.randint(0, 256, (3, stored_h, W_), dtype=torch.uint8, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.3
    wh = stem_weights(w).double()                                   # [2][32][32], k = (ci*3 + r)*3 + s
    wk = torch.zeros(2, 32, 48, dtype=torch.float64)                # kernel weight tile: k' = combo*4 + s, combo = ci*3 + r
    for combo in range(9):
        for s in range(3):
            wk[:, :, combo * 4 + s] = wh[:, :, combo * 3 + s]
    x = F.pad(raw.double(), (0, 0, pad_top, H - stored_h - pad_top))                       # zero rows = TMA out-of-bounds fill
    IW = 144
    for hflip in (0, 1):
        xin = x.flip(-1) if hflip else x
        ref = F.conv2d(F.pad(xin[None] / 255.0, (0, 1, 0, 1)), w.double(), stride=2)[0]   # [32][H/2][W/2]
        for ox0 in (0, 64):
            start = W_ - IW - 2 * ox0 if hflip else 2 * ox0                               # TMA x coordinate of the tile
            for oy in (0, 5, H // 2 - 1):
                tile = torch.zeros(3, 3, IW, dtype=torch.float64)                         # rows 2*oy + r of the three planes
                for r in range(3):
                    y = 2 * oy + r
                    for j in range(IW):
                        xx = start + j
                        if 0 <= xx < W_ and y < H:
                            tile[:, r, j] = x[:, y, xx]
                for cp in (0, 7, 31):                                                     # pixel pair (2cp, 2cp+1) of the tile row
                    a = torch.zeros(2, 48, dtype=torch.float64)
                    for combo in range(9):
                        ci, r = divmod(combo, 3)
                        row = tile[ci, r]
                        if hflip:
                            xs = [row[143 - (4 * cp + i)] for i in range(6)]
                        else:
                            xs = [row[4 * cp + i] for i in range(6)]
                        a[0, combo * 4:combo * 4 + 4] = torch.stack(xs[0:4])               # x0 x1 x2 [x3]
                        a[1, combo * 4:combo * 4 + 4] = torch.stack(xs[2:6])               # x2 x3 x4 [x5]
                    acc = a @ (wk[0] + wk[1]).t()                                          # [2 pixels][32 cout]
                    for px in range(2):
                        ox = ox0 + 2 * cp + px
                        if ox < W_ // 2:
                            got = acc[px] / 255.0
                            assert torch.allclose(got, ref[:, oy, ox], rtol=0, atol=2e-6), (hflip, ox0, oy, cp, px)


def test_conv_tc_index_arithmetic():
    """CPU emulation of conv_tc_kernel's implicit-GEMM addressing (csrc/conv_tc.cuh) on a random single-channel image:
    (a) stride 2: the input is read as its four (row, column) parity phases, each a dense (TW+1) x (TH+1) tile with TMA zero fill, and
        tap (r, s) of output pixel p (linear index in the phase tile) is phase (r&1, s&1) at p + (r>>1)*(TW+1) + (s>>1);
    (b) stride 1 with the column taps folded into N: D[p][s] = sum_r in[p + r*PW] * w[r][s] on the UNSHIFTED rows of a PW = 32 wide tile,
        and output column c of a tile row is D[c][0] + D[c+1][1] + D[c+2][2] (lanes c, c+1, c+2 of one warp)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    H, W = 20, 44
    x = torch.randn(H, W, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, generator=g, dtype=torch.float64)
    # (a) stride 2, TF-SAME on even sizes = pad bottom / right
    ref2 = F.conv2d(F.pad(x[None, None], (0, 1, 0, 1)), w[None, None], stride=2)[0, 0]
    TW, TH = 8, 4
    PW, PH = TW + 1, TH + 1
    for y0 in range(0, H // 2, TH):
        for x0 in range(0, W // 2, TW):
            phases = torch.zeros(4, PH * PW + PW + 2, dtype=torch.float64)      # + slack: the last M tile over-reads
            for py in range(2):
                for px in range(2):
                    for yy in range(PH):
                        for xx in range(PW):
                            iy, ix = 2 * (y0 + yy) + py, 2 * (x0 + xx) + px
                            if iy < H and ix < W:                                # out of bounds = TMA zero fill
                                phases[py * 2 + px, yy * PW + xx] = x[iy, ix]
            for ry in range(TH):
                for cx in range(TW):
                    oy, ox = y0 + ry, x0 + cx
                    if oy >= H // 2 or ox >= W // 2:
                        continue
                    p = ry * PW + cx
                    acc = sum(w[r, s] * phases[(r & 1) * 2 + (s & 1), p + (r >> 1) * PW + (s >> 1)] for r in range(3) for s in range(3))
                    assert abs(acc - ref2[oy, ox]) < 1e-12
    # (b) stride 1, pad 1, column taps folded into N
    ref1 = F.conv2d(x[None, None], w[None, None], padding=1)[0, 0]
    TW, TH, PW = 30, 16, 32
    for y0 in range(0, H, TH):
        for x0 in range(0, W, TW):
            tile = torch.zeros((TH + 2) * PW, dtype=torch.float64)
            for yy in range(TH + 2):
                for xx in range(PW):
                    iy, ix = y0 - 1 + yy, x0 - 1 + xx
                    if 0 <= iy < H and 0 <= ix < W:
                        tile[yy * PW + xx] = x[iy, ix]
            for ry in range(TH):
                D = torch.zeros(PW, 3, dtype=torch.float64)                       # one tile row = one warp = 32 lanes
                for lane in range(PW):
                    p = ry * PW + lane
                    for s in range(3):
                        D[lane, s] = sum(tile[p + r * PW] * w[r, s] for r in range(3))
                for c in range(TW):
                    oy, ox = y0 + ry, x0 + c
                    if oy < H and ox < W:
                        assert abs(D[c, 0] + D[c + 1, 1] + D[c + 2, 2] - ref1[oy, ox]) < 1e-12


def test_silu4_algebra_and_clamp():
    """float32 emulation of silu4 (csrc/common.cuh): SiLU of four values with one reciprocal, 1/(1+e_i) = prod_{j != i}(1+e_j) / prod_j(1+e_j),
    inputs clamped at -20 so that the product of the four (1 + e) factors stays finite.  Against float64 SiLU: relative error < 2e-6 for
    x >= -20 (before the SFU approximations, which add ~2 ulp; fp16 storage rounds at 4.9e-4) and an absolute error below the smallest fp16 subnormal beyond the clamp."""
    import numpy as np
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 3, 4000), rng.uniform(-20, 20, 4000), np.array([-20.0, -19.999, 0.0, 1e-8, 30.0, 88.0, -1e-3, 5.5])])
    x = x[: len(x) // 4 * 4].astype(np.float32).reshape(-1, 4)
    tail = np.array([[-25.0, -100.0, -1e4, -20.5], [-3e38, 0.0, 3.0, -50.0]], dtype=np.float32)

    def silu4(v):
        v = np.maximum(v, np.float32(-20.0))
        e = np.exp2((v * np.float32(-1.4426950408889634)).astype(np.float32)).astype(np.float32)
        d = (np.float32(1.0) + e).astype(np.float32)
        p01, p23 = d[:, 0] * d[:, 1], d[:, 2] * d[:, 3]
        r = (np.float32(1.0) / (p01 * p23)).astype(np.float32)
        r01, r23 = r * p23, r * p01
        out = np.stack([v[:, 0] * (r01 * d[:, 1]), v[:, 1] * (r01 * d[:, 0]), v[:, 2] * (r23 * d[:, 3]), v[:, 3] * (r23 * d[:, 2])], 1)
        return out.astype(np.float32)

    got = silu4(x).astype(np.float64)
    xd = x.astype(np.float64)
    ref = xd / (1.0 + np.exp(-xd))
    assert np.all(np.isfinite(got))
    nz = np.abs(ref) > 1e-30
    assert np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])) < 2e-6
    gt = silu4(tail).astype(np.float64)
    td = tail.astype(np.float64)
    rt = td / (1.0 + np.exp(-np.maximum(td, -700.0)))
    beyond = td < -20.0
    assert np.all(np.isfinite(gt)) and np.max(np.abs(gt[beyond] - rt[beyond])) < 6e-8     # < the smallest fp16 subnormal (5.96e-8)
    assert np.max(np.abs(gt[~beyond] - rt[~beyond])) < 1e-6


def test_tcgen05_epilogue_partitions_cover_every_column_once():
    """Mirrors of the index arithmetic of the tensor-memory epilogues: every accumulator column is read by exactly one warp.
    gemm_tc_kernel: warp `half` of `nhalf` takes the 16-column chunks half + nhalf*k, k < nk = ceil((BN/16 - half) / nhalf).
    conv_tc_kernel: PARTS warps per quadrant take CMID/PARTS consecutive columns in 8-column chunks; the projection's 16-column groups
    g = part + PARTS*j.  conv_tc_ws_kernel: 4 parts x 48 columns, packed P columns (col / 2) stay inside the low half of the accumulator."""
    for BN in range(32, 257, 16):
        for nhalf in (1, 2):
            ngroups = BN // 16
            seen = []
            for half in range(nhalf):
                nk = (ngroups - half + nhalf - 1) // nhalf
                seen += [half + nhalf * k for k in range(nk)]
            assert sorted(seen) == list(range(ngroups)), (BN, nhalf)
    for cmid, cproj in ((16, 0), (64, 32), (128, 32), (128, 48)):
        parts = 2 if cmid >= 32 else 1
        cols = cmid // parts
        assert cols % 16 == 0 and (cols // 8) % 2 == 0
        covered = sorted(c for part in range(parts) for ch in range(cols // 8) for c in range(part * cols + ch * 8, part * cols + ch * 8 + 8))
        assert covered == list(range(cmid))
        if cproj:
            ng2 = cproj // 16
            nj = 2 if ng2 > parts else 1
            groups = sorted(part + parts * j for part in range(parts) for j in range(nj) if part + parts * j < ng2)
            assert groups == list(range(ng2))
    cmid, parts = 192, 4
    cols = cmid // parts
    pcols = sorted(c for part in range(parts) for j in range(cols // 16) for c in range((part * cols + j * 16) // 2, (part * cols + j * 16) // 2 + 8))
    assert pcols == list(range(cmid // 2))                  # 96 packed columns = the low half of the 192-column accumulator


def test_conv_tc_tile_geometry_stays_inside_the_tile_allocation():
    """ConvTcCfg / ConvWsCfg (csrc/conv_tc.cuh, conv_tc_ws.cuh): the A operand of M tile m, tap (r, s) starts at pixel m*128 + off and spans
    128 rows; with the OVER slack the last M tile never reads past the allocation of its plane set."""
    # (stride, TW, TH, fold)
    for stride, TW, TH, fold in ((1, 30, 16, True), (1, 32, 15, False), (2, 32, 15, False), (2, 32, 7, False), (2, 32, 3, False)):
        PW = TW + 2 if stride == 1 else TW + 1
        PH = TH + 2 if stride == 1 else TH + 1
        PIX = PW * PH
        MT = (TH * PW + 127) // 128
        max_off = (2 * PW + 2) if stride == 1 else (PW + 1)
        over = max(0, MT * 128 + max_off - PIX)
        offs = [r * PW for r in range(3)] if fold else ([r * PW + s for r in range(3) for s in range(3)] if stride == 1
                                                        else [(r >> 1) * PW + (s >> 1) for r in range(3) for s in range(3)])
        last_read = (MT - 1) * 128 + max(offs) + 127
        assert last_read < PIX + over, (stride, TW, TH, fold)
        if fold:
            assert PW == 32 and TH * PW % 128 == 0 and last_read < PIX      # tile rows = TMEM lane quadrants, nothing read past the tile
        # every valid output pixel of the tile lies in an M tile that is computed
        rows = TH
        nm = (rows * PW + 127) // 128
        assert (rows - 1) * PW + (TW - 1) < nm * 128
