"""GPU: the MBConv tail kernels — depthwise + BN + SiLU + SE squeeze/excitation (mds_k_dwconv_se, TMA-staged), the gating
projection GEMM (mds_k_gemm_gate) and the single-launch fused variant (mds_k_mbconv_tail) — against the fp32 torch ops of the oracle's layer (timm InvertedResidual conv_dw..bn3, multidim_stacker.py:110-134) on
identical fp16-rounded inputs.  Tolerance (north_star): max|got - ref| / max|ref| <= 1e-3 on every output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-3
DEV = "cuda:0"


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def rel(got, ref):
    return ((got.float().cpu() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-12)).item()


def make_case(n, T, H, W, C_, kt, stride, rd, N, res, seed=0):
    x = torch.randn((n, C_, T, H, W), generator=gen(seed + 1)).to(torch.float16)
    taps = 9 * kt
    if kt == 1:
        w = torch.randn(C_, 1, 3, 3, generator=gen(seed + 2)) * 0.25
    else:
        w = torch.randn(C_, 1, 3, 3, 3, generator=gen(seed + 2)) * 0.15
    b = torch.randn(C_, generator=gen(seed + 3)) * 0.1
    w1 = torch.randn(rd, C_, generator=gen(seed + 4)) * (2.0 / rd) ** 0.5 * 0.2
    b1 = torch.randn(rd, generator=gen(seed + 5)) * 0.1
    w2 = torch.randn(C_, rd, generator=gen(seed + 6)) * (2.0 / C_) ** 0.5
    b2 = torch.randn(C_, generator=gen(seed + 7)) * 0.1
    wp = (torch.randn(N, C_, generator=gen(seed + 8)) * (1.0 / C_) ** 0.5).to(torch.float16)
    bp = torch.randn(N, generator=gen(seed + 9)) * 0.1
    # reference (fp32 torch ops)
    xf = x.float()
    if kt == 1:
        xi = xf[:, :, 0]
        if stride == 2:
            xi = F.pad(xi, (0, 1, 0, 1))                      # TF-SAME on even input
            y = F.conv2d(xi, w, b, stride=2, groups=C_)
        else:
            y = F.conv2d(xi, w, b, padding=1, groups=C_)
        y = F.silu(y)[:, :, None]                             # (n, C, 1, Ho, Wo)
    else:
        y = F.silu(F.conv3d(xf, w, b, padding=1, groups=C_))
    mean = y.mean((2, 3, 4))
    gate = torch.sigmoid(F.silu(mean @ w1.t() + b1) @ w2.t() + b2)
    yg = y * gate[:, :, None, None, None]
    rows = yg.permute(0, 2, 3, 4, 1).reshape(n, -1, C_)      # (n, T*Ho*Wo, C)
    out = rows @ wp.float().t() + bp
    r = None
    if res:
        r = torch.randn(out.shape, generator=gen(seed + 10)).to(torch.float16)
        out = out + r.float()
    return dict(x=x, w=w, b=b, w1=w1, b1=b1, w2=w2, b2=b2, wp=wp, bp=bp, r=r, y=y, gate=gate, out=out, taps=taps)


def run_tail(lib, case, n, T, H, W, C_, kt, stride, rd, N, rows_per_chunk=0, with_proj=True):
    from ball_action_spotting_b200.packer import bias_matrix
    Ho, Wo = H // stride, W // stride
    d = lambda t: t.contiguous().to(DEV)
    m1 = d(case["x"].permute(0, 2, 3, 4, 1))                                   # (n, T, H, W, C)
    dw_w = d(case["w"].reshape(C_, case["taps"]).t())                          # [taps][C]
    m2 = torch.zeros((n, T, Ho, Wo, C_), dtype=torch.float16, device=DEV)
    parts = torch.full((n, 64, C_), float("nan"), dtype=torch.float32, device=DEV)
    gate = torch.zeros((n, C_), dtype=torch.float32, device=DEV)
    sync = torch.zeros((3, n), dtype=torch.int32, device=DEV)
    out = torch.zeros((n, T * Ho * Wo, N), dtype=torch.float16, device=DEV)
    t = {k: d(case[k]) for k in ("b", "w1", "b1", "b2", "wp")}
    w2t = d(case["w2"].t())
    bm = d(bias_matrix(case["bp"]))
    r = None if case["r"] is None else d(case["r"])
    rc = lib.mds_k_mbconv_tail(m1.data_ptr(), m2.data_ptr(), dw_w.data_ptr(), t["b"].data_ptr(), parts.data_ptr(),
                               t["w1"].data_ptr(), t["b1"].data_ptr(), w2t.data_ptr(), t["b2"].data_ptr(), gate.data_ptr(),
                               sync.data_ptr(), t["wp"].data_ptr(), bm.data_ptr(), None if r is None else r.data_ptr(),
                               out.data_ptr(), n, T, H, W, C_, kt, stride, rd, N if with_proj else 0, rows_per_chunk, None)
    assert rc == 0, lib.mds_last_error().decode()
    torch.cuda.synchronize()
    return m2, gate, out, sync


CASES = [
    # n, T, H,  W,  C,    kt, stride, rd, N,  res
    (3, 1, 16, 44, 192, 1, 1, 12, 96, 0),
    (3, 1, 20, 48, 192, 1, 2, 12, 96, 0),
    (2, 1, 23, 40, 1152, 1, 1, 48, 192, 1),
    (2, 1, 46, 80, 672, 1, 1, 28, 112, 1),
    (2, 1, 46, 80, 672, 1, 2, 28, 192, 0),
    (5, 1, 46, 80, 384, 1, 1, 24, 96, 1),
    (2, 5, 7, 9, 576, 3, 1, 24, 192, 1),
    (2, 5, 23, 40, 576, 3, 1, 24, 192, 1),
    (1, 11, 23, 40, 576, 3, 1, 24, 192, 1),
]


@pytest.mark.parametrize("n,T,H,W,C_,kt,stride,rd,N,res", CASES)
def test_depthwise_and_se_only(lib, n, T, H, W, C_, kt, stride, rd, N, res):
    """N = 0 mode: depthwise output and SE gate (no projection items, no tcgen05)."""
    case = make_case(n, T, H, W, C_, kt, stride, rd, N, res)
    m2, gate, _, _ = run_tail(lib, case, n, T, H, W, C_, kt, stride, rd, N, with_proj=False)
    assert rel(m2.permute(0, 4, 1, 2, 3), case["y"]) <= TOL
    assert rel(gate, case["gate"]) <= TOL


@pytest.mark.parametrize("n,T,H,W,C_,kt,stride,rd,N,res", CASES)
def test_fused_tail(lib, n, T, H, W, C_, kt, stride, rd, N, res):
    case = make_case(n, T, H, W, C_, kt, stride, rd, N, res)
    m2, gate, out, sync = run_tail(lib, case, n, T, H, W, C_, kt, stride, rd, N)
    assert rel(m2.permute(0, 4, 1, 2, 3), case["y"]) <= TOL
    assert rel(gate, case["gate"]) <= TOL
    assert rel(out, case["out"]) <= TOL
    assert int(sync.abs().sum()) == 0                       # the sync words are left clean for the next launch
    # bit-reproducible, and independent of the other images in the batch
    m2b, gateb, outb, _ = run_tail(lib, case, n, T, H, W, C_, kt, stride, rd, N)
    assert torch.equal(out, outb) and torch.equal(gate, gateb) and torch.equal(m2, m2b)
    if n > 1:
        one = {k: (v[:1] if k in ("x", "r") and v is not None else v) for k, v in case.items()}
        _, gate1, out1, _ = run_tail(lib, one, 1, T, H, W, C_, kt, stride, rd, N)
        assert torch.equal(out1[0], out[0]) and torch.equal(gate1[0], gate[0])


def test_many_images(lib):
    """Many items per CTA in both streams: exercises the ring phases across items and the per-image flags."""
    n, T, H, W, C_, kt, stride, rd, N = 40, 1, 23, 40, 384, 1, 1, 24, 96
    case = make_case(n, T, H, W, C_, kt, stride, rd, N, 1)
    _, gate, out, sync = run_tail(lib, case, n, T, H, W, C_, kt, stride, rd, N, rows_per_chunk=6)
    assert int(sync.abs().sum()) == 0
    assert rel(gate, case["gate"]) <= TOL
    assert rel(out, case["out"]) <= TOL
    assert int(sync.abs().sum()) == 0


def run_two_launch(lib, case, n, T, H, W, C_, kt, stride, rd, N, rows_per_chunk=0):
    """mds_k_dwconv_se followed by mds_k_gemm_gate: the default path of the engine (mds_set_tail_mode(2))."""
    import ctypes as C
    from ball_action_spotting_b200.packer import bias_matrix
    Ho, Wo = H // stride, W // stride
    d = lambda t: t.contiguous().to(DEV)
    m1 = d(case["x"].permute(0, 2, 3, 4, 1))
    dw_w = d(case["w"].reshape(C_, case["taps"]).t())
    m2 = torch.zeros((n, T, Ho, Wo, C_), dtype=torch.float16, device=DEV)
    parts = torch.full((n, 64, C_), float("nan"), dtype=torch.float32, device=DEV)
    gate = torch.zeros((n, C_), dtype=torch.float32, device=DEV)
    done = torch.zeros((n,), dtype=torch.int32, device=DEV)
    out = torch.zeros((n, T * Ho * Wo, N), dtype=torch.float16, device=DEV)
    t = {k: d(case[k]) for k in ("b", "w1", "b1", "b2", "wp")}
    w2t = d(case["w2"].t())
    bm = d(bias_matrix(case["bp"]))
    r = None if case["r"] is None else d(case["r"])
    nparts = C.c_int(0)
    rc = lib.mds_k_dwconv_se(m1.data_ptr(), m2.data_ptr(), dw_w.data_ptr(), t["b"].data_ptr(), parts.data_ptr(), C.byref(nparts),
                             t["w1"].data_ptr(), t["b1"].data_ptr(), w2t.data_ptr(), t["b2"].data_ptr(), gate.data_ptr(),
                             done.data_ptr(), n, T, H, W, C_, kt, stride, rd, rows_per_chunk, None)
    assert rc == 0, lib.mds_last_error().decode()
    rc = lib.mds_k_gemm_gate(m2.data_ptr(), t["wp"].data_ptr(), gate.data_ptr(), bm.data_ptr(), None if r is None else r.data_ptr(),
                             out.data_ptr(), T * Ho * Wo, n, N, C_, 0, None)
    assert rc == 0, lib.mds_last_error().decode()
    torch.cuda.synchronize()
    return m2, gate, out, done, parts.view(-1)[: n * nparts.value * C_].view(n, nparts.value, C_)      # [n][nparts][C] contiguous


@pytest.mark.parametrize("n,T,H,W,C_,kt,stride,rd,N,res", CASES)
def test_dwconv_se_and_gating_gemm(lib, n, T, H, W, C_, kt, stride, rd, N, res):
    case = make_case(n, T, H, W, C_, kt, stride, rd, N, res)
    m2, gate, out, done, parts = run_two_launch(lib, case, n, T, H, W, C_, kt, stride, rd, N)
    assert rel(m2.permute(0, 4, 1, 2, 3), case["y"]) <= TOL
    assert rel(parts.sum(1), case["y"].sum((2, 3, 4))) <= 1e-4        # fp32 squeeze sums of the un-rounded SiLU output
    assert rel(gate, case["gate"]) <= TOL
    assert rel(out, case["out"]) <= TOL
    assert int(done.abs().sum()) == 0
    m2b, gateb, outb, _, _ = run_two_launch(lib, case, n, T, H, W, C_, kt, stride, rd, N)
    assert torch.equal(out, outb) and torch.equal(gate, gateb) and torch.equal(m2, m2b)      # bit-reproducible
    if n > 1:                                                                                # and independent of the batch
        one = {k: (v[:1] if k in ("x", "r") and v is not None else v) for k, v in case.items()}
        _, gate1, out1, _, _ = run_two_launch(lib, one, 1, T, H, W, C_, kt, stride, rd, N)
        assert torch.equal(out1[0], out[0]) and torch.equal(gate1[0], gate[0])


def test_dwconv_se_row_chunks(lib):
    """Ragged row chunks (Ho not a multiple of the chunk) and single-row chunks."""
    n, T, H, W, C_, kt, stride, rd, N = 2, 1, 23, 40, 192, 1, 1, 12, 96
    case = make_case(n, T, H, W, C_, kt, stride, rd, N, 0)
    for rows in (1, 5, 23, 64):
        m2, gate, out, done, _ = run_two_launch(lib, case, n, T, H, W, C_, kt, stride, rd, N, rows_per_chunk=rows)
        assert rel(m2.permute(0, 4, 1, 2, 3), case["y"]) <= TOL, rows
        assert rel(gate, case["gate"]) <= TOL and rel(out, case["out"]) <= TOL, rows
