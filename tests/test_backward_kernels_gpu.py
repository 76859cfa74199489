"""GPU parity of the training-path kernels against torch autograd on the CPU (fp64 where it matters):
norm+activation backward, weight gradient (tcgen05, MN-major operands, split-K), data gradient (= conv_tc with the
layer's own weight tensor in the transposed role)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import ops as _ops
    return _ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(shape, generator=g) * 2 - 1) * scale


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def split16(x, dtype=torch.bfloat16):
    hi = x.to(dtype)
    return hi, (x - hi.float()).to(dtype)


@pytest.mark.parametrize("N,H,W,C,pooled,act", [(2, 16, 16, 64, False, 1), (3, 9, 13, 128, True, 2), (1, 32, 32, 256, False, 0),
                                                 (2, 8, 8, 512, True, 1)])
def test_norm_bwd(ops, N, H, W, C, pooled, act):
    y = (_rand((N, C, H, W), 1) * 2 + _rand((1, C, 1, 1), 2)).double().requires_grad_(True)
    gamma = (1 + 0.2 * _rand((C,), 3)).double().requires_grad_(True)
    beta = (0.2 * _rand((C,), 4)).double().requires_grad_(True)
    dout = _rand((N, C, H, W), 5).double()
    dout2 = _rand((N, C, H, W), 6).double()
    if pooled:
        n = F.batch_norm(y, None, None, gamma, beta, True, 0.0, 1e-5)
    else:
        n = F.instance_norm(y, None, None, gamma, beta, True, 0.0, 1e-5)
    a = [n, F.relu(n), F.leaky_relu(n, 0.2)][act]
    a.backward(dout + dout2)
    yd = nhwc(y.detach().float()).cuda()
    sc, sh, mean, rstd = ops.norm_stats(yd, gamma.detach().float().cuda(), beta.detach().float().cuda(), pooled, want_stats=True)
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    f32, hi, lo = ops.norm_bwd(nhwc(dout.float()).cuda(), yd, sc, sh, mean, rstd, act, dout2=nhwc(dout2.float()).cuda(),
                               pooled=pooled, dgamma=dg, dbeta=db, want_f32=True, want_split=True)
    ref = y.grad
    tol = 2e-5 * max(1.0, ref.abs().max().item())
    assert (nchw(f32.cpu()).double() - ref).abs().max().item() < tol
    assert (nchw((hi.float() + lo.float()).cpu()).double() - ref).abs().max().item() < tol + 2 ** -15 * ref.abs().max().item()
    assert (dg.cpu().double() - gamma.grad).abs().max().item() < 1e-4 * max(1.0, gamma.grad.abs().max().item())
    assert (db.cpu().double() - beta.grad).abs().max().item() < 1e-4 * max(1.0, beta.grad.abs().max().item())
    # layer without norm: dy = dout * act'(y)
    f32b, _, _ = ops.norm_bwd(nhwc(dout.float()).cuda(), yd, act=act, want_f32=True, want_split=False)
    yy = y.detach()
    dref = dout * [torch.ones_like(yy), (yy > 0).double(), torch.where(yy > 0, 1.0, 0.2)][act]
    assert (nchw(f32b.cpu()).double() - dref).abs().max().item() < 1e-6


WG_CASES = [
    # name, N, H, W, Cin, Cout, R, stride, pad, transposed, outpad
    ("k3s1_64_64", 2, 16, 16, 64, 64, 3, 1, 1, False, 0),
    ("k3s1_256_256", 1, 32, 32, 256, 256, 3, 1, 1, False, 0),
    ("k3s2_64_128", 2, 32, 32, 64, 128, 3, 2, 1, False, 0),
    ("k4s2_128_256", 1, 16, 16, 128, 256, 4, 2, 1, False, 0),
    ("k4s1_256_512", 1, 16, 16, 256, 512, 4, 1, 1, False, 0),
    ("ct3s2_256_128", 2, 8, 8, 256, 128, 3, 2, 1, True, 1),
    ("ct4s2_128_64", 1, 16, 16, 128, 64, 4, 2, 1, True, 0),
    ("k7x1_64_64_stem", 1, 22, 40, 64, 64, (7, 1), 1, 0, False, 0),
    ("k3s1_w200", 1, 5, 200, 64, 64, 3, 1, 1, False, 0),
    ("k3s1_tiny", 3, 4, 4, 128, 64, 3, 1, 1, False, 0),
    # multi-tap kernel edge cases: extents that are not multiples of the 8 x 8 pixel tile, odd outputs, many K splits
    ("k4s1_63_dlayer", 1, 63, 63, 64, 128, 4, 1, 1, False, 0),
    ("k3s1_odd_21x19", 2, 21, 19, 64, 128, 3, 1, 1, False, 0),
    ("k4s2_64_64_dfirst", 2, 64, 64, 64, 64, 4, 2, 1, False, 0),
    ("ct3s2_128_64_up", 1, 24, 40, 128, 64, 3, 2, 1, True, 1),
    ("k7x1_head_big", 2, 70, 64, 64, 64, (7, 1), 1, 0, False, 0),
    ("k3s1_256_256_64x64", 2, 64, 64, 256, 256, 3, 1, 1, False, 0),
]


@pytest.mark.parametrize("case", WG_CASES, ids=[c[0] for c in WG_CASES])
def test_conv_wgrad_and_dgrad(ops, case):
    name, N, H, W, Cin, Cout, R, st, pad, tr, op = case
    R, S = (R if isinstance(R, tuple) else (R, R))
    x = _rand((N, Cin, H, W), 41).double().requires_grad_(True)
    w = _rand((Cin, Cout, R, S) if tr else (Cout, Cin, R, S), 42, 0.05).double().requires_grad_(True)
    y = F.conv_transpose2d(x, w, None, stride=st, padding=pad, output_padding=op) if tr else F.conv2d(x, w, None, stride=st, padding=pad)
    dy = _rand(tuple(y.shape), 43).double()
    y.backward(dy)
    d = ops.conv_desc(N, H, W, [Cin], Cout, R, S, st, pad, tr, op)
    xh, xl = split16(nhwc(x.detach().float()))
    dh, dl = split16(nhwc(dy.float()))
    # ---- weight gradient ---------------------------------------------------------------------------------------
    dw = ops.conv_wgrad(d, xh.cuda(), xl.cuda(), dh.cuda(), dl.cuda())
    torch.cuda.synchronize()
    ref = w.grad
    err = (dw.cpu().double() - ref).abs().max().item()
    print(f"{name}: wgrad max|d| {err:.3e} of scale {ref.abs().max().item():.3e}")
    assert err < 3e-5 * ref.abs().max().item()
    dw2 = ops.conv_wgrad(d, xh.cuda(), xl.cuda(), dh.cuda(), dl.cuda(), dw=dw.clone(), accumulate=True)
    assert (dw2.cpu().double() - 2 * ref).abs().max().item() < 6e-5 * ref.abs().max().item()
    # ---- data gradient: the forward weight tensor used in the opposite (transposed) role ---------------------------
    oh, ow = tuple(y.shape[2:])
    if tr:      # dgrad of ConvTranspose2d = Conv2d(dy, w) with the same stride / padding
        dd = ops.conv_desc(N, oh, ow, [Cout], Cin, R, S, st, pad, False, 0)
    else:       # dgrad of Conv2d = ConvTranspose2d(dy, w); output_padding restores the input extent
        oph = H - ((oh - 1) * st - 2 * pad + R)
        dd = ops.conv_desc(N, oh, ow, [Cout], Cin, R, S, st, pad, True, oph)
    w_hi, w_lo = ops.pack_weights_tc(dd, w.detach().float().cuda(), ops.FMT_BF16, True)
    dx = ops.conv_tc(dd, [dh.cuda()], [dl.cuda()], w_hi, w_lo, None, ops.FMT_BF16, True)
    errx = (nchw(dx.cpu()).double() - x.grad).abs().max().item()
    print(f"{name}: dgrad max|d| {errx:.3e} of scale {x.grad.abs().max().item():.3e}")
    assert tuple(dx.shape) == (N, H, W, Cin)
    assert errx < 6e-5 * x.grad.abs().max().item()      # K up to 8192: tensor-core fp32 accumulation ~3e-5 rel


def test_head_bwd_pack(ops):
    N, CO, H, W, S = 2, 3, 6, 10, 7
    dzz = _rand((N, CO, H, W), 51)
    hi, lo = ops.head_bwd_pack(dzz.cuda(), S)
    got = (hi.float() + lo.float()).cpu()
    ref = torch.zeros((N, H, W + S - 1, 64))
    for s in range(S):
        for co in range(CO):
            ref[:, :, s:s + W, s * 4 + co] = dzz[:, co]
    assert (got - ref).abs().max().item() < 2 ** -15
