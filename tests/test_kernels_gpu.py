"""Per-kernel parity on the B200: every C-ABI entry point against the same op written with torch
(fp32 math on the GPU, TF32 off).  fp32 storage must agree to 1e-4 (norm-wise), bf16 storage to 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from _util import rel_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def tol(dt):
    return 1e-4 if dt == torch.float32 else 1e-2


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rnd(*shape, dt=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return ((torch.rand(*shape, generator=g) * 2 - 1) * scale).cuda().to(dt)


DW_CASES = [
    # B, C, T, H, W, k, s, p
    (2, 16, 4, 12, 12, (1, 3, 3), (1, 1, 1), (1, 1, 1)),      # mobilenet k3 s1: time padded
    (2, 64, 5, 13, 11, (1, 3, 3), (2, 2, 2), (1, 1, 1)),      # mobilenet k3 s2: time strided, odd sizes
    (2, 72, 4, 14, 14, (1, 5, 5), (2, 2, 2), (2, 2, 2)),      # k5 s2
    (1, 120, 3, 7, 7, (1, 5, 5), (1, 1, 1), (2, 2, 2)),       # k5 s1 small spatial
    (2, 40, 6, 9, 9, (3, 3, 3), (1, 1, 1), (1, 1, 1)),        # movinet 3x3x3
    (1, 96, 6, 10, 10, (3, 3, 3), (1, 2, 2), (1, 1, 1)),      # movinet 3x3x3 spatial stride
    (1, 240, 7, 6, 6, (5, 3, 3), (1, 2, 2), (2, 1, 1)),       # movinet 5x3x3
    (1, 960, 3, 7, 7, (1, 5, 5), (1, 1, 1), (2, 2, 2)),       # widest layer
    (2, 8, 1, 1, 1, (1, 3, 3), (1, 1, 1), (1, 1, 1)),         # degenerate 1x1x1 frame
    (2, 104, 3, 12, 26, (1, 5, 5), (1, 1, 1), (2, 2, 2)),      # 5x5 stride 1, ragged channel block (C = 104 = 64 + 40)
    (1, 184, 2, 30, 17, (1, 5, 5), (1, 1, 1), (2, 2, 2)),      # ... odd width, strips of 4
    (1, 128, 2, 40, 60, (1, 5, 5), (1, 1, 1), (2, 2, 2)),      # ... wide rows, several w-tiles
    # MobileNetLarge3D's own layer shapes, 2 clips
    (2, 16, 8, 112, 112, (1, 3, 3), (1, 1, 1), (1, 1, 1)),    # block2.0
    (2, 64, 10, 112, 112, (1, 3, 3), (2, 2, 2), (1, 1, 1)),   # block2.1
    (2, 72, 6, 56, 56, (1, 3, 3), (1, 1, 1), (1, 1, 1)),      # block2.2
    (2, 72, 8, 56, 56, (1, 5, 5), (2, 2, 2), (2, 2, 2)),      # block3.0
    (2, 120, 6, 28, 28, (1, 5, 5), (1, 1, 1), (2, 2, 2)),     # block3.1
    (2, 240, 14, 28, 28, (1, 3, 3), (2, 2, 2), (1, 1, 1)),    # block4.0
    (2, 184, 10, 14, 14, (1, 3, 3), (1, 1, 1), (1, 1, 1)),    # block4.2
    (2, 672, 16, 14, 14, (1, 3, 3), (1, 1, 1), (1, 1, 1)),    # block4.5
    (2, 672, 18, 14, 14, (1, 5, 5), (2, 2, 2), (2, 2, 2)),    # block5.0
    (2, 960, 11, 7, 7, (1, 5, 5), (1, 1, 1), (2, 2, 2)),      # block5.1
    # window / tile edges of the tensor-core stride-1 kernels (dwconv_mma.cu)
    (3, 64, 1, 13, 31, (1, 3, 3), (1, 1, 1), (1, 1, 1)),      # one frame in, odd sizes, 2 windows with a ragged second one
    (1, 200, 5, 29, 12, (1, 5, 5), (1, 1, 1), (2, 2, 2)),     # 12 columns = exactly one 16-wide chunk, odd height
    (2, 136, 3, 9, 61, (1, 3, 3), (1, 1, 1), (1, 1, 1)),      # 61 columns: 3 windows padded to 4, ragged channel block
    (1, 64, 2, 5, 113, (1, 5, 5), (1, 1, 1), (2, 2, 2)),      # 113 columns: 5 windows padded to 8
    (5, 72, 3, 7, 7, (1, 5, 5), (1, 1, 1), (2, 2, 2)),        # 7x7 planes, odd frame count (two frames per tile)
    # MoViNetA2's own (kT,3,3) layer shapes (movinet.py:99-136), 2 clips of 8 frames
    (2, 64, 8, 56, 56, (3, 3, 3), (1, 1, 1), (1, 1, 1)),      # block2.2
    (2, 96, 8, 56, 56, (3, 3, 3), (1, 2, 2), (1, 1, 1)),      # block3.0
    (2, 240, 8, 28, 28, (5, 3, 3), (1, 2, 2), (2, 1, 1)),     # block4.0
    (2, 240, 8, 14, 14, (5, 3, 3), (1, 1, 1), (2, 1, 1)),     # block5.0
    (2, 480, 8, 14, 14, (5, 3, 3), (1, 2, 2), (2, 1, 1)),     # block6.0
    (2, 480, 8, 7, 7, (3, 3, 3), (1, 1, 1), (1, 1, 1)),       # block6.5
]

# (case index) -> must be served by the TMA fast paths in all three directions (every MobileNet class shape)
def _is_mobilenet_class(k, s, p):
    return k[0] == 1 and k[1] == k[2] and k[1] in (3, 5) and s[0] == s[1] == s[2] and s[0] in (1, 2) \
        and p[0] == p[1] == p[2] == k[1] // 2


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", DW_CASES + [(64, 120, 3, 28, 28, (1, 5, 5), (1, 1, 1), (1, 2, 2)),
                                             (70, 672, 2, 14, 14, (1, 3, 3), (1, 1, 1), (1, 1, 1)),
                                             (5, 72, 6, 20, 20, (1, 5, 5), (2, 2, 2), (1, 2, 2))])
def test_dwconv_fwd_pool(case, dt):
    """pb_dwconv3d_fwd_pool: the squeeze-excite average pool accumulated by the depthwise kernel while it stores its
    outputs (strip kernels; other paths pool in a second pass) equals pooling the stored tensor."""
    from picklebot_b200 import ops
    B, C, T, H, W, k, s, p = case
    x = rnd(B, T, H, W, C, dt=dt, seed=1)
    w_tc = ops.dw_weight_tapmajor(rnd(C, 1, *k, seed=2, scale=0.5), dt)
    y0 = ops.dwconv_fwd(x, w_tc, k, s, p)
    y, pooled = ops.dwconv_fwd_pool(x, w_tc, k, s, p)
    assert rel_err(y.float(), y0.float()) < 1e-4          # the tensor-core kernel may have served y0: a few 1-ulp bf16 flips
    ref = y.float().mean(dim=(1, 2, 3))
    assert pooled.shape == (B, C) and rel_err(pooled, ref) < 1e-5
    assert rel_err(pooled, ops.pool_fwd(y, B, C)) < 1e-5



@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", DW_CASES)
@pytest.mark.parametrize("mma", ["default", "forced"])
def test_dwconv_fwd_dgrad_wgrad(case, dt, mma, monkeypatch):
    """mma = forced: PB_DW_MMA=1 sends every stride-1 (1,k,k) shape with >= 64 channels through the tensor-core
    kernels of dwconv_mma.cu (by default only the shapes where they measured faster take that path)."""
    from picklebot_b200 import ops
    if mma == "forced":
        if dt != torch.bfloat16 or case[6] != (1, 1, 1) or case[5][0] != 1 or case[1] < 64:
            pytest.skip("not a shape of the mma kernels")
        monkeypatch.setenv("PB_DW_MMA", "1")
    B, C, T, H, W, k, s, p = case
    from picklebot_b200 import _lib
    x = rnd(B, T, H, W, C, dt=dt, seed=1)
    w = rnd(C, 1, *k, seed=2, scale=0.5)
    w_tc = ops.dw_weight_tapmajor(w, dt)
    _lib.path_reset()
    y = ops.dwconv_fwd(x, w_tc, k, s, p)
    # torch reference in fp32 on the same (rounded) operands
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wr = w.to(dt).float().requires_grad_(True)
    yr = F.conv3d(xr, wr, None, s, p, 1, C)
    assert tuple(y.shape) == tuple(yr.permute(0, 2, 3, 4, 1).shape)
    assert rel_err(y.float(), yr.permute(0, 2, 3, 4, 1)) < tol(dt)
    dy = rnd(*y.shape, dt=dt, seed=3)
    yr.backward(dy.float().permute(0, 4, 1, 2, 3))
    dx = ops.dwconv_dgrad(dy, w_tc, x.shape, k, s, p)
    assert rel_err(dx.float(), xr.grad.permute(0, 2, 3, 4, 1)) < tol(dt)
    dw_tc = ops.dwconv_wgrad(x, dy, k, s, p)
    dw = ops.dw_weight_grad_from_tapmajor(dw_tc, w.shape)
    assert rel_err(dw, wr.grad) < (1e-4 if dt == torch.float32 else 2e-3)
    paths = _lib.path_counts()
    movinet3d = k[0] in (3, 5) and k[1:] == (3, 3) and s[0] == 1 and s[1] == s[2] and p == (k[0] // 2, 1, 1)
    if dt == torch.bfloat16 and _is_mobilenet_class(k, s, p) and (s[0] == 1 or min(H, W) >= 2):
        # production path: no silent fall-through to the one-pixel-per-thread kernels
        assert paths["dw_fwd_tma"] == 1 and paths["dw_dgrad_tma"] == 1 and paths["dw_wgrad_tma"] == 1, paths
    elif dt == torch.bfloat16 and movinet3d:
        # MoviNetBottleneck.conv (kT,3,3): TMA-tiled forward; its gradients still run on the general-shape kernels
        assert paths["dw_fwd_tma"] == 1 and paths["dw_dgrad_generic"] == 1 and paths["dw_wgrad_generic"] == 1, paths
    else:
        assert paths["dw_fwd_generic"] == 1 and paths["dw_dgrad_generic"] == 1 and paths["dw_wgrad_generic"] == 1, paths


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("kt", [1, 3, 5])
def test_stream_dwconv_chunking_invariance(kt, dt):
    """8 chunks through the stream kernel == one causal pass (F.pad left by kT-1, movinet.py:34-39)."""
    from picklebot_b200 import ops
    B, C, T, H, W = 2, 24, 12, 9, 9
    k, s, p = (kt, 3, 3), (1, 2, 2), (0, 1, 1)
    x = rnd(B, T, H, W, C, dt=dt, seed=4)
    w = rnd(C, 1, *k, seed=5, scale=0.5)
    w_tc = ops.dw_weight_tapmajor(w, dt)
    xr = F.pad(x.float().permute(0, 4, 1, 2, 3), (0, 0, 0, 0, kt - 1, 0))
    yr = F.conv3d(xr, w.to(dt).float(), None, s, p, 1, C).permute(0, 2, 3, 4, 1)
    outs, buf = [], None
    for chunk in (x[:, :5], x[:, 5:6], x[:, 6:]):     # ragged chunks, one shorter than kT-1
        y, buf = ops.stream_dwconv_fwd(chunk.contiguous(), buf, w_tc, k, s, p)
        outs.append(y)
    y = torch.cat(outs, 1)
    assert rel_err(y.float(), yr) < tol(dt)
    from picklebot_b200 import _lib
    _lib.path_reset()
    whole, _ = ops.stream_dwconv_fwd(x, None, w_tc, k, s, p)
    assert torch.equal(whole, y)                      # chunking is bit-exact
    paths = _lib.path_counts()
    if dt == torch.bfloat16:                          # TMA-tiled: (1,3,3) stateless kernel, (kT,3,3) stream-buffer kernel
        assert paths["dw_stream_tma"] == 1 and paths["dw_stream_generic"] == 0, paths


@pytest.mark.parametrize("kt", [3, 5])
def test_stream_dwconv_movinet_shapes(kt):
    """The streaming kernel on MoViNetA2 layer shapes: an 8-frame chunk with a random (kT-1)-frame history."""
    from picklebot_b200 import _lib, ops
    for (C, H, W, s) in ((64, 56, 56, 1), (96, 56, 56, 2), (240, 14, 14, 1), (480, 14, 14, 2), (480, 7, 7, 1)):
        B, T = 2, 8
        k, st, p = (kt, 3, 3), (1, s, s), (0, 1, 1)
        x = rnd(B, T, H, W, C, dt=torch.bfloat16, seed=4)
        hist = rnd(B, kt - 1, H, W, C, dt=torch.bfloat16, seed=6)
        w = rnd(C, 1, *k, seed=5, scale=0.5)
        w_tc = ops.dw_weight_tapmajor(w, torch.bfloat16)
        _lib.path_reset()
        y, buf = ops.stream_dwconv_fwd(x, hist, w_tc, k, st, p)
        assert _lib.path_counts()["dw_stream_tma"] == 1
        xr = torch.cat([hist, x], 1).float().permute(0, 4, 1, 2, 3)
        yr = F.conv3d(xr, w.to(torch.bfloat16).float(), None, st, p, 1, C).permute(0, 2, 3, 4, 1)
        assert rel_err(y.float(), yr) < 1e-2, (C, H, W, s)
        assert torch.equal(buf, torch.cat([hist, x], 1)[:, -(kt - 1):])


GEMM_CASES = [(1, 300, 16, 16), (1, 1000, 24, 72), (3, 131, 72, 40), (2, 257, 960, 160), (1, 64, 960, 1280),
              (1, 64, 1280, 2), (1, 5, 8, 13)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm_simt_fwd_dgrad_wgrad(case, dt):
    from picklebot_b200 import ops
    Bt, R, K, N = case
    if dt == torch.bfloat16 and (N % 8 or K % 8):
        pytest.skip("bf16 activations need channel counts that are multiples of 8")
    A = rnd(Bt * R, K, dt=dt, seed=1)
    W = rnd(N, K, seed=2, scale=0.3)
    bias = rnd(N, seed=3)
    gate = (rnd(Bt, K, seed=4).abs() + 0.1).contiguous()
    Wr = W.to(dt).float()
    Ar = A.float().view(Bt, R, K)
    # forward with gate + bias
    C = ops.gemm_simt(A, W, N, K, K, 1, bias=bias, ascale=gate, Bt=Bt)
    As = (Ar * gate[:, None, :]).to(dt).float()
    Cr = As @ Wr.t() + bias
    assert rel_err(C.float().view(Bt, R, N), Cr) < tol(dt)
    # epilogue scale/add
    cs, ca = rnd(Bt, N, seed=5), rnd(Bt, N, seed=6)
    C2 = ops.gemm_simt(A, W, N, K, K, 1, colscale=cs, coladd=ca, Bt=Bt)
    Cr2 = (Ar @ Wr.t()) * cs[:, None, :] + ca[:, None, :]
    assert rel_err(C2.float().view(Bt, R, N), Cr2) < tol(dt)
    # dgrad form: dA = dC x W
    dC = rnd(Bt * R, N, dt=dt, seed=7)
    dA = ops.gemm_simt(dC, W, K, N, 1, K)
    assert rel_err(dA.float(), dC.float() @ Wr) < tol(dt)
    # wgrad with gate and bias
    dW, db = ops.wgrad_simt(A, dC, K, N, ascale=gate, Bt=Bt, want_bias=True)
    dWr = torch.einsum("brn,brk->nk", dC.float().view(Bt, R, N), As)
    assert rel_err(dW, dWr) < (1e-4 if dt == torch.float32 else 2e-3)
    assert rel_err(db, dC.float().sum(0)) < (1e-4 if dt == torch.float32 else 2e-3)


ACTS = [("relu", F.relu), ("hswish", F.hardswish), ("lrelu", lambda t: F.leaky_relu(t, 0.01)), ("none", lambda t: t)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("act", ACTS, ids=[a[0] for a in ACTS])
@pytest.mark.parametrize("training", [True, False])
def test_bn_act_dropout_fwd_bwd(act, training, dt):
    from picklebot_b200 import blocks, ops
    name, fn = act
    B, R, C = 3, 77, 40
    z = (rnd(B, R, C, seed=1) * 2 + 0.5).to(dt).contiguous()
    gamma, beta = rnd(C, seed=2) + 1.5, rnd(C, seed=3)
    rm, rv = rnd(C, seed=4), rnd(C, seed=5).abs() + 0.5
    mask = (torch.empty(B, C).bernoulli_(0.8, generator=torch.Generator().manual_seed(6)) / 0.8).cuda()
    rm0, rv0 = rm.clone(), rv.clone()
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    out, st = blocks.bn_forward(z.view(-1, C), B, C, gamma, beta, rm, rv, nbt, training, 1e-5, 0.1,
                                ops.ACT_CODES[name], 0.01, mask)
    zr = z.float().permute(0, 2, 1).requires_grad_(True)            # (B,C,R)
    g_r, b_r = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = rm0.clone(), rv0.clone()
    ref = fn(F.batch_norm(zr, rm_r, rv_r, g_r, b_r, training, 0.1, 1e-5)) * mask[:, :, None]
    assert rel_err(out.float().view(B, R, C), ref.permute(0, 2, 1)) < tol(dt)
    if training:
        assert rel_err(rm, rm_r) < 1e-5 and rel_err(rv, rv_r) < 1e-5 and int(nbt) == 1
    dout = rnd(B, R, C, dt=dt, seed=7)
    ref.backward(dout.float().permute(0, 2, 1))
    dz, dg, db = ops.bn_act_bwd(dout.view(-1, C), False, z.view(-1, C), *st, mask, B, C, ops.ACT_CODES[name],
                                training, 0.01)
    assert rel_err(dz.float().view(B, R, C), zr.grad.permute(0, 2, 1)) < tol(dt)
    assert rel_err(dg, g_r.grad) < (1e-4 if dt == torch.float32 else 5e-3)
    assert rel_err(db, b_r.grad) < (1e-4 if dt == torch.float32 else 5e-3)


@pytest.mark.parametrize("dt", DTYPES)
def test_bn_bwd_broadcast_dout(dt):
    """Gradient of a global average pool broadcast over the rows (block6 -> classifier pool)."""
    from picklebot_b200 import blocks, ops
    B, R, C = 4, 33, 64
    z = rnd(B, R, C, dt=dt, seed=1)
    gamma, beta = rnd(C, seed=2) + 1.5, rnd(C, seed=3)
    _, st = blocks.bn_forward(z.view(-1, C), B, C, gamma, beta, None, None, None, True, 1e-5, 0.1,
                              ops.ACT_HSWISH, 0.0, None)
    dfeat = rnd(B, C, seed=4)
    zr = z.float().requires_grad_(True)
    g_r = gamma.clone().requires_grad_(True)
    a = F.hardswish(F.batch_norm(zr.permute(0, 2, 1), None, None, g_r, beta, True, 0.1, 1e-5))
    (a.mean(2) * dfeat).sum().backward()
    dz, dg, _ = ops.bn_act_bwd((dfeat / R).contiguous(), True, z.view(-1, C), *st, None, B, C, ops.ACT_HSWISH, True)
    assert rel_err(dz.float().view(B, R, C), zr.grad) < tol(dt)
    assert rel_err(dg, g_r.grad) < (1e-4 if dt == torch.float32 else 5e-3)


@pytest.mark.parametrize("dt", DTYPES)
def test_squeeze_excite_pieces(dt):
    from picklebot_b200 import ops
    B, R, C, Ch = 3, 50, 72, 18
    x = rnd(B, R, C, dt=dt, seed=1)
    W1, b1, W2, b2 = rnd(Ch, C, seed=2, scale=0.4), rnd(Ch, seed=3), rnd(C, Ch, seed=4, scale=0.8), rnd(C, seed=5)
    pooled = ops.pool_fwd(x, B, C)
    assert rel_err(pooled, x.float().mean(1)) < 1e-5
    hidden, gate = ops.se_fc_fwd(pooled, W1, b1, W2, b2)
    pr = pooled.clone().requires_grad_(True)
    W1r, b1r, W2r, b2r = [t.clone().requires_grad_(True) for t in (W1, b1, W2, b2)]
    hr = F.relu(pr @ W1r.t() + b1r)
    gr = F.hardsigmoid(hr @ W2r.t() + b2r)
    assert rel_err(hidden, hr) < 1e-5 and rel_err(gate, gr) < 1e-5
    y = ops.rowscale(x, gate, B, C)
    assert rel_err(y.float(), x.float() * gate[:, None, :]) < tol(dt)
    g = rnd(B, R, C, dt=dt, seed=6)
    dgate = ops.rowdot(g, x, B, C)
    assert rel_err(dgate, (g.float() * x.float()).sum(1)) < 1e-4
    gr.backward(dgate)
    dmean, dW1, db1, dW2, db2 = ops.se_fc_bwd(dgate, pooled, hidden, gate, W1, W2, 1.0 / R)
    assert rel_err(dmean, pr.grad / R) < 1e-4
    for mine, ref in ((dW1, W1r.grad), (db1, b1r.grad), (dW2, W2r.grad), (db2, b2r.grad)):
        assert rel_err(mine, ref) < 1e-4
    g2 = g.clone()
    ops.scale_add_(g2, gate, dmean, B, C)
    assert rel_err(g2.float(), g.float() * gate[:, None, :] + dmean[:, None, :]) < tol(dt)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("src", ["u8_ndhwc", "f32_ncdhw", "act_ndhwc"])
@pytest.mark.parametrize("conv", [((3, 3, 3), (2, 2, 2), (1, 1, 1), True), ((1, 3, 3), (1, 2, 2), (0, 1, 1), False)],
                         ids=["mobilenet", "movinet"])
def test_stem_fwd_wgrad(conv, src, dt):
    from picklebot_b200 import ops
    k, s, p, has_bias = conv
    B, T, H, W = 2, 5, 17, 15
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (B, T, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    if src == "u8_ndhwc":
        x = u8.permute(0, 4, 1, 2, 3)                       # raw clip, /255 fused in the kernel
        xr = (u8.permute(0, 4, 1, 2, 3).to(dt) / 255).float()
    elif src == "f32_ncdhw":
        x = (u8.permute(0, 4, 1, 2, 3).float() / 255).contiguous()
        xr = x.to(dt).float()
    else:
        x = u8.permute(0, 4, 1, 2, 3).to(dt) / 255           # exactly train.py:106
        xr = x.float()
    w = rnd(16, 3, *k, seed=2, scale=0.3)
    bias = rnd(16, seed=3) if has_bias else None
    y = ops.stem_fwd(x, w, bias, k, s, p, dt)
    wr = w.to(dt).float().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if has_bias else None
    yr = F.conv3d(xr, wr, br, s, p)
    assert rel_err(y.float(), yr.permute(0, 2, 3, 4, 1)) < tol(dt)
    dy = rnd(*y.shape, dt=dt, seed=4)
    yr.backward(dy.float().permute(0, 4, 1, 2, 3))
    dw, db = ops.stem_wgrad(x, dy, w.shape, k, s, p, has_bias)
    assert rel_err(dw, wr.grad) < (1e-4 if dt == torch.float32 else 2e-3)
    if has_bias:
        assert rel_err(db, br.grad) < (1e-4 if dt == torch.float32 else 2e-3)


@pytest.mark.parametrize("conv", [((3, 3, 3), (2, 2, 2), (1, 1, 1), True), ((1, 3, 3), (1, 2, 2), (0, 1, 1), False),
                                  ((3, 3, 3), (1, 1, 1), (1, 1, 1), True)], ids=["mobilenet", "movinet", "stride1"])
@pytest.mark.parametrize("shape", [(2, 6, 64, 64), (3, 5, 44, 48), (1, 4, 224, 224), (2, 3, 30, 16)],
                         ids=["64x64", "ragged_row_groups", "224x224", "narrow"])
def test_stem_tma_path(conv, shape, monkeypatch):
    """uint8 clips whose rows are 16-byte multiples go through the TMA-staged tcgen05 stem (stem_tc.cu, round 2):
    exact integer inputs, fp16 weights, /255 behind the accumulator.  Checked against fp32 convolution of the exact
    x/255 with the fp16-rounded weights (forward: only the bf16 output rounding is left) and, for the weight
    gradient, with the same bf16 upstream gradient (nothing is rounded before the products).  The gather kernels
    (PB_STEM_GATHER=1) must agree with it to bf16 input-rounding accuracy."""
    from picklebot_b200 import _lib, ops
    k, s, p, has_bias = conv
    B, T, H, W = shape
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (B, T, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    x = u8.permute(0, 4, 1, 2, 3)
    w = rnd(16, 3, *k, seed=2, scale=0.3)
    bias = rnd(16, seed=3) if has_bias else None
    _lib.path_reset()
    y = ops.stem_fwd(x, w, bias, k, s, p, torch.bfloat16)
    xr = x.float() / 255
    wr = w.half().float().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if has_bias else None
    yr = F.conv3d(xr, wr, br, s, p)
    assert rel_err(y.float(), yr.permute(0, 2, 3, 4, 1)) < 3e-3
    dy = rnd(*y.shape, dt=torch.bfloat16, seed=4)
    yr.backward(dy.float().permute(0, 4, 1, 2, 3))
    dw, db = ops.stem_wgrad(x, dy, w.shape, k, s, p, has_bias)
    tma = 2 if y.shape[3] <= 128 else 0                      # wider output rows stay on the gather kernels
    # TMA path: exact inputs, what is left is the tensor core's fp32 accumulation over ~1e5 pixels per accumulator;
    # gather path: x/255 rounded to bf16 first
    assert rel_err(dw, wr.grad) < (5e-4 if tma else 3e-3)
    if has_bias:
        assert rel_err(db, br.grad) < 1e-4
    c = _lib.path_counts()
    assert c["stem_tma"] == tma and c["stem_tc"] == 2 and c["stem_simt"] == 0, c
    monkeypatch.setenv("PB_STEM_GATHER", "1")
    y2 = ops.stem_fwd(x, w, bias, k, s, p, torch.bfloat16)
    dw2, _ = ops.stem_wgrad(x, dy, w.shape, k, s, p, has_bias)
    assert _lib.path_counts()["stem_tma"] == tma
    assert rel_err(y2.float(), y.float()) < 1e-2 and rel_err(dw2, dw) < 2e-3


def test_errors_are_loud():
    from picklebot_b200 import _lib, ops
    x = torch.zeros(1, 2, 2, 2, 12, device="cuda")           # C=12 is not a multiple of 8
    w_tc = torch.zeros(9, 12, device="cuda")
    with pytest.raises(_lib.PicklebotKernelError):
        ops.dwconv_fwd(x, w_tc, (1, 3, 3), (1, 1, 1), (1, 1, 1))
    with pytest.raises(RuntimeError):
        ops.colstats(torch.zeros(4, 8), 8)                    # CPU tensor
    assert _lib.lib().pb_device_check() == 0


@pytest.mark.parametrize("case", [(64, 1280, 960), (64, 2, 1280), (5, 100, 72), (64, 2048, 640), (1, 8, 2560), (9, 33, 8)])
def test_fc_fwd_dgrad(case):
    """Classifier-head nn.Linear layers (mobilenet.py:184-190, movinet.py:146-154): fp32, B <= a few dozen."""
    from picklebot_b200 import ops
    B, N, K = case
    X, W, b = rnd(B, K, seed=1), rnd(N, K, seed=2, scale=0.1), rnd(N, seed=3)
    assert rel_err(ops.fc_fwd(X, W, b), X @ W.t() + b) < 1e-5
    assert rel_err(ops.fc_fwd(X, W), X @ W.t()) < 1e-5
    if N <= 2560 - 8 * 33:
        dY = rnd(B, N, seed=4)
        assert rel_err(ops.fc_dgrad(dY, W, 0.25), 0.25 * (dY @ W)) < 1e-5


def test_fc_rejects_oversize():
    from picklebot_b200 import ops
    with pytest.raises(RuntimeError):
        ops.fc_fwd(rnd(2, 4096), rnd(8, 4096))
