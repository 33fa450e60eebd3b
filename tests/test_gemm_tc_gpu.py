"""tcgen05/TMEM/TMA GEMM against fp32 torch matmul on the bf16-rounded operands (B200 only)."""
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return ((torch.rand(*shape, generator=g) * 2 - 1) * scale).cuda()


# (Bt, R, K, N): pointwise layers of the three models, ragged row counts, K tails, N > 256
CASES = [(1, 128, 64, 16), (1, 1000, 16, 16), (1, 5000, 16, 64), (1, 777, 24, 72), (2, 931, 960, 160),
         (3, 200, 672, 112), (1, 3000, 160, 960), (1, 1234, 112, 672), (4, 37, 72, 40), (1, 4096, 184, 80),
         (2, 300, 80, 184), (1, 50000, 40, 240), (1, 129, 8, 8), (64, 735, 960, 160),
         # enough row tiles for the n-tile-resident schedule (a CTA keeps one of the 2-4 column tiles)
         (1, 40000, 160, 960), (1, 30011, 112, 672), (3, 9000, 80, 480)]


@pytest.mark.parametrize("case", CASES)
def test_gemm_tc_plain_and_epilogue(case):
    from picklebot_b200 import gemm_tc
    Bt, R, K, N = case
    A = rnd(Bt * R, K, seed=1).bfloat16()
    W = rnd(N, K, seed=2, scale=0.3).bfloat16()
    C = gemm_tc.gemm(A, W, N, K)
    ref = A.float() @ W.float().t()
    assert rel_err(C.float(), ref) < 6e-3
    bias, cs, ca = rnd(N, seed=3), rnd(Bt, N, seed=4), rnd(Bt, N, seed=5)
    C2 = gemm_tc.gemm(A, W, N, K, Bw=1, Bt=Bt, bias=bias, colscale=cs, coladd=ca)
    ref2 = ((ref + bias).view(Bt, R, N) * cs[:, None, :] + ca[:, None, :]).view(-1, N)
    assert rel_err(C2.float(), ref2) < 6e-3


@pytest.mark.parametrize("case", [(2, 931, 960, 160), (3, 200, 672, 112), (4, 37, 72, 40), (64, 49, 576, 96)])
def test_gemm_tc_per_sample_weights(case):
    """squeeze-excite gate folded into per-sample weights (pb_fold_gate_bf16)."""
    from picklebot_b200 import gemm_tc, ops
    Bt, R, K, N = case
    A = rnd(Bt * R, K, seed=1).bfloat16()
    W = rnd(N, K, seed=2, scale=0.3)
    gate = rnd(Bt, K, seed=3).abs() + 0.1
    Wb = ops.fold_gate(W, gate)
    assert rel_err(Wb.float(), W[None] * gate[:, None, :]) < 4e-3
    C = gemm_tc.gemm(A, Wb, N, K, Bw=Bt, Bt=Bt)
    ref = torch.einsum("brk,bnk->brn", A.float().view(Bt, R, K), Wb.float()).reshape(-1, N)
    assert rel_err(C.float(), ref) < 6e-3


@pytest.mark.parametrize("case", [(2, 931, 160, 960), (3, 3528, 112, 672), (64, 49, 96, 576), (5, 300, 40, 120),
                                  (2, 1000, 40, 72), (1, 50, 24, 64), (400, 130, 80, 480),
                                  # per-sample weights RESIDENT per (sample, column tile), swapped several times per CTA
                                  (8, 2000, 40, 120), (70, 700, 112, 480), (64, 931, 160, 960), (9, 5000, 24, 72)])
def test_gemm_tc_se_input_gradient(case, monkeypatch):
    """dy2 = (dz W) * gate + dmean of a squeeze-excite block (blocks.se_pw2_backward): the gate folded into the rows
    of per-sample transposed weights (pb_fold_gate_t_bf16), dmean pre-loaded into the TMEM accumulator by the
    epilogue warps (coladd alone) -- against fp32 math and against the fp32 epilogue-vector route it replaces.
    The last case has more tiles than 2 x 148, so every accumulator stage is re-loaded several times."""
    from picklebot_b200 import gemm_tc, ops
    Bt, R, Cout, Cexp = case
    dz = rnd(Bt * R, Cout, seed=1).bfloat16()
    W = rnd(Cout, Cexp, seed=2, scale=0.3)                   # the layer's weight [N = Cout][K = Cexp]
    gate = rnd(Bt, Cexp, seed=3).abs() + 0.1
    dmean = rnd(Bt, Cexp, seed=4, scale=2.0)
    Wtg = ops.fold_gate_t(W, gate)
    assert torch.equal(ops.fold_rows(W.t().contiguous(), gate), Wtg)      # the coalesced variant blocks.py uses
    assert Wtg.shape == (Bt, Cexp, Cout)
    assert rel_err(Wtg.float(), W.t()[None] * gate[:, :, None]) < 4e-3
    dy2 = gemm_tc.gemm(dz, Wtg, Cexp, Cout, Bw=Bt, Bt=Bt, coladd=dmean)
    ref = (torch.einsum("bro,oe->bre", dz.float().view(Bt, R, Cout), W) * gate[:, None, :] + dmean[:, None, :]).reshape(-1, Cexp)
    assert rel_err(dy2.float(), ref) < 6e-3
    old = gemm_tc.gemm(dz, W.t().contiguous().bfloat16(), Cexp, Cout, Bw=1, Bt=Bt, colscale=gate, coladd=dmean)
    assert rel_err(dy2.float(), old.float()) < 6e-3
    monkeypatch.setenv("PB_GEMM_NO_TINIT", "1")              # same call through the fp32 epilogue
    alt = gemm_tc.gemm(dz, Wtg, Cexp, Cout, Bw=Bt, Bt=Bt, coladd=dmean)
    assert rel_err(dy2.float(), alt.float()) < 3e-3


@pytest.mark.parametrize("case", [(1, 5000, 96, 24, "hswish", False), (4, 800, 672, 112, "hswish", True),
                                  (1, 3000, 64, 24, "relu", False), (64, 49, 960, 160, "hswish", True),
                                  (1, 777, 240, 80, "none", False), (3, 130, 72, 40, "relu", True)])
def test_gemm_tc_folded_bn_act(case):
    """Inference form of conv -> BatchNorm(eval) -> activation: scale (and the squeeze-excite gate) folded into bf16
    weights by pb_fold_scaled_bf16, shift pre-loaded into the accumulator, activation in the epilogue
    (pb_pw_gemm_tc_act) -- against the unfused fp32 math."""
    import torch.nn.functional as F
    from picklebot_b200 import gemm_tc, ops
    Bt, R, K, N, act, gated = case
    A = rnd(Bt * R, K, seed=1).bfloat16()
    W = rnd(N, K, seed=2, scale=0.3)
    gate = (rnd(Bt, K, seed=3).abs() + 0.1) if gated else None
    scale, shift = rnd(N, seed=4).abs() + 0.5, rnd(N, seed=5)
    Wb = ops.fold_scaled(W, gate, scale)
    full = W[None] * scale[None, :, None] * (gate[:, None, :] if gated else 1.0)
    assert Wb.shape == ((Bt if gated else 1), N, K) and rel_err(Wb.float(), full) < 4e-3
    C = gemm_tc.gemm(A, Wb if gated else Wb.view(N, K), N, K, Bw=Bt if gated else 1, Bt=Bt if gated else 1,
                     bias=shift, act=ops.ACT_CODES[act])
    pre = torch.einsum("brk,bnk->brn", A.float().view(Bt, R, K), Wb.float().expand(Bt, N, K)).reshape(-1, N) + shift
    ref = {"hswish": F.hardswish, "relu": F.relu, "none": lambda t: t}[act](pre)
    assert rel_err(C.float(), ref) < 6e-3


@pytest.mark.parametrize("case", [(4000, 16, 16, 4), (4000, 16, 64, 4), (1002, 24, 72, 2), (6000, 32, 96, 2)])
def test_gemm_tc_row_folded(case):
    """K <= 32 layers run as X'[rows/F][F*K] against diag(W, ..., W) (pb_block_diag_bf16): same bytes out."""
    from picklebot_b200 import gemm_tc, ops
    rows, K, N, F = case
    A = rnd(rows, K, seed=1).bfloat16()
    W = rnd(N, K, seed=2, scale=0.3).bfloat16()
    Wf = ops.block_diag(W, F)
    assert torch.equal(Wf.float(), torch.block_diag(*([W.float()] * F)))
    C = gemm_tc.gemm(A, Wf, N * F, K * F).view(-1, N)
    # zero blocks add exact zeros, but the MMA groups the K terms differently: agreement to bf16 round-off
    assert rel_err(C.float(), gemm_tc.gemm(A, W, N, K).float()) < 3e-3
    assert rel_err(C.float(), A.float() @ W.float().t()) < 6e-3


@pytest.mark.parametrize("case", [(1, 5000, 16, 16, 4), (1, 3001, 64, 24, 1), (2, 931, 960, 160, 1), (1, 4444, 72, 40, 1),
                                  (3, 777, 480, 112, 1), (1, 2000, 240, 80, 1), (1, 9000, 16, 64, 4), (64, 49, 576, 96, 1),
                                  (1, 1000, 144, 256, 1), (1, 130, 8, 8, 1)])
def test_gemm_tc_fused_bn_statistics(case):
    """pb_pw_gemm_tc(stats=...): the epilogue's column sums equal a statistics pass over the stored outputs."""
    from picklebot_b200 import gemm_tc, ops
    Bt, R, K, N, F = case
    A = rnd(Bt * R, K, seed=1).bfloat16()
    W = rnd(N, K, seed=2, scale=0.3).bfloat16()
    if F > 1:
        C, sums = gemm_tc.gemm(A, ops.block_diag(W, F), N * F, K * F, stat_mod=N)
        C = C.view(-1, N)
    else:
        C, sums = gemm_tc.gemm(A, W, N, K, Bw=1, Bt=Bt, stat_mod=N)
    assert rel_err(C.float(), A.float() @ W.float().t()) < 6e-3
    got = sums.sum(0)
    ref = torch.stack([C.double().sum(0), (C.double() ** 2).sum(0)])
    assert (got - ref).abs().max() / ref.abs().max() < 1e-5
    assert rel_err(got[1].float(), ref[1].float()) < 1e-5


def test_gemm_tc_matches_simt_bitwise_scale():
    """Same operands through the CUDA-core kernel: both accumulate in fp32, so they agree to bf16 round-off."""
    from picklebot_b200 import gemm_tc, ops
    R, K, N = 2000, 120, 40
    A = rnd(R, K, seed=1).bfloat16()
    W = rnd(N, K, seed=2, scale=0.3)
    a = gemm_tc.gemm(A, W.bfloat16(), N, K)
    b = ops.gemm_simt(A, W, N, K, K, 1)
    assert rel_err(a.float(), b.float()) < 3e-3


# (Bt, R, K, N): K = input channels of the layer, N = its output channels
WGRAD_CASES = [(1, 4096, 16, 16), (1, 100000, 16, 64), (1, 7000, 24, 72), (1, 50000, 72, 24), (2, 931, 960, 160),
               (1, 59584, 160, 960), (3, 777, 672, 112), (1, 3000, 112, 672), (4, 100, 72, 40), (1, 12544, 80, 184),
               (64, 735, 960, 160), (1, 65, 8, 8), (2, 5000, 144, 576), (1, 20000, 576, 144), (1, 6000, 40, 240),
               (1, 30000, 40, 120), (1, 9000, 120, 40), (1, 5000, 88, 24), (1, 3000, 96, 40), (2, 2000, 48, 144),
               (1, 2500, 32, 32), (1, 2500, 64, 64), (1, 2500, 128, 128), (1, 2500, 104, 56),
               (3, 1000, 16, 24), (2, 998, 24, 32), (5, 444, 32, 8)]   # row-folded plans


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_tc(case):
    from picklebot_b200 import gemm_tc
    Bt, R, K, N = case
    A = rnd(Bt * R, K, seed=1).bfloat16()
    dC = rnd(Bt * R, N, seed=2).bfloat16()
    dW, _ = gemm_tc.wgrad(A, dC, K, N)
    ref = dC.float().t() @ A.float()
    assert rel_err(dW, ref) < 1e-4
    # squeeze-excite form: per-sample gate and the gate gradient from the same per-sample products
    gate = rnd(Bt, K, seed=3).abs() + 0.1
    W = rnd(N, K, seed=4, scale=0.3)
    dW2, dgate = gemm_tc.wgrad(A, dC, K, N, gate=gate, W=W, Bt=Bt, want_dgate=True)
    P = torch.einsum("brn,brk->bnk", dC.float().view(Bt, R, N), A.float().view(Bt, R, K))
    assert rel_err(dW2, (P * gate[:, None, :]).sum(0)) < 1e-4
    assert rel_err(dgate, (P * W[None]).sum(1)) < 1e-4
