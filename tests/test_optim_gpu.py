"""Multi-tensor AdamW (pb_adamw_step) against torch.optim.AdamW on identical parameters and gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [(1,), (7,), (16, 3, 3, 3, 3), (4096,), (4097,), (960, 160, 1, 1, 1), (1280, 960), (3,), (12289,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in SIZES]


@pytest.mark.parametrize("wd", [0.0, 5e-4, 0.1])
def test_adamw_matches_torch(wd):
    from picklebot_b200.optim import AdamW
    ours, ref = _params(1), _params(1)
    o1 = AdamW(ours, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    o2 = torch.optim.AdamW(ref, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    g = torch.Generator().manual_seed(2)
    for step in range(5):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 if step == 2 else 1.0)
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 3:                              # a new gradient tensor per parameter: the address tables must follow
            for a in ours:
                a.grad = a.grad.clone()
        o1.step()
        o2.step()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (step, tuple(a.shape), float((a - b).abs().max()))
    for a, b in zip(ours, ref):
        assert torch.allclose(o1.state[a]["exp_avg"], o2.state[b]["exp_avg"], rtol=2e-6, atol=1e-7)
        assert torch.allclose(o1.state[a]["exp_avg_sq"], o2.state[b]["exp_avg_sq"], rtol=2e-6, atol=1e-9)
        assert o1.state[a]["step"] == 5


def test_adamw_grad_scale_state_dict_and_errors():
    from picklebot_b200.optim import AdamW
    ours, ref = _params(3), _params(3)
    o1, o2 = AdamW(ours, lr=1e-3), torch.optim.AdamW(ref, lr=1e-3)
    for a, b in zip(ours, ref):
        b.grad = torch.randn_like(b)
        a.grad = b.grad * 1024.0                  # a loss-scaled gradient (GradScaler)
    o1.step(grad_scale=1.0 / 1024.0)
    o2.step()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
    # state_dict round trip into a fresh optimiser continues identically
    clone = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o3 = AdamW(clone, lr=1e-3)
    o3.load_state_dict(copy.deepcopy(o1.state_dict()))     # load_state_dict keeps references to same-device tensors
    for a, c in zip(ours, clone):
        c.grad = a.grad.clone()
    o1.step(); o3.step()
    for a, c in zip(ours, clone):
        assert torch.equal(a, c)
    # no CPU fallback
    cpu = [torch.nn.Parameter(torch.randn(4))]
    cpu[0].grad = torch.randn(4)
    with pytest.raises(RuntimeError):
        AdamW(cpu).step()


@pytest.mark.parametrize("B,NC", [(64, 2), (7, 2), (33, 10), (4, 1000), (1, 3)])
def test_cross_entropy_loss_gradient_and_accuracy(B, NC):
    """pb_ce_loss against F.cross_entropy / argmax accuracy (train.py:110-114, 214, 266-267)."""
    import torch.nn.functional as F
    from picklebot_b200.loss import cross_entropy_with_accuracy
    g = torch.Generator().manual_seed(B * 1000 + NC)
    logits = (torch.randn(B, NC, generator=g) * 3).cuda().requires_grad_(True)
    ref_logits = logits.detach().clone().requires_grad_(True)
    labels = torch.randint(0, NC, (B,), generator=g).cuda()
    loss, correct = cross_entropy_with_accuracy(logits, labels, scale=0.125)
    ref = F.cross_entropy(ref_logits, labels) * 0.125
    assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    (loss * 3.0).backward()
    (ref * 3.0).backward()
    assert torch.allclose(logits.grad, ref_logits.grad, rtol=1e-5, atol=1e-8)
    assert int(correct) == int((ref_logits.argmax(1) == labels).sum())
    assert correct.dtype == torch.int32 and not correct.requires_grad
