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
