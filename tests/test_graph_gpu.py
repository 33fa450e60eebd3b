"""CUDA-graph capture of the training micro-batch (picklebot_b200.graph): replays must reproduce the eager
forward/backward bit for bit where the arithmetic is deterministic, compose with gradient accumulation, keep
following the optimizer's weight updates, and leave BatchNorm buffers as eager training would."""
import pytest
import torch
import torch.nn.functional as F

from _util import golden, rel_err, synthetic_checkpoint
from picklebot_b200 import synth

pytestmark = pytest.mark.gpu


def _build(model):
    import picklebot_b200 as pb
    g = golden(model)
    m = pb.valid_models[model](num_classes=g["num_classes"])
    m.initialize_weights()
    m.load_state_dict(synthetic_checkpoint(model))
    return m.cuda().train(), g


def _inputs(shape, nc, seed):
    clips = synth.synthetic_clips_u8(*shape, seed=seed).cuda()            # (B,T,H,W,3) uint8
    return clips.permute(0, 4, 1, 2, 3), synth.synthetic_labels(shape[0], nc, seed=seed + 1).cuda()


def _run_eager(m, opt, batches, autocast):
    opt.zero_grad(set_to_none=False)
    losses = []
    for x, y in batches:
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = F.cross_entropy(m(x), y)
        else:
            loss = F.cross_entropy(m(x.float() / 255.0), y)
        loss.backward()
        losses.append(float(loss.detach()))
    return losses, torch.cat([p.grad.flatten() for p in m.parameters()]).clone()


@pytest.mark.parametrize("model,autocast", [("MobileNetLarge3D", False), ("MobileNetSmall3D", False), ("MoViNetA2", False),
                                            ("MobileNetLarge3D", True)])
def test_graphed_step_matches_eager(model, autocast, monkeypatch):
    """Two passes of three accumulated micro-batches, the weights rescaled in between: replayed graph vs eager twin.
    fp32 storage: 1e-4 (only the order of atomic partial sums differs).  bf16 autocast: the kernels' atomics make
    even two eager runs differ, and at this tiny batch a single flipped bf16 rounding moves the loss by ~1e-2 (two
    eager twins are printed for reference), so the bf16 bar is a sanity bound, the fp32 one is the proof.
    (MoViNetA2 under bf16 at this size is chaotic -- two eager runs' gradients differ by ~100 % -- and is left out.)"""
    from picklebot_b200 import blocks
    from picklebot_b200.graph import GraphedTrainStep
    # Dropout3d noise off (mask of ones) so that the passes are comparable number for number
    monkeypatch.setattr(blocks, "draw_dropout3d_mask",
                        lambda B, C, p, dtype, device: torch.ones((B, C), dtype=torch.float32, device=device))
    (m_e, g), (m_e2, _), (m_g, _) = _build(model), _build(model), _build(model)
    if model == "MoViNetA2":
        for m in (m_e, m_e2, m_g):
            m.classifier[5].p = 0.0                                       # nn.Dropout before the last Linear
    shape, nc = g["train_shape"], g["num_classes"]
    batches = [_inputs(shape, nc, seed) for seed in (11, 22, 33)]
    opts = [torch.optim.SGD(m.parameters(), lr=0.05) for m in (m_e, m_e2, m_g)]
    if autocast:
        step = GraphedTrainStep(m_g, *batches[0])
    else:
        step = GraphedTrainStep(m_g, batches[0][0].float() / 255.0, batches[0][1], autocast_dtype=None)
    assert step.launches > 100
    assert 50 < step.launches_warm < step.launches                      # the warm graph skips the weight casts
    sd_e, sd_g = m_e.state_dict(), m_g.state_dict()
    for k in sd_e:                                                        # construction left no trace
        assert torch.equal(sd_e[k], sd_g[k]), k
    for it in range(2):
        l_e, g_e = _run_eager(m_e, opts[0], batches, autocast)
        l_e2, g_e2 = _run_eager(m_e2, opts[1], batches, autocast)
        opts[2].zero_grad(set_to_none=False)
        # gradient accumulation: only the first micro-batch after a weight change re-derives the shadow weights
        l_g = [float(step(x if autocast else x.float() / 255.0, y, weights_changed=(i == 0)))
               for i, (x, y) in enumerate(batches)]
        g_g = torch.cat([p.grad.flatten() for p in m_g.parameters()])
        noise_l = max(abs(a - b) for a, b in zip(l_e, l_e2))
        noise_g = rel_err(g_e2, g_e)
        d_l, d_g = max(abs(a - b) for a, b in zip(l_e, l_g)), rel_err(g_g, g_e)
        print(f"\n{model} pass {it}: eager-vs-eager loss {noise_l:.2e} grads {noise_g:.2e}; graph-vs-eager loss "
              f"{d_l:.2e} grads {d_g:.2e}")
        if autocast:
            assert d_l <= max(3 * noise_l, 3e-2) and d_g <= max(3 * noise_g, 0.3)
        else:
            # the proof: same arithmetic, only the atomics reorder partial sums (the eager twins bound that;
            # MoViNetA2's 26 train-mode BN layers over <= 128 samples amplify it to the 1e-3 level)
            # MobileNetLarge3D: with the synthetic checkpoint one block4.5 activation sits at u = 3.000000 +- 2e-6,
            # the discontinuity of Hardswish' (1.5 -> 1.0, same in torch); the 1e-6 forward round-off decides on which
            # side it falls, which moves the gradients by 0.7-1.2 % in about a third of all runs, eager or replayed.
            floor_l, floor_g = {"MoViNetA2": (5e-4, 5e-3), "MobileNetLarge3D": (2e-4, 2e-2)}.get(model, (2e-4, 2e-4))
            assert d_l <= max(3 * noise_l, floor_l) and d_g <= max(3 * noise_g, floor_g)
        if it == 0:
            # change the weights the way an optimizer does (in place, bumping the version): the replayed graph must
            # follow, i.e. re-cast its bf16 / transposed / block-diagonal shadow copies
            first_losses = l_g
            with torch.no_grad():
                for m in (m_e, m_e2, m_g):
                    for p in m.parameters():
                        if p.dim() > 1:                  # not a rescaling: BatchNorm makes the nets blind to those
                            p.view(-1).add_(p.view(-1).roll(1), alpha=0.3)
        else:
            assert max(abs(a - b) for a, b in zip(first_losses, l_g)) > 10 * max(d_l, 1e-4)    # it did follow
    sd_e, sd_g = m_e.state_dict(), m_g.state_dict()
    for k in sd_e:
        if k.endswith("num_batches_tracked"):
            assert int(sd_e[k]) == int(sd_g[k]) == 6, k
