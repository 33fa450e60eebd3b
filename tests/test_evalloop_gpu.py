"""estimate_loss (picklebot_b200.evalloop, the reference's train.py:123-153) on the CUDA modules: loss and accuracy
over a small validation set against the oracle evaluated the reference's way, batch by batch in fp32."""
import os

import pytest
import torch
import torch.nn.functional as F

from _util import GOLDEN
from oracle import picklebot_oracle as O
from picklebot_b200 import synth

pytestmark = pytest.mark.gpu


def test_estimate_loss_matches_reference_style_evaluation():
    import picklebot_b200 as pb
    from picklebot_b200 import loss as pbloss
    from picklebot_b200.evalloop import estimate_loss
    model = "MobileNetSmall3D"
    path = os.path.join(GOLDEN, f"{model}_trained.pt")
    if not os.path.exists(path):
        pytest.skip("trained fixture missing")
    fx = torch.load(path)
    state = {k: (v.float() if v.is_floating_point() else v) for k, v in fx["state"].items()}
    batches = []
    for i in range(3):
        clips, labels = synth.synthetic_task_clips_u8(8, 8, 64, 64, seed=4242 + i)
        if i == 1:
            labels = 1 - labels                                  # make one batch wrong on purpose: accuracy < 1
        batches.append((clips, labels.view(-1, 1)))              # the loader's (B,T,H,W,C) uint8 / (B,1) layout
    m = pb.valid_models[model](num_classes=2)
    m.load_state_dict(state)
    m = m.cuda().train()                                         # estimate_loss must switch to eval and back
    loss, acc = estimate_loss(m, batches, pbloss.cross_entropy_with_accuracy, "cuda", use_autocast=True)
    assert m.training
    # reference-style: per-batch CrossEntropyLoss().item() summed / len(loader), correct / samples (fp32 oracle)
    tot, correct, n = 0.0, 0, 0
    with torch.no_grad():
        for clips, labels in batches:
            logits = O.MODELS[model](state, synth.clips_to_features(clips, torch.float32).contiguous())
            tot += float(F.cross_entropy(logits, labels.view(-1)))
            correct += int((logits.argmax(1) == labels.view(-1)).sum())
            n += labels.shape[0]
    assert acc == pytest.approx(correct / n)
    assert 0.3 < acc < 0.9
    assert loss == pytest.approx(tot / len(batches), rel=2e-2, abs=2e-3)
    # a plain callable criterion takes the torch.max route of train.py:110-114
    loss2, acc2 = estimate_loss(m, batches, lambda o, y: F.cross_entropy(o.float(), y), "cuda", all_reduce=False)
    assert acc2 == pytest.approx(acc) and loss2 == pytest.approx(loss, rel=1e-3, abs=1e-4)
