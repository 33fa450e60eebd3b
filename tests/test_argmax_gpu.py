"""Ball/strike calls on a checkpoint that actually decides.

tests/golden/make_trained.py trained the synthetic checkpoints with the REFERENCE modules on a separable two-class
task and stored the reference's own eval logits of 96 held-out clips.  Both classes are predicted and the margins
are large against bf16 noise, so "argmax identical" is a real statement here (the seeded checkpoints of the other
tests predict one class for every clip).

MobileNetLarge3D and MobileNetSmall3D only: MoViNetA2 did not get off ln 2 within 450 end-to-end CPU steps, and the
pooled features of its synthetic checkpoint are effectively rank one on these clips (a closed-form fit of the head
reaches 55 %), so no two-sided MoViNet fixture exists; its parity rests on the golden / train-step / stream tests."""
import os

import pytest
import torch

from _util import GOLDEN, MODEL_NAMES, rel_err
from oracle import picklebot_oracle as O
from picklebot_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _load(model):
    path = os.path.join(GOLDEN, f"{model}_trained.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    fx = torch.load(path)
    state = {k: (v.float() if v.is_floating_point() else v) for k, v in fx["state"].items()}
    return fx, state


@pytest.mark.parametrize("model", ["MobileNetLarge3D", "MobileNetSmall3D"])
def test_argmax_identical_on_trained_checkpoint(model):
    import picklebot_b200 as pb
    fx, state = _load(model)
    n, T, H, W = fx["eval_shape"]
    clips, labels = synth.synthetic_task_clips_u8(n, T, H, W, seed=fx["eval_seed"])
    assert torch.equal(labels, fx["eval_labels"])
    ref = fx["eval_logits"]                                   # the reference module's fp32 eval logits
    ref_pred = ref.argmax(1)
    counts = torch.bincount(ref_pred, minlength=2)
    assert counts.min() >= n // 4, f"fixture is one-sided: {counts.tolist()}"
    m = pb.valid_models[model](num_classes=2)
    m.load_state_dict(state)                                  # strict: the reference's key layout
    m = m.cuda().eval()
    x = clips.cuda().permute(0, 4, 1, 2, 3)                   # raw uint8 clips, /255 fused into the stem
    with torch.no_grad():
        fp32 = m(synth.clips_to_features(clips.cuda(), torch.float32))
        _lib.path_reset()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            bf16 = m(x)
        paths = _lib.path_counts()
        sd = O.clone_state(state, device="cuda")
        with torch.autocast("cuda", dtype=torch.bfloat16):
            torch_bf16 = O.MODELS[model](sd, synth.clips_to_features(clips.cuda(), torch.bfloat16)).float()
    assert paths["gemm_tc"] > 0 and paths["gemm_simt"] <= 2, paths          # only the tiny FC head may use CUDA cores
    if model != "MoViNetA2":
        assert paths["dw_fwd_generic"] == 0, paths
    # fp32 path: the oracle-level 1e-4 bar and identical calls
    assert rel_err(fp32.cpu(), ref) < 1e-4
    assert torch.equal(fp32.argmax(1).cpu(), ref_pred)
    # bf16 production path: margin gate, then identical calls
    err = (bf16.float().cpu() - ref).abs().max()
    margin = (ref[:, 1] - ref[:, 0]).abs()
    decisive = margin > 10 * err
    print(f"\n{model}: calls {counts.tolist()}, bf16 max abs logit error {float(err):.3e} (torch-autocast "
          f"{float((torch_bf16.cpu() - ref).abs().max()):.3e}), rel {rel_err(bf16.float().cpu(), ref):.2e}, "
          f"margin min {float(margin.min()):.3f} median {float(margin.median()):.3f}, decisive clips {int(decisive.sum())}/{n}")
    assert decisive.float().mean() >= 0.9, "margins are not large against the bf16 error: the check would be vacuous"
    assert torch.equal(bf16.argmax(1).cpu()[decisive], ref_pred[decisive])
    assert torch.bincount(ref_pred[decisive], minlength=2).min() >= n // 5
    assert rel_err(bf16.float().cpu(), ref) < 2e-2
