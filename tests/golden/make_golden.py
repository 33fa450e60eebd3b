"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE modules.

Run in the build container only (needs /root/reference, which never travels to the GPU box):

    python tests/golden/make_golden.py

Writes, per model M in {MobileNetLarge3D, MobileNetSmall3D, MoViNetA2}:
  M_bnstats.pt   calibrated BatchNorm running statistics (the one data-dependent part of the
                 synthetic checkpoint; weights come from picklebot_b200.synth, seed 1234)
  M_golden.pt    reference outputs: eval logits at full and reduced size, block taps digests, one
                 train step (logits, loss, per-parameter gradient digests, updated running stats)
and statedict_keys.json (key -> shape for all three models).

The reference has no tests or golden vectors of its own; these files are what pins the oracle.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from picklebot_b200 import synth  # noqa: E402

import mobilenet as ref_mobilenet  # noqa: E402  (reference)
import movinet as ref_movinet  # noqa: E402  (reference)

torch.set_num_threads(os.cpu_count() or 1)

CASES = {
    # name: (ctor, num_classes, full-size (B,T,H,W), small (B,T,H,W), train (B,T,H,W))
    "MobileNetLarge3D": (ref_mobilenet.MobileNetLarge3D, 2, (4, 16, 224, 224), (3, 8, 64, 64), (4, 8, 64, 64)),
    "MobileNetSmall3D": (ref_mobilenet.MobileNetSmall3D, 2, (4, 16, 224, 224), (3, 8, 64, 64), (4, 8, 64, 64)),
    "MoViNetA2": (ref_movinet.MoViNetA2, 13, (2, 8, 224, 224), (3, 8, 64, 64), (4, 8, 64, 64)),
}


def probe_vector(n: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1


def grad_digest(name: str, g: torch.Tensor):
    """(L2 norm, projection on a fixed pseudo-random vector, first 8 values)."""
    flat = g.detach().double().flatten()
    pv = probe_vector(flat.numel(), (hash_name(name) % 100000) + 11)
    return float(flat.norm()), float((flat * pv).sum()), flat[:8].float().clone()


def hash_name(name: str) -> int:
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def features(shape, seed):
    B, T, H, W = shape
    clips = synth.synthetic_clips_u8(B, T, H, W, seed=seed)
    return synth.clips_to_features(clips, torch.float32).contiguous()


def calibrate(model, shape):
    """Cumulative-average BN statistics over two train-mode passes (dropout off)."""
    import torch.nn as nn
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm3d, nn.BatchNorm1d)):
            m.momentum = None
            m.reset_running_stats()
    model.train()
    for m in model.modules():
        if isinstance(m, (nn.Dropout3d, nn.Dropout)):
            m.eval()
    with torch.no_grad():
        for s in (100, 101):
            model(features(shape, s))
    stats = {k: v.clone() for k, v in model.state_dict().items()
             if k.endswith("running_mean") or k.endswith("running_var")}
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm3d, nn.BatchNorm1d)):
            m.momentum = 0.1
    return stats


def main():
    keys = {}
    for name, (ctor, nc, full, small, train_shape) in CASES.items():
        print("==", name)
        torch.manual_seed(synth.SEED_WEIGHTS)
        model = ctor(num_classes=nc)
        model.initialize_weights()
        template = model.state_dict()
        keys[name] = {k: list(v.shape) for k, v in template.items()}
        sd0 = synth.synthetic_state_dict(template)
        model.load_state_dict(sd0)
        stats = calibrate(model, full)
        torch.save(stats, os.path.join(HERE, f"{name}_bnstats.pt"))
        sd = synth.synthetic_state_dict(template, bn_stats=stats)
        model.load_state_dict(sd)
        gold = {"num_classes": nc, "full_shape": full, "small_shape": small, "train_shape": train_shape,
                "digest": synth.state_dict_digest(sd)}

        model.eval()
        with torch.no_grad():
            gold["eval_full_logits"] = model(features(full, synth.SEED_DATA)).clone()
            gold["eval_small_logits"] = model(features(small, synth.SEED_DATA)).clone()
        print(" eval full logits\n", gold["eval_full_logits"])
        print(" cross-sample logit std", gold["eval_full_logits"].std(0))

        # one training step: CE loss, dropout drawn from the global RNG with seed 7
        model.train()
        x = features(train_shape, synth.SEED_DATA)
        labels = synth.synthetic_labels(train_shape[0], nc)
        torch.manual_seed(synth.SEED_DROPOUT)
        logits = model(x)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        model.zero_grad()
        loss.backward()
        gold["train_logits"] = logits.detach().clone()
        gold["train_loss"] = loss.detach().clone()
        gold["train_grads"] = {k: grad_digest(k, p.grad) for k, p in model.named_parameters()}
        after = model.state_dict()
        gold["train_running"] = {k: after[k].clone() for k in after
                                 if k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
        print(" train loss", float(loss), "logits", logits[0].tolist()[:4])
        torch.save(gold, os.path.join(HERE, f"{name}_golden.pt"))

    # CausalConv3d (movinet.py:7-39): the specification of the streaming semantics
    torch.manual_seed(5)
    cc = {}
    for kt in (1, 2, 3, 5):
        mod = ref_movinet.CausalConv3d(8, 8, (kt, 3, 3), stride=(1, 1, 1), padding=(0, 1, 1), groups=8, bias=False)
        g = torch.Generator().manual_seed(50 + kt)
        w = torch.rand(mod.conv3d.weight.shape, generator=g) - 0.5
        x = torch.rand((2, 8, 6, 5, 5), generator=g)
        with torch.no_grad():
            mod.conv3d.weight.copy_(w)
            y = mod(x)
        cc[kt] = {"w": w, "x": x, "y": y.clone()}
    torch.save(cc, os.path.join(HERE, "causalconv3d_golden.pt"))

    with open(os.path.join(HERE, "statedict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()
