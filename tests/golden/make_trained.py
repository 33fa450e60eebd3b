"""Train small synthetic checkpoints with the REFERENCE modules so that the argmax parity test is not vacuous.

Run in the build container only (needs /root/reference):

    python tests/golden/make_trained.py

The reference's weights/*.pth are absent (SURVEY.md finding 1) and the seeded synthetic checkpoint predicts one
class for every clip, so "argmax identical" proved nothing.  This script takes the synthetic checkpoint, runs a few
hundred SGD steps of the reference model (train mode: batch statistics, Dropout3d) on the two-class toy task of
``picklebot_b200.synth.synthetic_task_clips_u8`` (reduced clips, 8x64x64) and stores

  M_trained.pt   {"state": state_dict with every float tensor rounded to bf16 and stored as bf16 (half the bytes;
                  the checkpoint IS the rounded one), "eval_shape", "eval_seed", "eval_logits" (reference, fp32,
                  eval mode, on the stored checkpoint), "eval_labels", "train_log"}

for M in MobileNetLarge3D, MobileNetSmall3D (num_classes=2).  tests/test_argmax_gpu.py replays them.  (The MoViNetA2
branch is kept for the record: neither end-to-end steps nor a closed-form head fit give a two-sided fixture, see
tests/test_argmax_gpu.py.)
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from picklebot_b200 import synth  # noqa: E402

import mobilenet as ref_mobilenet  # noqa: E402  (reference)
import movinet as ref_movinet  # noqa: E402  (reference)

torch.set_num_threads(os.cpu_count() or 1)

CLIP = (8, 64, 64)
EVAL_CLIPS, EVAL_SEED = 96, 900001
CASES = {
    "MobileNetLarge3D": (ref_mobilenet.MobileNetLarge3D, 300, 0.002),
    "MobileNetSmall3D": (ref_mobilenet.MobileNetSmall3D, 300, 0.02),
    "MoViNetA2": (ref_movinet.MoViNetA2, 450, 0.004),
}


def feats(clips):
    return synth.clips_to_features(clips, torch.float32).contiguous()


def train_movinet_head(model):
    """MoViNetA2 does not get anywhere on the toy task in a few hundred end-to-end steps on CPU (26 blocks, loss stuck
    near ln 2), so only its classifier head is trained: the backbone (calibrated synthetic checkpoint, eval mode) is
    run once over 640 training clips, the pooled 640-d features are cached, and classifier.{2,3,6} (Linear,
    BatchNorm1d, Linear of the reference module) are fitted to them with Adam."""
    import torch.nn as nn
    # the calibrated statistics of the synthetic checkpoint belong to 224x224 clips: redo them for the toy clips
    bns = [m for m in model.modules() if isinstance(m, (nn.BatchNorm3d, nn.BatchNorm1d))]
    for m in bns:
        m.momentum = None
        m.reset_running_stats()
    model.train()
    for m in model.modules():
        if isinstance(m, (nn.Dropout3d, nn.Dropout)):
            m.eval()
    with torch.no_grad():
        for it in range(6):
            clips, _ = synth.synthetic_task_clips_u8(16, *CLIP, seed=700000 + it)
            model(feats(clips))
    for m in bns:
        m.momentum = 0.1
    model.eval()
    feats_, labels_ = [], []
    backbone = nn.Sequential(model.block1, model.block2, model.block3, model.block4, model.block5, model.block6,
                             model.conv, model.classifier[0], model.classifier[1])
    t0 = time.time()
    with torch.no_grad():
        for it in range(40):
            clips, labels = synth.synthetic_task_clips_u8(16, *CLIP, seed=500000 + it)
            feats_.append(backbone(feats(clips)))
            labels_.append(labels)
    X, Y = torch.cat(feats_), torch.cat(labels_)
    print(f"MoViNetA2: cached {X.shape[0]} pooled features ({time.time() - t0:.0f}s), std over clips {float(X.std(0).mean()):.4f}")
    # closed form instead of SGD (Adam on this head oscillates around ln 2): calibrate the BatchNorm1d running
    # statistics on the cached features, then a ridge regression of +-4 logit targets on the hidden activations
    # gives classifier.6 (rows -w / +w).  Every tensor involved is a parameter/buffer of the reference module.
    fc1, bn1, fc2 = model.classifier[2], model.classifier[3], model.classifier[6]
    with torch.no_grad():
        # the first 640 hidden units pass the standardised pooled features through (the random 640 -> 2048 layer of
        # the synthetic checkpoint is dominated by one common component and leaves nothing to separate)
        mu, sd_ = X.mean(0), X.std(0) + 1e-6
        fc1.weight[:640].zero_()
        fc1.weight[:640] += torch.diag(1.0 / sd_)
        fc1.bias[:640] = -mu / sd_
        Z = fc1(X)
        bn1.running_mean.copy_(Z.mean(0))
        bn1.running_var.copy_(Z.var(0, unbiased=True))
        A = torch.nn.functional.hardswish(bn1(Z)).double()            # eval mode: running statistics
        y = (Y.double() * 2 - 1) * 4.0
        A1 = torch.cat([A, torch.ones(A.shape[0], 1, dtype=torch.float64)], 1)
        lam = 1e-3 * float((A1 * A1).sum()) / A1.shape[1]
        w = torch.linalg.solve(A1.T @ A1 + lam * torch.eye(A1.shape[1], dtype=torch.float64), A1.T @ y)
        fc2.weight.copy_(torch.stack([-w[:-1], w[:-1]]).float() / 2)
        fc2.bias.copy_(torch.stack([-w[-1], w[-1]]).float() / 2)
        pred = (A1 @ w) > 0
        print(f"MoViNetA2: ridge fit, train accuracy {float((pred == (Y > 0)).double().mean()):.3f}")


def finish(name, model, log):
    # the stored checkpoint: float tensors rounded to bf16
    state = {}
    for k, v in model.state_dict().items():
        state[k] = v.detach().to(torch.bfloat16) if v.is_floating_point() else v.detach().clone()
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in state.items()})
    model.eval()
    clips, labels = synth.synthetic_task_clips_u8(EVAL_CLIPS, *CLIP, seed=EVAL_SEED)
    with torch.no_grad():
        logits = model(feats(clips))
    pred = logits.argmax(1)
    margin = (logits[:, 1] - logits[:, 0]).abs()
    print(f"{name}: eval acc {float((pred == labels).float().mean()):.3f}, predicted class counts "
          f"{torch.bincount(pred, minlength=2).tolist()}, margin min {float(margin.min()):.3f} "
          f"median {float(margin.median()):.3f}")
    torch.save({"state": state, "eval_shape": (EVAL_CLIPS,) + CLIP, "eval_seed": EVAL_SEED,
                "eval_logits": logits.clone(), "eval_labels": labels.clone(), "train_log": log},
               os.path.join(HERE, f"{name}_trained.pt"))


def main():
    only = sys.argv[1:] or list(CASES)
    for name in only:
        ctor, steps, lr = CASES[name]
        torch.manual_seed(synth.SEED_WEIGHTS)
        model = ctor(num_classes=2)
        model.initialize_weights()
        stats = torch.load(os.path.join(HERE, f"{name}_bnstats.pt"))
        template = model.state_dict()
        sd0 = synth.synthetic_state_dict(template)
        for k, v in stats.items():          # calibrated running statistics where the shapes still match
            if k in sd0 and sd0[k].shape == v.shape:
                sd0[k] = v.clone().float()
        model.load_state_dict(sd0)
        import torch.nn as nn
        if name == "MoViNetA2":
            train_movinet_head(model)
            finish(name, model, [])
            continue
        if name == "MobileNetSmall3D":
            opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, nesterov=True)
        else:   # the deeper nets do not get anywhere in 300 SGD steps: Adam, and the Dropout modules switched off
            opt = torch.optim.Adam(model.parameters(), lr=lr)
        log = []
        t0 = time.time()
        model.train()
        if name != "MobileNetSmall3D":
            for m in model.modules():
                if isinstance(m, (nn.Dropout3d, nn.Dropout)):
                    m.eval()
        for it in range(steps):
            clips, labels = synth.synthetic_task_clips_u8(16, *CLIP, seed=500000 + it)
            torch.manual_seed(it)
            logits = model(feats(clips))
            loss = torch.nn.functional.cross_entropy(logits, labels)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
            acc = float((logits.argmax(1) == labels).float().mean())
            log.append((float(loss.detach()), acc))
            if it % 20 == 0 or it == steps - 1:
                print(f"{name} step {it:4d} loss {float(loss):.4f} acc {acc:.2f} ({time.time() - t0:.0f}s)", flush=True)
        # re-calibrate the BatchNorm running statistics for the final weights (cumulative average over 8 batches,
        # dropout off): with momentum 0.1 they lag the fast-moving weights of such a short run
        bns = [m for m in model.modules() if isinstance(m, (nn.BatchNorm3d, nn.BatchNorm1d))]
        for m in bns:
            m.momentum = None
            m.reset_running_stats()
        for m in model.modules():
            if isinstance(m, (nn.Dropout3d, nn.Dropout)):
                m.eval()
        with torch.no_grad():
            for it in range(8):
                clips, _ = synth.synthetic_task_clips_u8(16, *CLIP, seed=700000 + it)
                model(feats(clips))
        for m in bns:
            m.momentum = 0.1
        finish(name, model, log)


if __name__ == "__main__":
    main()
