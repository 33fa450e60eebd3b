"""Drop-in replacements for the reference's 3D MobileNetV3 modules (``/root/reference/mobilenet.py``).

Same class names, constructor arguments, sub-module names (hence ``state_dict`` keys)
and ``forward`` contract -- so ``train.py:155-184`` can build them and ``load_state_dict`` the reference's
checkpoints -- but ``forward`` runs hand-written sm_100a kernels through ``libpicklebot_b200.so`` on NDHWC
buffers.  Parameters live in ordinary ``nn.Conv3d`` / ``nn.BatchNorm3d`` containers that are never called:
they only hold the tensors in the reference layout.

There is no CPU path: calling ``forward`` on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import blocks, ops
from .blocks import BlockCfg, BottleneckFn, MobileNetTailFn, StemFn, WeightCache


def _act_code(m: nn.Module):
    """Map the activation *module* the reference passes around (mobilenet.py:56) to a kernel code."""
    if isinstance(m, nn.Hardswish):
        return ops.ACT_HSWISH, 0.0
    if isinstance(m, nn.ReLU):
        return ops.ACT_RELU, 0.0
    if isinstance(m, nn.LeakyReLU):
        return ops.ACT_LRELU, float(m.negative_slope)
    if isinstance(m, nn.Hardsigmoid):
        return ops.ACT_HSIGMOID, 0.0
    if isinstance(m, nn.Identity):
        return ops.ACT_NONE, 0.0
    raise NotImplementedError(f"picklebot_b200: activation {type(m).__name__} has no kernel epilogue")


def _bn_args(bn: nn.BatchNorm3d):
    # momentum=None is nn.BatchNorm's cumulative moving average; pb_bn_finalize takes it as a negative momentum
    mom = -1.0 if bn.momentum is None else float(bn.momentum)
    return float(bn.eps), mom, bn.running_mean, bn.running_var, bn.num_batches_tracked


def _require_cuda(x: torch.Tensor, who: str):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: picklebot_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")


class SEBlock3D(nn.Module):
    """mobilenet.py:11-26.  Inside the bottlenecks the gate is fused into pointwise_conv2; called on its
    own (MobileNetSmall3D.block4) it pools, runs the two FCs and scales."""

    def __init__(self, channels):
        super().__init__()
        self.se = nn.Sequential(
            nn.AdaptiveAvgPool3d(1),
            nn.Conv3d(channels, channels // 4, kernel_size=1),
            nn.ReLU(inplace=True),
            nn.Conv3d(channels // 4, channels, kernel_size=1),
            nn.Hardsigmoid(),
        )

    def params(self):
        return self.se[1].weight, self.se[1].bias, self.se[3].weight, self.se[3].bias

    def forward(self, x):
        _require_cuda(x, "SEBlock3D")
        return _SEFn.apply(x, *self.params())


class _SEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x5 = blocks.to_ndhwc(x)
        B, T, H, W, C = x5.shape
        pooled = ops.pool_fwd(x5, B, C)
        hidden, gate = ops.se_fc_fwd(pooled, w1.detach().reshape(w1.shape[0], -1), b1.detach(),
                                     w2.detach().reshape(w2.shape[0], -1), b2.detach())
        y = ops.rowscale(x5, gate, B, C)
        ctx.save_for_backward(x5, pooled, hidden, gate, w1, w2)
        return blocks.from_ndhwc(y)

    @staticmethod
    def backward(ctx, dy):
        x5, pooled, hidden, gate, w1, w2 = ctx.saved_tensors
        B, T, H, W, C = x5.shape
        d5 = blocks.to_ndhwc(dy)
        if d5.dtype != x5.dtype:
            d5 = d5.to(x5.dtype)
        d5 = d5.clone()   # never scale the incoming gradient buffer in place
        dgate = ops.rowdot(d5, x5, B, C)
        dmean, dW1, db1, dW2, db2 = ops.se_fc_bwd(dgate, pooled, hidden, gate, w1.detach().reshape(w1.shape[0], -1),
                                                  w2.detach().reshape(w2.shape[0], -1), 1.0 / float(T * H * W))
        ops.scale_add_(d5, gate, dmean, B, C)
        return blocks.from_ndhwc(d5), dW1.view(w1.shape), db1, dW2.view(w2.shape), db2


class Bottleneck3D(nn.Module):
    """mobilenet.py:47-93: pointwise_conv1 -> depthwise_conv (1,k,k) with SCALAR stride and padding (time is
    padded and strided as well) -> optional squeeze_excite -> pointwise_conv2 -> batchnorm -> nonlinearity
    -> Dropout3d.  No residual connection."""

    def __init__(self, in_channels: int, out_channels: int, expanded_channels: int, stride: int = 1,
                 use_se: bool = False, kernel_size: int = 3, nonlinearity=nn.Hardswish(), batchnorm: bool = True,
                 dropout: float = 0, bias: bool = False):
        super().__init__()
        if bias:
            raise NotImplementedError("picklebot_b200: Bottleneck3D(bias=True) is not used by any reference model "
                                      "and has no kernel")
        if not batchnorm:
            # the reference would call None(x) in forward (mobilenet.py:90); refuse early instead
            raise NotImplementedError("picklebot_b200: Bottleneck3D(batchnorm=False) is not callable in the reference")
        self.pointwise_conv1 = nn.Conv3d(in_channels, expanded_channels, kernel_size=1, bias=bias)
        self.depthwise_conv = nn.Conv3d(expanded_channels, expanded_channels, groups=expanded_channels,
                                        kernel_size=(1, kernel_size, kernel_size), stride=stride,
                                        padding=kernel_size // 2, bias=bias)
        self.squeeze_excite = SEBlock3D(expanded_channels) if use_se else None
        self.pointwise_conv2 = nn.Conv3d(expanded_channels, out_channels, kernel_size=1, bias=bias)
        self.batchnorm = nn.BatchNorm3d(out_channels)
        self.nonlinearity = nonlinearity
        self.dropout = nn.Dropout3d(p=dropout)
        self._cache = WeightCache()

    def _cfg(self) -> BlockCfg:
        dw = self.depthwise_conv
        act, slope = _act_code(self.nonlinearity)
        eps, mom, _, _, _ = _bn_args(self.batchnorm)
        return BlockCfg(tuple(dw.kernel_size), tuple(dw.stride), tuple(dw.padding), act, slope,
                        self.squeeze_excite is not None, float(self.dropout.p), eps, mom)

    def forward(self, x, _mask=None):
        _require_cuda(x, "Bottleneck3D")
        dt = blocks.compute_dtype(x)
        if x.dtype != dt:
            x = x.to(dt)
        cfg = self._cfg()
        training = self.training
        bn = self.batchnorm
        mask = _mask
        if mask is None and self.dropout.training and cfg.p_drop > 0:
            mask = blocks.draw_dropout3d_mask(x.shape[0], bn.num_features, cfg.p_drop, dt, x.device)
        se = self.squeeze_excite.params() if self.squeeze_excite is not None else (None, None, None, None)
        if mask is None and blocks.eval_fold_ok(dt, bn.training, bn.running_mean):
            return blocks.bottleneck_eval(x, cfg, self._cache, bn.running_mean, bn.running_var,
                                          self.pointwise_conv1.weight, self.depthwise_conv.weight,
                                          self.pointwise_conv2.weight, bn.weight, bn.bias, *se)
        return BottleneckFn.apply(x, cfg, self._cache, bn.training, mask,
                                  bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                  self.pointwise_conv1.weight, self.depthwise_conv.weight,
                                  self.pointwise_conv2.weight, bn.weight, bn.bias, *se)


class _MobileNet3DBase(nn.Module):
    """Shared forward of MobileNetLarge3D / MobileNetSmall3D (mobilenet.py:192-201, 258-265)."""

    _tail_se = False

    def _stem(self, x, dt):
        conv, bn = self.block1[0], self.block1[1]
        eps, mom, rm, rv, nbt = _bn_args(bn)
        if (blocks.eval_fold_ok(dt, bn.training, rm) and x.dtype == torch.uint8 and x.stride(1) == 1
                and tuple(conv.kernel_size)[1:] == (3, 3)):
            return blocks.stem_eval(x, tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding), dt, eps, rm, rv,
                                    conv.weight, conv.bias, bn.weight, bn.bias)
        return StemFn.apply(x, tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding), dt, bn.training,
                            eps, mom, rm, rv, nbt, conv.weight, conv.bias, bn.weight, bn.bias)

    def _bottlenecks(self):
        raise NotImplementedError

    def _tail_modules(self):
        raise NotImplementedError

    def forward(self, x, _masks=None):
        """x: (B,3,T,H,W) float (values in [0,1], any strides; channels-last-3d is free) or uint8 (raw
        0..255 clip; the /255 of train.py:106 is fused into the stem).  Returns fp32 logits (B,num_classes).
        ``_masks``: optional list of [B][C] Dropout3d masks, one per bottleneck with p>0 (tests)."""
        _require_cuda(x, type(self).__name__)
        dt = blocks.compute_dtype(x)
        masks = list(_masks) if _masks is not None else None
        x = self._stem(x, dt)
        for blk in self._bottlenecks():
            m = None
            if masks is not None and blk.dropout.training and blk.dropout.p > 0:
                m = masks.pop(0).to(device=x.device, dtype=torch.float32).contiguous()
            x = blk(x, m)
        conv, se, bn, fc1, fc2 = self._tail_modules()
        eps, mom, rm, rv, nbt = _bn_args(bn)
        sep = se.params() if se is not None else (None, None, None, None)
        logits = MobileNetTailFn.apply(x, self._cache, bn.training, eps, mom, se is not None, rm, rv, nbt,
                                       conv.weight, conv.bias, bn.weight, bn.bias,
                                       fc1.weight, fc1.bias, fc2.weight, fc2.bias, *sep)
        return logits.view(logits.shape[0], self.num_classes)


class MobileNetLarge3D(_MobileNet3DBase):
    """mobilenet.py:133-210."""

    def __init__(self, num_classes=2):
        super().__init__()
        self.num_classes = num_classes
        self._cache = WeightCache()
        hs, relu = nn.Hardswish, nn.ReLU
        self.block1 = nn.Sequential(nn.Conv3d(3, 16, kernel_size=3, stride=2, padding=1), nn.BatchNorm3d(16),
                                    nn.Hardswish())
        B = Bottleneck3D
        self.block2 = nn.Sequential(
            B(16, 16, 16, stride=1, nonlinearity=relu(), dropout=0.2),
            B(16, 24, 64, stride=2, nonlinearity=relu(), dropout=0.2),
            B(24, 24, 72, stride=1, nonlinearity=relu(), dropout=0.2))
        self.block3 = nn.Sequential(
            B(24, 40, 72, stride=2, use_se=True, kernel_size=5, nonlinearity=relu(), dropout=0.2),
            B(40, 40, 120, stride=1, use_se=True, kernel_size=5, nonlinearity=relu(), dropout=0.2),
            B(40, 40, 120, stride=1, use_se=True, kernel_size=5, nonlinearity=relu(), dropout=0.2))
        self.block4 = nn.Sequential(
            B(40, 80, 240, stride=2, nonlinearity=hs(), dropout=0.2),
            B(80, 80, 240, stride=1, nonlinearity=hs(), dropout=0.2),
            B(80, 80, 184, stride=1, nonlinearity=hs(), dropout=0.2),
            B(80, 80, 184, stride=1, nonlinearity=hs(), dropout=0.2),
            B(80, 112, 480, stride=1, use_se=True, nonlinearity=hs(), dropout=0.2),
            B(112, 112, 672, stride=1, use_se=True, nonlinearity=hs(), dropout=0.2))
        self.block5 = nn.Sequential(
            B(112, 160, 672, stride=2, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(160, 160, 960, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(160, 160, 960, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2))
        self.block6 = nn.Sequential(nn.Conv3d(160, 960, kernel_size=1), nn.BatchNorm3d(960), nn.Hardswish())
        self.classifier = nn.Sequential(
            nn.AdaptiveAvgPool3d((1, 1, 1)),
            nn.Conv3d(960, 1280, kernel_size=1),
            nn.Hardswish(),
            nn.Conv3d(1280, self.num_classes, kernel_size=1))

    def _bottlenecks(self):
        for seq in (self.block2, self.block3, self.block4, self.block5):
            yield from seq

    def _tail_modules(self):
        return self.block6[0], None, self.block6[1], self.classifier[1], self.classifier[3]

    def initialize_weights(self):
        """mobilenet.py:203-210 tests ``hasattr(module, "nonlinearity")`` on nn.Conv3d / nn.Linear modules,
        which is never true, so the reference's initialiser changes nothing.  Same here."""
        return None


class MobileNetSmall3D(_MobileNet3DBase):
    """mobilenet.py:213-278."""

    def __init__(self, num_classes=2):
        super().__init__()
        self.num_classes = num_classes
        self._cache = WeightCache()
        hs, lrelu = nn.Hardswish, nn.LeakyReLU
        self.block1 = nn.Sequential(nn.Conv3d(3, 16, kernel_size=3, stride=2, padding=1), nn.BatchNorm3d(16),
                                    nn.Hardswish())
        B = Bottleneck3D
        self.block2 = nn.Sequential(
            B(16, 16, 16, stride=2, use_se=True, nonlinearity=lrelu(), dropout=0.2),
            B(16, 24, 72, stride=2, nonlinearity=lrelu(), dropout=0.2),
            B(24, 24, 88, stride=1, nonlinearity=lrelu(), dropout=0.2))
        self.block3 = nn.Sequential(
            B(24, 40, 96, stride=2, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(40, 40, 240, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(40, 40, 240, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(40, 48, 120, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(48, 48, 144, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(48, 96, 288, stride=2, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(96, 96, 576, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2),
            B(96, 96, 576, stride=1, use_se=True, kernel_size=5, nonlinearity=hs(), dropout=0.2))
        self.block4 = nn.Sequential(nn.Conv3d(96, 576, kernel_size=1), SEBlock3D(channels=576),
                                    nn.BatchNorm3d(576), nn.Hardswish())
        self.classifier = nn.Sequential(
            nn.AdaptiveAvgPool3d((1, 1, 1)),
            nn.Conv3d(576, 1024, kernel_size=1),
            nn.Hardswish(),
            nn.Conv3d(1024, self.num_classes, kernel_size=1))

    def _bottlenecks(self):
        for seq in (self.block2, self.block3):
            yield from seq

    def _tail_modules(self):
        return self.block4[0], self.block4[1], self.block4[2], self.classifier[1], self.classifier[3]

    def initialize_weights(self):
        """mobilenet.py:268-278: the Conv3d/Linear branch is unreachable (no module has a ``nonlinearity``
        attribute); the BatchNorm3d branch resets affine parameters to (1, 0)."""
        for module in self.modules():
            if isinstance(module, nn.BatchNorm3d):
                nn.init.constant_(module.weight, 1)
                nn.init.constant_(module.bias, 0)
