"""picklebot_b200 -- B200-native (sm_100a) forward/backward of Picklebot's 3D mobile CNNs.

Drop-in ``nn.Module`` classes with the reference's names, constructor arguments and ``state_dict`` layout
(``/root/reference/mobilenet.py``, ``/root/reference/movinet.py``), backed by hand-written CUDA kernels in
``libpicklebot_b200.so`` (C ABI: ``include/picklebot_b200.h``).  No CPU fallback.
"""
from .mobilenet import Bottleneck3D, MobileNetLarge3D, MobileNetSmall3D, SEBlock3D
from .movinet import CausalConv3d, MoViNetA2, MoviNetBottleneck

# the subset of train.py:156-161's registry that is on the hot path
valid_models = {
    "MobileNetLarge3D": MobileNetLarge3D,
    "MobileNetSmall3D": MobileNetSmall3D,
    "MoViNetA2": MoViNetA2,
}

__all__ = ["SEBlock3D", "Bottleneck3D", "MobileNetLarge3D", "MobileNetSmall3D", "CausalConv3d",
           "MoviNetBottleneck", "MoViNetA2", "valid_models"]
