import sys, torch
sys.path.insert(0, "/root/repo")
from picklebot_b200 import ops, _lib
g = torch.Generator().manual_seed(1)
u8 = torch.randint(0, 256, (2, 4, 224, 224, 3), generator=g, dtype=torch.uint8).cuda()
x = u8.permute(0, 4, 1, 2, 3)
w = (torch.rand(16, 3, 3, 3, 3, generator=g) - 0.5).cuda()
b = torch.zeros(16).cuda()
y = ops.stem_fwd(x, w, b, (3, 3, 3), (2, 2, 2), (1, 1, 1), torch.bfloat16)
dy = torch.randn_like(y)
dw, db = ops.stem_wgrad(x, dy, w.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1), True)
torch.cuda.synchronize()
xr = x.float() / 255
wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
yr = torch.nn.functional.conv3d(xr, wr, br, (2, 2, 2), (1, 1, 1))
yr.backward(dy.float().permute(0, 4, 1, 2, 3))
print("ok", _lib.path_counts()["stem_tma"], float((dw - wr.grad).norm() / wr.grad.norm()), float((db - br.grad).norm() / br.grad.norm()))
