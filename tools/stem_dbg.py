import sys, torch
sys.path.insert(0, "/root/repo")
from picklebot_b200 import ops
g = torch.Generator().manual_seed(1)
u8 = torch.randint(0, 256, (2, 6, 64, 64, 3), generator=g, dtype=torch.uint8).cuda()
x = u8.permute(0, 4, 1, 2, 3)
w = (torch.rand(16, 3, 3, 3, 3, generator=g) - 0.5).cuda()
b = torch.zeros(16).cuda()
y = ops.stem_fwd(x, w, b, (3, 3, 3), (2, 2, 2), (1, 1, 1), torch.bfloat16)
torch.cuda.synchronize()
yr = torch.nn.functional.conv3d(x.float() / 255, w.half().float(), b, (2, 2, 2), (1, 1, 1)).permute(0, 2, 3, 4, 1)
print("ok", float((y.float() - yr).norm() / yr.norm()))
