import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, panogrf_b200 as pg
dev=torch.device("cuda")
B, Hc, Wc, C, D = 1, 256, 512, 32, 64
g = torch.Generator(device=dev).manual_seed(0)
images = torch.randn(B, 2, Hc, Wc, C, device=dev, generator=g)
rots = torch.eye(3, device=dev).expand(B, 2, 3, 3).contiguous()
trans = torch.zeros(B, 2, 3, device=dev); trans[:, 0, 2], trans[:, 1, 2] = 0.5, -0.5
depths = torch.linspace(0.1, 10, D, device=dev)
args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
for _ in range(3):
    pg.calculate_cost_volume_erp(args, images, depths, trans, rots)
torch.cuda.synchronize()
