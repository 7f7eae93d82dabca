"""ncu target: one steady-state forward of the image encoder (run with `ncu --profile-from-start off`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panogrf_b200.image_encoder import ResUNetLight
torch.manual_seed(0)
net = ResUNetLight({}, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=True).cuda()
x = torch.rand(2, 3, 512, 1024, device="cuda")
for _ in range(3): net(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
