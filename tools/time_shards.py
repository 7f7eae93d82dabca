"""Device time of net.render() for row shards of the bench view (what each rank does at N GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import panogrf_b200 as pg
torch.manual_seed(0)
cfg = bench.cfg_dict(); cfg["mlp_dtype"] = "bf16"
net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
for rows in (512, 256, 128, 64, 32):
    que, ref = bench.make_inputs(torch, (0, rows))
    q = {k: v.cuda() for k, v in que.items()}; r = {k: v.cuda() for k, v in ref.items()}
    for _ in range(3): net.render(q, r, False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.render(q, r, False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"rows {rows:4d} rays {rows*bench.W:7d}: {ms:8.3f} ms  {rows*bench.W/ms/1e3:7.2f} M rays/s")
