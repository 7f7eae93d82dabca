"""Device and host time of net.render() for row shards of the bench view (what each rank does at N GPUs): device time between two
events around the call, host wall time of the call itself (the GPU idles for the part of it that precedes the first launch), and
the same with a deep launch queue (the GPU never waits for the host)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import panogrf_b200 as pg
torch.manual_seed(0)
cfg = bench.cfg_dict(); cfg["mlp_dtype"] = "bf16"
net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
for rows in (512, 128, 64, 32):
    que, ref = bench.make_inputs(torch, (0, rows))
    q = {k: v.cuda() for k, v in que.items()}; r = {k: v.cuda() for k, v in ref.items()}
    for _ in range(3): net.render(q, r, False)
    torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); net.render(q, r, False); hs.append(time.perf_counter() - t0); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net.render(q, r, False)
    e1.record(); torch.cuda.synchronize()
    ms, back = sorted(ts)[3], e0.elapsed_time(e1) / 10
    print(f"rows {rows:4d} rays {rows*bench.W:7d}: event-to-event {ms:7.3f} ms, back-to-back {back:7.3f} ms, host call {1e3*sorted(hs)[3]:6.3f} ms"
          f"  -> {rows*bench.W/ms/1e3:7.2f} M rays/s")
