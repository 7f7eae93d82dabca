import os, sys, json
sys.path.insert(0, "/root/repo")
import torch, bench
import panogrf_b200 as pg
dev = torch.device("cuda")
v, t, o, idx, Wd = bench.oracle_run(8192)
cfg = bench.cfg_dict()
que, ref = bench.make_inputs(torch)
q = {k: v_.to(dev) for k, v_ in que.items()}
q["coords"] = q["coords"][:, idx.to(dev)]
r = {k: v_.to(dev) for k, v_ in ref.items()}
net = pg.NeuralRayBaseRenderer({**cfg, "mlp_dtype": "fp32"}).to(dev).eval()
net.load_state_dict(Wd, strict=False)
out = {k: v_.float().cpu() for k, v_ in net.render(q, r, False).items() if v_.dtype.is_floating_point}
for k, s in (("pixel_colors_nr", 1.0), ("render_depth", 14.5), ("pixel_colors_nr_fine", 1.0), ("render_depth_fine", 14.5)):
    d = (out[k] - o[k]).abs()
    if d.dim() == 3: d = d.amax(-1)
    d = d[0]
    tol = 1e-4 * o[k].abs().reshape(d.shape[0], -1).amax(-1) + 1e-4 * s
    bad = torch.nonzero(d > tol).flatten()
    print(k, "n_bad", bad.numel(), "of", d.numel(), "quantiles", [float(torch.quantile(d.double(), p)) for p in (0.5, 0.99, 0.999, 1.0)])
    c = que["coords"][0, idx][bad[:12]]
    print("   bad coords (x,y):", c.tolist(), "errs", d[bad[:12]].tolist())
