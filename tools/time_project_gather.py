import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.manual_seed(0)
que, ref = bench.make_inputs(torch)
ref_d = {k: v.cuda() for k, v in ref.items()}
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
print(json.dumps(bench.time_project_gather(torch, None, ref_d, flush, bench.measured_peaks())))
