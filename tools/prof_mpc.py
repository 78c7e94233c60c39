"""Where does an MPC scoring pass spend its time?  torch.profiler table of one c5-pickplace / c5-stack pass (diagnostic)."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import op3_b200
from op3_b200.mpc import Cost
from op3_b200 import parallel as par

name = sys.argv[1] if len(sys.argv) > 1 else "c5-pickplace"
cfg = bench.CONFIGS[name]
op3_b200.set_precision("f16x3")
dev = torch.device("cuda", 0)
model = bench.build_model(cfg, dev)
model.eval_mode = True
model.inference_chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Na, K, D = cfg["Na"], cfg["K"], cfg["D"]
g = torch.Generator().manual_seed(31)
sched = np.array(cfg["schedule"], dtype=np.float64)
seed_fr = (torch.randint(0, 256, (1, 3, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0).to(dev)
seed_ac = (2 * torch.rand(1, 3, 4, generator=g) - 1).to(dev)
info = model.batch_internal_inference(seed_fr, seed_ac, None, np.array([0., 0., 1., 0., 1., 0.]))
cand = (torch.rand(Na, 1, 4, generator=g)).to(dev)
goal_info = {"goal_image": torch.rand(1, 3, D, D, generator=g).to(dev)}
cost = Cost(None, core_type="final_recon", aggregate="sum")
fn = lambda: par.sharded_rollout_costs(model, cost, goal_info, cand, info["state"], sched)
for _ in range(5):
    fn()
torch.cuda.synchronize()
import time
for rep in range(3):
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); t_host = time.time() - t0
    torch.cuda.synchronize()
    print("pass %d: device %.2f ms, host enqueue %.2f ms" % (rep, e0.elapsed_time(e1), t_host * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fn()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
