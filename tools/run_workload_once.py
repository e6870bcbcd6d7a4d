import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name = sys.argv[1]
wl = bench.Workload(name)
if name == "cfg5": wl.B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl.build("cuda")
x, prm = wl.host_inputs()
x = x.cuda(); prm = bench.tree_to(prm, "cuda")
with torch.no_grad():
    for _ in range(3):
        y = wl.forward(x, prm)
torch.cuda.synchronize()
