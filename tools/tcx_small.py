import sys, torch
sys.path.insert(0, "/root/repo")
from diffudf_b200 import SIREN
torch.manual_seed(0)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
eng = m._engine_synced()
x = torch.rand(4096, 3, device="cuda") * 2 - 1
for order in (1, 0, 2):
    f, g, H, _ = eng.query(x, order, "tcx3")
    torch.cuda.synchronize()
    print("order", order, "ok", float(f.abs().max()))
