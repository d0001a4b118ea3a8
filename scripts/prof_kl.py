import sys, torch
sys.path.insert(0, ".")
from mimic_b200 import ops
R, V = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1540, 152000)
z = (torch.randn(R, V, device="cuda") * 2).bfloat16().requires_grad_(True)
y = (torch.randn(R, V, device="cuda") * 2).bfloat16()
rows = torch.arange(R, dtype=torch.int32, device="cuda")
for _ in range(4):
    ops.kl_loss(z, y, rows)
torch.cuda.synchronize()
print("ok")
