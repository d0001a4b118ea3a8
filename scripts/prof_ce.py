import sys, torch
sys.path.insert(0, ".")
from mimic_b200 import ops
R, V = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1540, 152000)
logits = (torch.randn(R, V, device="cuda") * 2).bfloat16().requires_grad_(True)
labels = torch.randint(0, V, (R,), device="cuda")
for _ in range(4):
    ops.ce_loss(logits, labels)
torch.cuda.synchronize()
print("ok")
