import torch,sys
sys.path.insert(0,".")
from mimic_b200 import ops
for (B,T) in ((2,3200),(2,128)):
    qkv = torch.randn(B, T, (32 + 16) * 128, device="cuda").bfloat16()
    qv, kv_ = qkv[..., : 32 * 128].view(B, T, 32, 128), qkv[..., 32 * 128: 40 * 128].view(B, T, 8, 128)
    cs = torch.randn(B, T, 128, device="cuda").bfloat16()
    for _ in range(3): ops.rope_qk(qv, kv_, cs, cs)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ops.rope_qk(qv, kv_, cs, cs)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/50
    print("rope", B, T, round(ms*1e3,1), "us", round(2*40*128*2*B*T/ms/1e6), "GB/s")
