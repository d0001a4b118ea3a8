"""SM clock / power while one kernel family runs back to back for ~2 s (is a TFLOP/s number power-capped?).
    python scripts/clock_probe.py bwd|fwd T active"""
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from mimic_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "bwd"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
active = (sys.argv[3] == "1") if len(sys.argv) > 3 else True
B, H, d = 16384 // T, 32, 128
q, k, v = (torch.randn(B, H, T, d, device="cuda").bfloat16() for _ in range(3))
do = torch.randn(B, T, H, d, device="cuda").bfloat16()
sh, w, b = (torch.randn(H, d, device="cuda"), torch.randn(H, d, device="cuda") * 0.1, torch.randn(H, device="cuda")) if active else (None,) * 3
out, lse_c, lse_f, mu, saved = ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
if which == "bwd":
    fn = lambda: ops.shifted_attention_backward(saved, out, do, lse_c, lse_f, mu, causal=True, active=active, scale=d ** -0.5)
    fl = (8 if active else 5) * B * H * T * T * d
else:
    fn = lambda: ops.shifted_attention_forward(q, k, v, None, sh, w, b, causal=True, active=active, scale=d ** -0.5)
    fl = (3 if active else 2) * B * H * T * T * d
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits",
                         "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
for _ in range(3):
    fn()
torch.cuda.synchronize()
time.sleep(0.3)
n0 = len(rows)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
reps = 0
e0.record()
while time.time() - t0 < 2.0:
    for _ in range(20):
        fn()
    reps += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
proc.terminate()
load = rows[n0 + 2:]
mhz = sorted(float(r.split(",")[0]) for r in load if r)
pw = sorted(float(r.split(",")[1]) for r in load if r)
print(f"{which} T={T} active={int(active)}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s | SM clock median {mhz[len(mhz) // 2]:.0f} MHz "
      f"(min {mhz[0]:.0f}, max {mhz[-1]:.0f}), power median {pw[len(pw) // 2]:.0f} W, samples {len(mhz)}, "
      f"power-capped samples {sum('Active' in r and 'Not' not in r for r in load)}")
