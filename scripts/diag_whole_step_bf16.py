"""Whole-step bf16 comparison in three arms (tests/test_gpu_backbones.py): CUDA kernels in a bf16 model, oracle attention in
the same bf16 model, and oracle attention with the same rounded weights in fp32 arithmetic ("truth").  Prints, per trainable
parameter, the rel-L2 gradient error of each bf16 arm against the truth and against each other."""
import os
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gpu_backbones as TB

import mimic_b200
impl = os.environ.get("DIAG_IMPL")
if impl:                                    # e.g. simt: the same step through the SIMT kernels
    _orig = mimic_b200.AttnApproximator.__init__
    def _init(self, *a, **k):
        k.setdefault("attn_impl", impl)
        _orig(self, *a, **k)
    mimic_b200.AttnApproximator.__init__ = _init
which = sys.argv[1] if len(sys.argv) > 1 else "llava"
if which == "llava":
    kw = dict(num_layers=3, hidden_size=256, intermediate_size=512, num_attention_heads=2, num_key_value_heads=2, vocab_size=1000)
    args = dict(preset="llava-interleave-7b", lora=True)
else:
    kw = dict(num_layers=5, hidden_size=256, intermediate_size=512, num_attention_heads=2, num_key_value_heads=2, vocab_size=512,
              vision_embed_dim=64)
    args = dict(preset="idefics-9b", lora=False)
bf = torch.bfloat16
for rep in range(2):
    _, _, l_k, g_k = TB._step(args["preset"], bf, False, lora=args["lora"], **kw)
    _, _, l_o, g_o = TB._step(args["preset"], bf, True, lora=args["lora"], **kw)
    _, _, l_t, g_t = TB._step(args["preset"], bf, True, lora=args["lora"], upcast=True, **kw)
    print(f"rep {rep} impl={impl or 'auto'} losses kernels {l_k} | oracle-bf16 {l_o} | truth {l_t}")
    for n in g_k:
        print(f"  {n[-60:]:60s} kernels-vs-truth {TB.rel(g_k[n], g_t[n]):.4f}  oracle-vs-truth {TB.rel(g_o[n], g_t[n]):.4f}  "
              f"kernels-vs-oracle {TB.rel(g_k[n], g_o[n]):.4f}")
