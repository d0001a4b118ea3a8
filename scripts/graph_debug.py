import sys, types, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_trainer import _setup
from mimic_b200 import ops
from mimic_b200.trainer import MimicTrainer

variant = sys.argv[1] if len(sys.argv) > 1 else "full"
main_stream = torch.cuda.Stream()
torch.cuda.set_stream(main_stream)      # everything (eager warm-up included) on the stream that will capture
sm, enc, full, qa = _setup(torch.float32, seed=7)
tr = MimicTrainer(sm, lr=5e-3, warmup_step=2, total_steps=50)
for i in range(3):
    tr.micro_step({"input_ids": full}, {"input_ids": qa})
torch.cuda.synchronize()


def body():
    if variant == "fwd":
        with torch.no_grad():
            sm.forward_inputs({"input_ids": full}, {"input_ids": qa})
    elif variant == "ce_only":
        from mimic_b200 import Strategy
        sm.strategy = Strategy.LM_LOSS
        ld = sm.forward_inputs(None, {"input_ids": qa})
        ld["loss"].backward()
    elif variant == "nosink":
        enc.grad_sinks = None
        ld = sm.forward_inputs({"input_ids": full}, {"input_ids": qa})
        ld["loss"].backward()
    else:
        ld = sm.forward_inputs({"input_ids": full}, {"input_ids": qa})
        ld["loss"].backward()


body()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=main_stream):
        body()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    print(variant, "OK")
except Exception as e:
    print(variant, "FAIL", str(e).split("\n")[0])
