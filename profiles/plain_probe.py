"""configs[3] plain MLP2 train step (B=1024, dropout 0.1) through the module API: launch list under ncu / graph replay time."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200.graphs import GraphedStep
from dmi_b200.model.mlp2 import plain_mlp2
from dmi_b200.model.projector import Projector
from dmi_b200.utils.args import ProjectorArgs
dev = "cuda"
D, H, B = 768, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
base = Projector(ProjectorArgs(proj_dropout=0.1), H, D, dev)
base.train()
g = torch.Generator(device=dev).manual_seed(1)
x, dy = torch.randn(B, D, device=dev, generator=g), torch.randn(B, H, device=dev, generator=g) / math.sqrt(H)
for q in base.parameters():
    q.grad = torch.zeros_like(q)
for _ in range(3):
    base(x).backward(dy)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: exactly one step (all threads: the backward runs on autograd's)
base(x).backward(dy)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if os.environ.get("PLAIN_GRAPH", "1") == "1":
    lin0, lin1 = base.net[0], base.net[-1]
    static = dict(x=x.clone(), dy=dy.clone())
    def step():
        plain_mlp2(static["x"], lin0.weight, lin0.bias, lin1.weight, lin1.bias, dropout_p=0.1, cache=None, grad_in_place=True).backward(static["dy"])
    gs = GraphedStep(step, static, params=[])          # gradients are accumulated in place: no AccumulateGrad nodes in the capture
    for _ in range(3):
        gs(x=x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        gs(x=x)
    e1.record()
    torch.cuda.synchronize()
    print(f"plain MLP2 B={B} fwd+bwd graph replay: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
