"""Graph-captured hypernet micro-step (v4 shape) in a clean process: replay time and parity with eager."""
import math, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import numpy as np, torch
from dmi_b200 import augment as A
from dmi_b200.graphs import GraphedStep
from dmi_b200.model.hypernet import HyperNetWrapper
from dmi_b200.model.projector import Projector
from dmi_b200.utils.args import HypnetArgs, ProjectorArgs
dev = "cuda"
D, H, r, B, K = 768, 2048, 32, 4, 128
torch.manual_seed(0)
base = Projector(ProjectorArgs(), H, D, dev)
with tempfile.NamedTemporaryFile(suffix=".pt") as f:
    torch.save({"projector_state_dict": base.state_dict()}, f.name)
    w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                        ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
w.train()
g = torch.Generator(device=dev).manual_seed(1)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
mm, m, t, p = rn(B, D), rn(K, D), rn(K, D), rn(1, D)
R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(0))
dy = rn(B, H) / math.sqrt(H)
keep = (torch.rand(2, 3 + 2 * K, device=dev, generator=g) >= 0.05)
for fused in (1, 0):
    w.hypernet.fuse_generator_grad_accumulation = bool(fused)
    static = dict(mm=mm.clone(), m=m.clone(), t=t.clone(), p=p.clone(), R=R.clone(), dy=dy.clone(), keep=keep.clone())
    def gstep():
        x2, z = A.process_embeddings(static["mm"], (static["m"], static["t"], static["p"]), R=static["R"], normalize=True)
        a_w, b_w, biases = w.hypernet(z, keep_mask=static["keep"])
        y = w.projector.lora_forward(x2, a_w, b_w, biases)
        y.backward(static["dy"])
        return y.detach()
    gs = GraphedStep(gstep, static, params=list(w.hypernet.parameters()))
    for _ in range(3):
        gs(mm=mm, R=R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        gs(mm=mm, R=R)
    e1.record()
    torch.cuda.synchronize()
    print(f"graphed micro-step (fused_grad={fused}): {e0.elapsed_time(e1)/50:.3f} ms per replay", flush=True)
    for q in w.hypernet.parameters():
        if q.grad is not None:
            q.grad.zero_()
    y_g = gs(mm=mm, R=R).clone()
    torch.cuda.synchronize()
    g_graph = {n: q.grad.clone() for n, q in w.hypernet.named_parameters() if q.grad is not None}
    del gs
    for q in w.hypernet.parameters():
        q.grad = None
    w.hypernet.fuse_generator_grad_accumulation = False
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
        a_w, b_w, biases = w.hypernet(z, keep_mask=keep)
        y_e = w.projector.lora_forward(x2, a_w, b_w, biases)
        y_e.backward(dy)
    torch.cuda.synchronize()
    worst = max(((g_graph[n] - q.grad).norm() / q.grad.norm().clamp_min(1e-20)).item() for n, q in w.hypernet.named_parameters()
                if q.grad is not None and q.grad.norm().item() > 1e-9)        # k.bias grad is mathematically zero (softmax shift invariance)
    print(f"GRAPH_PARITY fused={fused} worst={worst:.3e} y={((y_g - y_e.detach()).norm() / y_e.detach().norm()).item():.3e} n_grads={len(g_graph)}", flush=True)
    print({n: (round(((g_graph[n] - q.grad).norm() / q.grad.norm().clamp_min(1e-20)).item(), 6), float(q.grad.norm())) for n, q in w.hypernet.named_parameters() if q.grad is not None})
    print("   graph vs eager: y rel diff", ((y_g - y_e.detach()).norm() / y_e.detach().norm()).item(), " worst grad rel diff", worst, len(g_graph), flush=True)
    del a_w, b_w, biases, y_e, x2, z
    for q in w.hypernet.parameters():
        q.grad = None

# ---- gradient accumulation as ONE graph, generator gradient kept as rank-1 factors (wrapper path: generator 0 only) vs eager dense accumulation ----
from dmi_b200.parallel import Rank1FactorSync
GA = 3
with tempfile.NamedTemporaryFile(suffix=".pt") as f:          # fresh parameters: their AccumulateGrad nodes belong to this capture only
    torch.save({"projector_state_dict": base.state_dict()}, f.name)
    w2 = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                         ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
w2.load_state_dict(w.state_dict())
w2.train()
mms = [rn(B, D) for _ in range(GA)]
gen0 = w2.hypernet.generators[0]
small = [q for n, q in w2.hypernet.named_parameters() if not n.startswith("generators")]
sink = Rank1FactorSync(gen0.weight.shape[0], D, dev, max_terms=GA)
w2.hypernet.fuse_generator_grad_accumulation = True
w2.hypernet.factor_sinks = {0: sink}
static2 = dict(mms=torch.stack(mms).clone())


def ga_loop():
    sink.n = 0
    for j in range(GA):
        x2, z = A.process_embeddings(static2["mms"][j], (m, t, p), R=R, normalize=True)
        a_w, b_w, biases = w2.hypernet(z, keep_mask=keep, n_layers=1)
        w2.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0]).backward(dy)


gs = GraphedStep(ga_loop, static2, params=small)
for q in small:
    q.grad.zero_()
gs(mms=torch.stack(mms))
gen0.weight.grad, gen0.bias.grad = torch.zeros_like(gen0.weight), torch.zeros_like(gen0.bias)
sink.n = GA
sink.apply_(gen0.weight.grad, gen0.bias.grad)
torch.cuda.synchronize()
g_graph = {n: q.grad.clone() for n, q in w2.hypernet.named_parameters() if q.grad is not None}
del gs
for q in w.hypernet.parameters():
    q.grad = None
w.hypernet.fuse_generator_grad_accumulation = False
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for j in range(GA):                                        # eager, dense, through the public wrapper forward (reference call path)
        x2, z = A.process_embeddings(mms[j], (m, t, p), R=R, normalize=True)
        a_w, b_w, biases = w.hypernet(z, keep_mask=keep)
        w.projector.lora_forward(x2, a_w, b_w, biases).backward(dy)
torch.cuda.synchronize()
ref = {n: q.grad for n, q in w.hypernet.named_parameters() if q.grad is not None}
worst = max(((g_graph[n] - ref[n]).norm() / ref[n].norm().clamp_min(1e-20)).item() for n in ref if ref[n].norm().item() > 1e-9)
print(f"GRAPH_PARITY_GA worst={worst:.3e} n_grads={len(g_graph)} n_ref={len(ref)}", flush=True)
