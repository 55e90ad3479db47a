"""Where the graphed hypernet micro-step goes: CUDA-graph replay time of each stage alone (v4 shape, B=4, K=128)."""
import math, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import numpy as np, torch
from dmi_b200 import augment as A
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
w.hypernet.fuse_generator_grad_accumulation = True
for q in w.hypernet.parameters():
    q.grad = torch.zeros_like(q)
g = torch.Generator(device=dev).manual_seed(1)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
mm, m, t, p = rn(B, D), rn(K, D), rn(K, D), rn(1, D)
R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(0))
dy = rn(B, H) / math.sqrt(H)
keep = (torch.rand(2, 3 + 2 * K, device=dev, generator=g) >= 0.05)
x2s, zs = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)


def replay_us(fn, reps=200):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        fn()
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def f_aug():
    A.process_embeddings(mm, (m, t, p), R=R, normalize=True)


def f_hyp():
    with torch.no_grad():
        w.hypernet(zs, keep_mask=keep, n_layers=1)


def f_proj():
    with torch.no_grad():
        a_w, b_w, biases = hold
        w.projector.lora_forward_first_layer(x2s, a_w[0], b_w[0], biases[0])


def f_fwd():
    with torch.no_grad():
        x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
        a_w, b_w, biases = w.hypernet(z, keep_mask=keep, n_layers=1)
        w.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0])


def f_all():
    x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
    a_w, b_w, biases = w.hypernet(z, keep_mask=keep, n_layers=1)
    w.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0]).backward(dy)


with torch.no_grad():
    hold = w.hypernet(zs, keep_mask=keep, n_layers=1)
print(f"augment                    {replay_us(f_aug):7.1f} us")
print(f"hypernet forward (1 gen)   {replay_us(f_hyp):7.1f} us")
print(f"projector forward (B=4)    {replay_us(f_proj):7.1f} us")
print(f"whole forward              {replay_us(f_fwd):7.1f} us")
print(f"forward + backward (dense) {replay_us(f_all):7.1f} us", flush=True)
