"""One hypernet micro-step at the v4 shape (B=4, K=128, D=768, H=2048, r=32) exactly as bench.py times it: augment + HyperNetWrapper
forward as written (generator 0 only) + backward with the generator gradient accumulated in place.  Under
`ncu --profile-from-start off` the cudaProfilerStart/Stop pair brackets exactly one micro-step."""
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
g = torch.Generator(device=dev).manual_seed(1)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
mm, m, t, p = rn(B, D), rn(K, D), rn(K, D), rn(1, D)
R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(0))
dy = rn(B, H) / math.sqrt(H)
keep = (torch.rand(2, 3 + 2 * K, device=dev, generator=g) >= 0.05)


def micro_step():
    x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
    a_w, b_w, biases = w.hypernet(z, keep_mask=keep, n_layers=1)
    w.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0]).backward(dy)


for _ in range(5):
    micro_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
micro_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    micro_step()
e1.record()
torch.cuda.synchronize()
print(f"hypernet micro-step, eager: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
