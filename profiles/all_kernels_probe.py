"""Runs every kernel class of libdmi_b200 once inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`):
the bench step (full adapted MLP fwd+bwd, B=16384), a plain-MLP2 training step (MN-major dW GEMMs, dropout), one hypernet
micro-step (augment, pooling, generators, backward) and the splice."""
import math, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import numpy as np, torch
from dmi_b200 import augment as A, ops
from dmi_b200.model.hypernet import HyperNetWrapper
from dmi_b200.model.mmmodel import splice_prefix
from dmi_b200.model.projector import Projector
from dmi_b200.utils.args import HypnetArgs, ProjectorArgs
dev = "cuda"
D, H, r, B = 768, 2048, 32, 16384
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
torch.manual_seed(0)
base = Projector(ProjectorArgs(proj_dropout=0.1), H, D, dev)
with tempfile.NamedTemporaryFile(suffix=".pt") as f:
    torch.save({"projector_state_dict": base.state_dict()}, f.name)
    w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                        ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
w.train()
base.train()
x, dy = rn(B, D), rn(B, H) / math.sqrt(H)
leaves = [t.requires_grad_(True) for t in (rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, rn(H) * 0.1, rn(H * r) / math.sqrt(H), rn(r * H) * 0.1, rn(H) * 0.1)]
mm, m, t, p = rn(4, D), rn(128, D), rn(128, D), rn(1, D)
R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(0))
dy4 = rn(4, H)
table = rn(32000, H).to(torch.bfloat16)
ids = torch.randint(0, 32000, (32, 320), device=dev, generator=g)


def everything():
    base.lora_forward_mode = "full"
    a0, b0, be0, a1, b1, be1 = leaves
    y = base.lora_forward(x, [a0, a1], [b0, b1], [be0, be1])          # full adapted MLP fwd
    y.backward(dy)                                                      # + bwd to the adapter
    base(x).backward(dy)                                                # plain MLP2 with dropout: dW / db (MN-major GEMMs)
    x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)    # augmentation
    a_w, b_w, biases = w.hypernet(z)                                    # pooling + generators
    base.lora_forward_mode = "as_written"
    w.projector.lora_forward(x2, a_w, b_w, biases).backward(dy4)        # H1 projector + hypernet backward
    splice_prefix(rn(32, H), table, ids, None, None, torch.float32)


everything()
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done, launches:", ops.launch_count())
