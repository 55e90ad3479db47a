"""One launch of each kernel class added after the first all-kernels capture (optimizer, embedding gather, Haar draw,
register-streaming side kernels, merged pack), with a small memory footprint, inside a cudaProfilerStart/Stop window:
  ncu --profile-from-start off --metrics <light set> python profiles/new_kernels_probe.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import numpy as np, torch
from dmi_b200 import augment as A, ops
from dmi_b200.data import EmbeddingStore
from dmi_b200.optim import FusedAdamW
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
D, H, r, B = 768, 2048, 32, 32768
params = [torch.nn.Parameter(rn(92160, 768) * 0.05), torch.nn.Parameter(rn(768, 768))]           # generators.0-sized + q
for q in params:
    q.grad = rn(*q.shape) * 1e-3
opt = FusedAdamW(params, lr=1e-4, betas=(0.9, 0.95), weight_decay=5e-6)
store = EmbeddingStore(rn(100000, 1024), selected_features=np.arange(768), mean=rn(768) * 0.01)
sidx = torch.randint(0, 100000, (B,), device=dev, generator=g)
sout, sbf = torch.empty(B, 768, device=dev), torch.empty(B, 768, device=dev, dtype=torch.bfloat16)
hb = rn(B, H).to(torch.bfloat16)
Wr = (rn(r, H) / math.sqrt(H)).to(torch.bfloat16)
lq = torch.empty(ops.lq_words(B, r), device=dev, dtype=torch.int32)
Gs, cs = torch.zeros(r, H, device=dev), torch.zeros(H, device=dev)
w1, w2 = rn(H, D) / math.sqrt(D), rn(H, H) / math.sqrt(H)
zv = torch.zeros(H, device=dev)
pk = ops.PackedProjector(D, H, r, dev, merged=True)
A0, B0, A1, B1 = rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, rn(H * r) / math.sqrt(H), rn(r * H) * 0.1


def everything():
    opt.step(max_grad_norm=1.0)
    store.gather(sidx, out=sout, out_bf16=sbf)
    A.get_rotation_matrix_device(256, dev, generator=g)
    ops.stream_project(hb, Wr, out_lq=lq)
    ops.stream_reduce(lq, r, hb, Gs, colsum=cs)
    pk.pack_adapter_merged(w1, w2, A0, B0, zv, A1, B1, zv, zv, zv)


everything()
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
