"""Driver used under ncu: one hypernet micro-step at the v4 shape (B=4, K=128, D=768, H=2048, r=32), fwd+bwd.
   python profiles/hyper_probe.py [--reps 3]"""
import argparse
import math
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dmi_b200 import augment as A  # noqa: E402
from dmi_b200.model.hypernet import HyperNetWrapper  # noqa: E402
from dmi_b200.model.projector import Projector  # noqa: E402
from dmi_b200.utils.args import HypnetArgs, ProjectorArgs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fused-grad", type=int, default=0)
a = ap.parse_args()
dev = "cuda"
D, H, r, B, K = 768, 2048, 32, 4, 128
torch.manual_seed(0)
base = Projector(ProjectorArgs(), H, D, dev)
with tempfile.NamedTemporaryFile(suffix=".pt") as f:
    torch.save({"projector_state_dict": base.state_dict()}, f.name)
    w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                        ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
w.train()
w.hypernet.fuse_generator_grad_accumulation = bool(a.fused_grad)
g = torch.Generator(device=dev).manual_seed(1)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
mm, m, t, p = rn(B, D), rn(K, D), rn(K, D), rn(1, D)
R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(0))
dy = rn(B, H) / math.sqrt(H)
keep = (torch.rand(2, 3 + 2 * K, device=dev, generator=g) >= 0.05)


def micro_step():
    x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
    a_w, b_w, biases = w.hypernet(z, keep_mask=keep)
    y = w.projector.lora_forward(x2, a_w, b_w, biases)
    y.backward(dy)


for _ in range(10):
    micro_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    micro_step()
e1.record()
torch.cuda.synchronize()
print(f"hypernet micro-step: {e0.elapsed_time(e1)/a.reps:.3f} ms (GPU events), {(time.perf_counter()-t0)/a.reps*1e3:.3f} ms wall, fused_grad={a.fused_grad}")


def stage_times():
    import time as _t
    def wall(fn, n=5):
        fn(); torch.cuda.synchronize()
        t0 = _t.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize()
        return (_t.perf_counter() - t0) / n * 1e3, r
    ta, (x2, z) = wall(lambda: A.process_embeddings(mm, (m, t, p), R=R, normalize=True))
    th, (a_w, b_w, biases) = wall(lambda: w.hypernet(z, keep_mask=keep))
    tp, y = wall(lambda: w.projector.lora_forward(x2, a_w, b_w, biases))
    def bwd():
        a_w, b_w, biases = w.hypernet(z, keep_mask=keep)
        y = w.projector.lora_forward(x2, a_w, b_w, biases)
        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        y.backward(dy)
        torch.cuda.synchronize()
        return (_t.perf_counter() - t0) * 1e3
    tb = [bwd() for _ in range(4)]
    print(f"stages (wall ms, synced): augment {ta:.3f}  hypernet fwd {th:.3f}  lora_forward fwd {tp:.3f}  backward {tb}")


stage_times()

import cProfile, pstats, io
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(4):
    x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
    a_w, b_w, biases = w.hypernet(z, keep_mask=keep)
    y = w.projector.lora_forward(x2, a_w, b_w, biases)
    y.backward(dy)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print(s.getvalue()[:3500])

# ---- host time spent inside each C-ABI call during an UNSYNCED loop ----
from dmi_b200 import _lib as L
lib = L.load()
acc = {}
import time as _t2
class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn
    def __call__(self, *a):
        t0 = _t2.perf_counter(); r = self.fn(*a); acc[self.name] = acc.get(self.name, 0.0) + (_t2.perf_counter() - t0); return r
for name in list(L.SIGNATURES):
    if name.startswith("dmi_") and name not in ("dmi_last_error",):
        setattr(lib, name, Timed(name, getattr(lib, name)))
_orig_empty_like = torch.empty_like
def timed_empty_like(*a, **k):
    t0 = _t2.perf_counter(); r = _orig_empty_like(*a, **k); acc["torch.empty_like"] = acc.get("torch.empty_like", 0.0) + (_t2.perf_counter() - t0); return r
torch.empty_like = timed_empty_like
torch.cuda.synchronize()
t0 = _t2.perf_counter()
for _ in range(6):
    micro_step()
t_loop = _t2.perf_counter() - t0
torch.cuda.synchronize()
print("unsynced loop host ms/iter", t_loop / 6 * 1e3, {k: round(v / 6 * 1e3, 3) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8]})

