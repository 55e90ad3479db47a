"""Runs every kernel class of libdmi_b200 once inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`):
the bench step (full adapted MLP fwd+bwd, B=32768), a plain-MLP2 training step (MN-major dW GEMMs, dropout), one hypernet
micro-step (augment, pooling, generators, backward dense and in factor mode), the few-shot mean adapter + merge, the splice, the
embedding-store gather, the device isometry draw, fused clip + AdamW and the one-shot all-reduce kernel (single rank).
   python profiles/all_kernels_probe.py   (plain)   |   ncu --profile-from-start off --metrics ... python profiles/all_kernels_probe.py"""
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
D, H, r, B = 768, 2048, 32, 32768
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
from dmi_b200 import _lib
from dmi_b200.data import EmbeddingStore
from dmi_b200.optim import FusedAdamW
lib = _lib.load()
zs4 = [A.process_embeddings(None, (rn(32, D), rn(32, D), rn(1, D)), R=None, normalize=True)[1] for _ in range(4)]
store = EmbeddingStore(rn(200000, D))
sidx = torch.randint(0, 200000, (B,), device=dev, generator=g)
sout, sbf = torch.empty(B, D, device=dev), torch.empty(B, D, device=dev, dtype=torch.bfloat16)
fo = FusedAdamW([q for n, q in w.hypernet.named_parameters() if "generators.1" not in n], lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=5e-6)
ar_in, ar_out = rn(225280), torch.zeros(225280, device=dev)
ar_flags = torch.zeros(int(lib.dmi_allreduce_flag_words()), dtype=torch.int32, device=dev)
epoch = [0]


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
    splice_prefix(rn(32, H), table, ids, None, None, torch.bfloat16)
    # factor mode: the generator backward only reads G
    from dmi_b200.parallel import Rank1FactorSync
    w.hypernet.factor_sinks = {0: Rank1FactorSync(w.hypernet.generators[0].weight.shape[0], D, dev, max_terms=2)}
    a_w, b_w, biases = w.hypernet(z, n_layers=1)
    w.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0]).backward(dy4)
    w.hypernet.factor_sinks = None
    # few-shot: mean adapter of 4 support sets through one generator pass + exact merge
    w.eval()
    w.generate_projector_from_multiple_adapters(zs4)
    w.generated_projector = None
    w.train()
    store.gather(sidx, out=sout, out_bf16=sbf)                          # embedding store
    A.get_rotation_matrix_device(D, dev, generator=g)                   # device isometry
    fo.step(max_grad_norm=1.0)                                          # fused clip + AdamW over the hypernet
    import ctypes as C
    arr = C.c_void_p * 1
    epoch[0] += 1
    _lib.check(lib.dmi_allreduce_oneshot(arr(ar_in.data_ptr()), arr(ar_flags.data_ptr()), None, 0, 1, C.c_void_p(ar_out.data_ptr()), ar_in.numel(), 1.0,
                                         epoch[0], C.c_void_p(torch.cuda.current_stream().cuda_stream)), "allreduce")


everything()
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done, launches:", ops.launch_count())
