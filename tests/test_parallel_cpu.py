"""Data-parallel host logic on the CPU with gloo, world_size 2 (the NCCL path is exercised on the GPU box by bench.py).

The reference has no DP; its analogue is gradient accumulation.  Equivalence pinned here:
  * hypernet training: one micro-step per rank, all-reduce(SUM) of grads of loss/(world*GA_local)  ==  the reference's
    single-process accumulation over GA = world*GA_local micro-steps (train_hypernet.py:119-149);
  * projector training: batch rows sharded across ranks, all-reduce(mean)  ==  single-process mean loss over the batch.
Gradients come from the CPU oracle here (test infrastructure); the buckets / reducer are the product code."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O

WORLD = 2
DIMS = dict(D=32, H=48, r=4, n_tokens=3, K=3, B=5)


def make_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    D, H, r = DIMS["D"], DIMS["H"], DIMS["r"]
    rn = lambda *s: torch.randn(*s, generator=g) * 0.2
    p = {"hypernet.prefix_tokens": rn(2, D)}
    for n in "qkv":
        p[f"hypernet.hypnet.{n}.weight"], p[f"hypernet.hypnet.{n}.bias"] = rn(D, D), rn(D)
    p["hypernet.generators.0.weight"], p["hypernet.generators.0.bias"] = rn(D * r + r * H + H, D), rn(D * r + r * H + H)
    p["hypernet.generators.1.weight"], p["hypernet.generators.1.bias"] = rn(H * r + r * H + H, D), rn(H * r + r * H + H)
    p["projector.net.0.weight"], p["projector.net.0.bias"] = rn(H, D), rn(H)
    p["projector.net.3.weight"], p["projector.net.3.bias"] = rn(H, H), rn(H)
    return p


def micro_batch(i):
    g = torch.Generator().manual_seed(1000 + i)
    D, H, K, B = DIMS["D"], DIMS["H"], DIMS["K"], DIMS["B"]
    x = O.l2_normalize(torch.randn(B, D, generator=g))
    z = O.l2_normalize(torch.randn(1 + 2 * K, D, generator=g))
    dy = torch.randn(B, H, generator=g)
    return x, z, dy


HYPER_KEYS = ["hypernet.prefix_tokens"] + [f"hypernet.hypnet.{n}.{w}" for n in "qkv" for w in ("weight", "bias")] + \
             ["hypernet.generators.0.weight", "hypernet.generators.0.bias"]


def hyper_grads(params, i, scale):
    leaves = {k: v.clone().requires_grad_(k in HYPER_KEYS) for k, v in params.items()}
    x, z, dy = micro_batch(i)
    out = O.hypernet_wrapper_forward(leaves, x, z, n_tokens=DIMS["n_tokens"], rank=DIMS["r"], alpha=8.0, lm_dim=DIMS["H"], mm_dim=DIMS["D"])
    ((out * dy).sum() * scale).backward()
    return {k: leaves[k].grad for k in HYPER_KEYS}


def _worker(rank, init_file, result_file):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=WORLD)
    from dmi_b200.parallel import BucketAllReducer, FlatGrads, allreduce_module_grads
    params = make_params()
    GA_local = 2
    # ---- hypernet: micro-steps rank, rank+WORLD, ... (rank-ordered interleave reproduces the single-process draw order) ----
    shapes = {k: tuple(params[k].shape) for k in HYPER_KEYS}
    buckets = [["hypernet.generators.0.weight", "hypernet.generators.0.bias"],                     # available first in backward
               [k for k in HYPER_KEYS if "generators" not in k]]
    fg = FlatGrads(shapes, buckets, "cpu")
    for j in range(GA_local):
        g = hyper_grads(params, j * WORLD + rank, 1.0 / (WORLD * GA_local))
        for k in HYPER_KEYS:
            fg[k].add_(g[k])
    red = BucketAllReducer(average=False)
    for b in fg.buckets:
        red.reduce_bucket(b)
    red.wait()
    # ---- projector: rows sharded, mean loss ----
    g = torch.Generator().manual_seed(7)
    X = O.l2_normalize(torch.randn(8, DIMS["D"], generator=g))
    DY = torch.randn(8, DIMS["H"], generator=g)
    lo, hi = rank * 4, rank * 4 + 4
    leaves = [params[k].clone().requires_grad_(True) for k in ("projector.net.0.weight", "projector.net.0.bias", "projector.net.3.weight", "projector.net.3.bias")]
    sd = dict(zip(("projector.net.0.weight", "projector.net.0.bias", "projector.net.3.weight", "projector.net.3.bias"), leaves))
    (O.projector_forward(sd, X[lo:hi]) * DY[lo:hi]).sum(1).mean().backward()
    allreduce_module_grads(leaves, average=True)
    if rank == 0:
        torch.save({"hyper": {k: fg[k].clone() for k in HYPER_KEYS}, "proj": [l.grad for l in leaves]}, result_file)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_dp_allreduce_equals_gradient_accumulation():
    with tempfile.TemporaryDirectory() as d:
        init_file, result_file = os.path.join(d, "init"), os.path.join(d, "res.pt")
        mp.spawn(_worker, args=(init_file, result_file), nprocs=WORLD, join=True)
        res = torch.load(result_file)
    params = make_params()
    GA = WORLD * 2
    ref = None
    for i in range(GA):                       # the reference's loop: loss / GA, backward, accumulate
        g = hyper_grads(params, i, 1.0 / GA)
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    for k in HYPER_KEYS:
        torch.testing.assert_close(res["hyper"][k], ref[k], rtol=1e-5, atol=1e-7)
    g = torch.Generator().manual_seed(7)
    X = O.l2_normalize(torch.randn(8, DIMS["D"], generator=g))
    DY = torch.randn(8, DIMS["H"], generator=g)
    names = ("projector.net.0.weight", "projector.net.0.bias", "projector.net.3.weight", "projector.net.3.bias")
    leaves = [params[k].clone().requires_grad_(True) for k in names]
    (O.projector_forward(dict(zip(names, leaves)), X) * DY).sum(1).mean().backward()
    for a, b in zip(res["proj"], leaves):
        torch.testing.assert_close(a, b.grad, rtol=1e-5, atol=1e-7)


def test_flat_grads_layout():
    from dmi_b200.parallel import FlatGrads
    fg = FlatGrads({"a": (3, 5), "b": (7,), "c": (2, 2)}, [["a", "b"], ["c"]], "cpu")
    assert fg["a"].shape == (3, 5) and fg["b"].shape == (7,) and fg["c"].shape == (2, 2)
    assert sum(b.numel() for b in fg.buckets) == fg.flat.numel()
    fg["b"].fill_(2.0)
    assert fg.buckets[0].sum().item() == 14.0 and fg.buckets[1].sum().item() == 0.0
    for v in fg.views.values():
        assert v.data_ptr() % 16 == 0
    fg.zero_()
    assert fg.flat.abs().sum().item() == 0
    with pytest.raises(AssertionError):
        FlatGrads({"a": (1,)}, [["a"], ["a"]], "cpu")


def _fewshot_worker(rank, init_file, result_file):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=WORLD)
    from dmi_b200.parallel import allreduce_sum_, shard_support_sets
    params = make_params()
    n_sets = 5
    zs = [micro_batch(100 + i)[1] for i in range(n_sets)]
    mine = shard_support_sets(list(range(n_sets)))
    partial = None
    for i in mine:
        a_w, b_w, bias = O.hypernetwork_forward(params, zs[i], n_tokens=DIMS["n_tokens"], rank=DIMS["r"], alpha=8.0, lm_dim=DIMS["H"], mm_dim=DIMS["D"])
        flat = torch.cat([t.reshape(-1) for t in (*a_w, *b_w, *bias)]) / n_sets
        partial = flat if partial is None else partial + flat
    allreduce_sum_(partial)
    if rank == 0:
        torch.save({"mean": partial, "mine": mine}, result_file)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_fewshot_support_sets_shard_and_mean_over_ranks():
    """support sets round-robin over the ranks + SUM all-reduce of the (1/N)-weighted partial results == the single-process mean
    adapter of generate_projector_from_multiple_adapters (hypernet.py:234-266); linear in the per-set results, so the same holds
    for the modality codes that HyperNetwork.mean_adapter(zs_local, n_total=N) reduces on the GPU."""
    with tempfile.TemporaryDirectory() as d:
        init_file, result_file = os.path.join(d, "init"), os.path.join(d, "res.pt")
        mp.spawn(_fewshot_worker, args=(init_file, result_file), nprocs=WORLD, join=True)
        res = torch.load(result_file)
    assert res["mine"] == [0, 2, 4]
    params = make_params()
    ref = None
    for i in range(5):
        a_w, b_w, bias = O.hypernetwork_forward(params, micro_batch(100 + i)[1], n_tokens=DIMS["n_tokens"], rank=DIMS["r"], alpha=8.0, lm_dim=DIMS["H"], mm_dim=DIMS["D"])
        flat = torch.cat([t.reshape(-1) for t in (*a_w, *b_w, *bias)]) / 5
        ref = flat if ref is None else ref + flat
    assert torch.allclose(res["mean"], ref, rtol=1e-5, atol=1e-7)


def _gradsync_worker(rank, init_file, result_file):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=WORLD)
    from dmi_b200.parallel import GradSync, Rank1FactorSync
    params = make_params()
    GA_local = 2
    # parameters in backward-availability order (generator first), small buckets so that several exist
    order = ["hypernet.generators.0.weight", "hypernet.generators.0.bias"] + [k for k in HYPER_KEYS if "generators" not in k]
    leaves = {k: torch.nn.Parameter(params[k].clone()) for k in order}
    sync = GradSync([leaves[k] for k in order], bucket_bytes=4096)
    views = {k: leaves[k].grad for k in order}            # the flat views must stay attached through backward
    sync.zero_grad()
    for j in range(GA_local):
        sync.enabled = j == GA_local - 1                   # no_sync for all but the last micro-step
        full = dict(params)
        full.update(leaves)
        x, z, dy = micro_batch(j * WORLD + rank)
        out = O.hypernet_wrapper_forward(full, x, z, n_tokens=DIMS["n_tokens"], rank=DIMS["r"], alpha=8.0, lm_dim=DIMS["H"], mm_dim=DIMS["D"])
        ((out * dy).sum() / (WORLD * GA_local)).backward()
    sync.finish()
    attached = all(leaves[k].grad.data_ptr() == views[k].data_ptr() for k in order)
    # ---- rank-1 factor exchange: dG = sum_k dw_k (x) e_k over ranks and local micro-steps ----
    g = torch.Generator().manual_seed(50 + rank)
    fs = Rank1FactorSync(24, 8, "cpu", max_terms=4)
    dense = torch.zeros(24, 8)
    dense_b = torch.zeros(24)
    for _ in range(GA_local):
        dw, e = torch.randn(24, generator=g), torch.randn(8, generator=g)
        fs.push(dw, e)
        dense += torch.outer(dw, e)
        dense_b += dw
    dist.all_reduce(dense)
    dist.all_reduce(dense_b)
    Gg, bg = torch.ones(24, 8), torch.ones(24)
    fs.apply_(Gg, bg)
    if rank == 0:
        torch.save({"grads": {k: leaves[k].grad.clone() for k in order}, "attached": attached, "n_buckets": len(sync.buckets),
                    "factor": (Gg - 1.0, dense, bg - 1.0, dense_b)}, result_file)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gradsync_buckets_no_sync_and_rank1_factors():
    """GradSync (p.grad = view into a flat bucket, hooks release buckets as autograd fills them, no_sync for the first GA_local-1
    micro-steps) reproduces the reference's accumulation over GA = world*GA_local micro-steps (train_hypernet.py:119-149), and the
    rank-1 factor all-gather gives the same generator gradient as a dense all-reduce."""
    with tempfile.TemporaryDirectory() as d:
        init_file, result_file = os.path.join(d, "init"), os.path.join(d, "res.pt")
        mp.spawn(_gradsync_worker, args=(init_file, result_file), nprocs=WORLD, join=True)
        res = torch.load(result_file)
    assert res["attached"] and res["n_buckets"] >= 2
    params = make_params()
    GA = WORLD * 2
    ref = None
    for i in range(GA):
        g = hyper_grads(params, i, 1.0 / GA)
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    for k in HYPER_KEYS:
        torch.testing.assert_close(res["grads"][k], ref[k], rtol=1e-5, atol=1e-7)
    Gg, dense, bg, dense_b = res["factor"]
    torch.testing.assert_close(Gg, dense, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bg, dense_b, rtol=1e-5, atol=1e-6)
