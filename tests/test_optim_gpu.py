"""Fused clip-grad-norm + AdamW (dmi_b200.optim, C ABI dmi_grad_sqnorm / dmi_grad_clip / dmi_adamw_step) against the golden
sequence produced by the reference's own calls (torch.nn.utils.clip_grad_norm_ + optim.AdamW.step(), train_hypernet.py:148-149)
and against torch's optimizer at hypernet-like sizes.  fp32: 1e-5 relative (north_star); untouched parameters bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def close(a, b, rtol=1e-5, atol=1e-7):
    return torch.allclose(a.detach().cpu(), b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("fused", [False, True])
def test_golden_sequence(golden_dir, fused):
    from dmi_b200.optim import FusedAdamW, clip_grad_norm_
    d = load(golden_dir, "optimizer_adamw_clip")
    n, steps, skip = int(d["n_params"]), int(d["steps"]), int(d["no_grad_index"])
    params = [torch.nn.Parameter(d[f"p0/{i}"].cuda()) for i in range(n)]
    opt = FusedAdamW(params=params, lr=float(d["lr"]), betas=(float(d["beta1"]), float(d["beta2"])), eps=float(d["eps"]),
                     weight_decay=float(d["weight_decay"]))
    mx = float(d["max_grad_norm"])
    for t in range(steps):
        for i, p in enumerate(params):
            p.grad = None if i == skip else d[f"g{t}/{i}"].cuda()
        if fused:
            opt.step(max_grad_norm=mx, write_clipped_grads=True)
            total = opt.last_grad_norm
        else:
            total = clip_grad_norm_(params, mx)            # the reference's two calls, one after the other
            opt.step()
        assert close(total, d[f"norm{t}"], rtol=1e-5)
        for i, p in enumerate(params):
            if i == skip:
                assert torch.equal(p.detach().cpu(), d[f"p{t + 1}/{i}"])
                continue
            assert close(p.grad, d[f"gclip{t}/{i}"], atol=1e-10), (t, i)
            assert close(p, d[f"p{t + 1}/{i}"]), (t, i)
            assert close(opt.state[p]["exp_avg"], d[f"m{t + 1}/{i}"], atol=1e-9)
            assert close(opt.state[p]["exp_avg_sq"], d[f"v{t + 1}/{i}"], atol=1e-12)


def test_state_dict_roundtrip_with_torch_adamw():
    """checkpoints are interchangeable with optim.AdamW: same state keys, and a torch optimizer continues from our state"""
    from dmi_b200.optim import FusedAdamW
    g = torch.Generator(device="cuda").manual_seed(0)
    hp = dict(lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01)
    w0 = torch.randn(257, 31, device="cuda", generator=g)
    pa, pb = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w0.clone())
    ours, ref = FusedAdamW([pa], **hp), torch.optim.AdamW([pb], **hp)
    for _ in range(2):
        gr = torch.randn(257, 31, device="cuda", generator=g)
        pa.grad, pb.grad = gr.clone(), gr.clone()
        ours.step(); ref.step()
    sd = ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    ref2 = torch.optim.AdamW([pa], **hp)
    ref2.load_state_dict(sd)
    gr = torch.randn(257, 31, device="cuda", generator=g)
    pa.grad, pb.grad = gr.clone(), gr.clone()
    ref2.step(); ref.step()
    assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7)


def test_hypernet_sized_step_matches_torch():
    """generator-sized tensors (92160 x 768 and friends), unaligned tails, 3 fused steps vs clip_grad_norm_ + optim.AdamW"""
    from dmi_b200.optim import FusedAdamW
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(92160, 768), (92160,), (768, 768), (768,), (2, 768), (1000003,)]
    hp = dict(lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=5e-6)
    base = [torch.randn(*s, device="cuda", generator=g) * 0.05 for s in shapes]
    pa = [torch.nn.Parameter(b.clone()) for b in base]
    pb = [torch.nn.Parameter(b.clone()) for b in base]
    ours, ref = FusedAdamW(pa, **hp), torch.optim.AdamW(pb, **hp)
    for t in range(3):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (1e-3 if t == 1 else 1e-5)
            a.grad, b.grad = gr.clone(), gr.clone()
        ours.step(max_grad_norm=1.0)
        total = torch.nn.utils.clip_grad_norm_(pb, 1.0)
        ref.step()
        assert torch.allclose(ours.last_grad_norm, total, rtol=1e-4)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
        assert torch.allclose(ours.state[a]["exp_avg_sq"], ref.state[b]["exp_avg_sq"], rtol=1e-4, atol=1e-14)


def test_cpu_parameters_raise():
    from dmi_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdamW([p], lr=1e-3).step()


def test_allreduce_oneshot_kernel_world_1():
    """dmi_allreduce_oneshot with a single rank (its own buffer as the only peer): out = scale * in, the flag protocol advances by two
    per call and survives repeated calls on the same slot.  The multi-rank path (peer-mapped symmetric memory, NVLS multicast) is
    checked against NCCL inside bench.py on the multi-GPU box (`dp_reduce.max_abs_diff_vs_nccl`)."""
    import ctypes as C

    from dmi_b200 import _lib
    lib = _lib.load()
    n = 225280                                     # the bench's flat adapter-gradient buffer (0.9 MB)
    g = torch.Generator(device="cuda").manual_seed(0)
    inp = torch.randn(n, device="cuda", generator=g)
    out = torch.zeros(n, device="cuda")
    flags = torch.zeros(int(lib.dmi_allreduce_flag_words()), dtype=torch.int32, device="cuda")
    arr = C.c_void_p * 1
    for epoch in (1, 2, 3):
        rc = lib.dmi_allreduce_oneshot(arr(inp.data_ptr()), arr(flags.data_ptr()), None, 0, 1, C.c_void_p(out.data_ptr()), n, 0.5, epoch,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "dmi_allreduce_oneshot")
        torch.cuda.synchronize()
        assert torch.equal(out, inp * 0.5)
        inp.add_(1.0)
    assert int(flags.max().item()) == 6 and int(flags[0].item()) == 6
    with pytest.raises(RuntimeError):              # in-place is refused: peers may still be reading the input
        _lib.check(lib.dmi_allreduce_oneshot(arr(inp.data_ptr()), arr(flags.data_ptr()), None, 0, 1, C.c_void_p(inp.data_ptr()), n, 1.0, 4,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dmi_allreduce_oneshot")
