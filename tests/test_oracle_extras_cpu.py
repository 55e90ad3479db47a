"""CPU checks that pin the oracle's restatements of the section-8f rows to the third-party code the reference actually calls:
torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ (train_hypernet.py:148-149), scipy.stats.ortho_group.rvs (:57), and the
literal collate + normalise sequence (data/base.py:222-232, model_utils.py:47-62)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O


@pytest.mark.parametrize("seed,wd,clip", [(0, 0.0, 1.0), (1, 5e-6, 0.05), (2, 0.1, 100.0)])
def test_adamw_and_clip_restatement_vs_torch(seed, wd, clip):
    g = torch.Generator().manual_seed(seed)
    shapes = [(17, 5), (3,), (64, 9)]
    hp = dict(lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd)
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    opt = torch.optim.AdamW(params, **hp)
    p = [q.detach().clone() for q in params]
    m = [torch.zeros_like(q) for q in p]
    v = [torch.zeros_like(q) for q in p]
    for t in range(4):
        grads = [torch.randn(*s, generator=g) * (10.0 if t % 2 else 0.01) for s in shapes]
        for q, gr in zip(params, grads):
            q.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(params, clip)
        opt.step()
        clipped, total_o = O.clip_grad_norm(grads, clip)
        assert torch.allclose(total_o, total, rtol=1e-6)
        for i in range(len(p)):
            assert torch.allclose(clipped[i], params[i].grad, rtol=1e-6, atol=1e-12)
            p[i], m[i], v[i] = O.adamw_step(p[i], clipped[i], m[i], v[i], t + 1, **hp)
            assert torch.allclose(p[i], params[i].detach(), rtol=1e-6, atol=1e-7)
            assert torch.allclose(m[i], opt.state[params[i]]["exp_avg"], rtol=1e-5, atol=1e-9)
            assert torch.allclose(v[i], opt.state[params[i]]["exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_haar_construction_is_orthogonal_and_has_the_moments_of_scipys_draw():
    from scipy.stats import ortho_group
    n, draws = 16, 600
    ours = np.stack([O.haar_from_gaussian(np.random.RandomState(s).randn(n, n)) for s in range(draws)])
    host = ortho_group.rvs(n, size=draws, random_state=np.random.RandomState(123))
    eye = np.eye(n)
    assert np.abs(np.einsum("bij,bik->bjk", ours, ours) - eye).max() < 1e-12
    for Q in (ours, host):
        tr = np.trace(Q, axis1=1, axis2=2)
        assert abs(tr.mean()) < 0.15 and 0.8 < tr.var() < 1.25                 # E tr = 0, Var tr = 1 on O(n)
        assert abs((Q[:, 0, 0] ** 2).mean() * n - 1.0) < 0.15                   # E Q_00^2 = 1/n
        assert abs((Q[:, 0, 0] * Q[:, 1, 1]).mean()) < 0.02                      # distinct entries uncorrelated
        dets = np.linalg.det(Q)
        assert np.allclose(np.abs(dets), 1.0) and 0.4 < (dets > 0).mean() < 0.6   # both components of O(n), equally often
    # the first column is uniform on the sphere in both: compare a tail probability
    assert abs((np.abs(ours[:, 0, 0]) > 0.4).mean() - (np.abs(host[:, 0, 0]) > 0.4).mean()) < 0.08


def test_haar_construction_reproduces_householder_qr_with_sign_fix():
    """For a given Gaussian matrix Z the oracle's reflectors, fed with the successive reduced columns of the Householder QR of Z,
    give exactly scipy's Q * sign(diag R) -- i.e. the construction IS ortho_group.rvs's algorithm, minus the matrix."""
    rs = np.random.RandomState(5)
    n = 12
    Z = rs.randn(n, n)
    # run Householder QR by hand, recording the reduced column each reflector sees
    A = Z.copy()
    gauss = np.zeros((n, n))
    for k in range(n):
        x = A[k:, k].copy()
        gauss[k, k:] = x
        sgn = 1.0 if x[0] >= 0 else -1.0
        v = x.copy()
        v[0] += sgn * np.linalg.norm(x)
        if v @ v > 0:
            A[k:, :] -= np.outer(v, (2.0 / (v @ v)) * (v @ A[k:, :]))
    q, r = np.linalg.qr(Z)
    ref = q * np.sign(np.diag(r))[None, :]
    got = O.haar_from_gaussian(gauss)
    assert np.abs(got - ref).max() < 1e-10


def test_collate_embeddings_matches_manual_numpy():
    rs = np.random.RandomState(0)
    items = [{"emb": rs.randn(40).astype(np.float32).tolist()} for _ in range(7)]
    sel = np.array([3, 1, 39, 20, 5])
    mean = torch.from_numpy(rs.randn(5).astype(np.float32))
    got = O.collate_embeddings(items, selected_features=sel, emb_mean=mean, normalize=True)
    raw = np.asarray([it["emb"] for it in items], dtype=np.float32)[:, sel] - mean.numpy()
    ref = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    assert np.allclose(got.numpy(), ref, rtol=1e-6, atol=1e-7)
