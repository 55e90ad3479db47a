"""On-device Haar orthogonal matrix (dmi_haar_orthogonal) against the float64 oracle restatement on the same Gaussian
samples (1e-5, fp32) and through size-independent properties: orthogonality, isometry of the rotation, and the Haar moments
E[tr Q] = 0, E[tr Q^2-ish] ... that scipy.stats.ortho_group.rvs (the reference's draw, train_hypernet.py:57) satisfies as well."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 7, 64, 65, 200, 768])
def test_matches_oracle_on_same_gaussian_samples(n):
    from dmi_b200 import augment as A
    g = torch.Generator().manual_seed(n)
    gauss = torch.randn(n, n, generator=g)
    Q = A.get_rotation_matrix_device(n, "cuda", gauss=gauss.cuda())
    ref = O.haar_from_gaussian(gauss.numpy())
    assert np.abs(Q.cpu().numpy().astype(np.float64) - ref).max() < 1e-5


@pytest.mark.parametrize("n", [768, 1024, 2048])
def test_orthogonal_and_isometric_at_encoder_widths(n):
    from dmi_b200 import augment as A
    gen = torch.Generator(device="cuda").manual_seed(n)
    Q = A.get_rotation_matrix_device(n, "cuda", generator=gen).double()
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert float((Q.T @ Q - eye).abs().max()) < 2e-5
    x = torch.randn(64, n, dtype=torch.float64, device="cuda")
    assert torch.allclose((x @ Q).norm(dim=1), x.norm(dim=1), rtol=1e-5)           # the augmentation is an isometry
    assert abs(abs(float(torch.linalg.det(Q))) - 1.0) < 1e-3


def test_haar_moments_match_the_host_draw():
    """400 draws at n = 24: trace mean 0 / variance 1, E[Q_00^2] = 1/n, det = +-1 equally often -- the same moments the
    reference's scipy.stats.ortho_group.rvs produces (checked side by side on the host)."""
    from scipy.stats import ortho_group
    from dmi_b200 import augment as A
    n, draws = 24, 400
    gen = torch.Generator(device="cuda").manual_seed(7)
    tr, q00, dets = [], [], []
    for _ in range(draws):
        Q = A.get_rotation_matrix_device(n, "cuda", generator=gen).double()
        tr.append(float(Q.trace())); q00.append(float(Q[0, 0] ** 2)); dets.append(float(torch.linalg.det(Q)) > 0)
    host = ortho_group.rvs(n, size=draws, random_state=np.random.RandomState(3))
    tr_h = np.trace(host, axis1=1, axis2=2)
    assert abs(np.mean(tr)) < 0.2 and abs(np.mean(tr_h)) < 0.2
    assert 0.75 < np.var(tr) < 1.3 and 0.75 < np.var(tr_h) < 1.3
    assert abs(np.mean(q00) * n - 1.0) < 0.2
    assert 0.38 < np.mean(dets) < 0.62
