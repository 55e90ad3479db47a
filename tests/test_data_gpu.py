"""EmbeddingStore (dmi_gather_rows) against the oracle's literal restatement of the reference collate + get_embeddings.
Row / column gathers are index work (bit-exact); mean subtraction and normalisation are fp32 (1e-5, here 1e-6)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def make_items(n, d, seed):
    rs = np.random.RandomState(seed)
    return {f"id{i:05d}": {"caption": f"c{i}", "emb": rs.randn(d).astype(np.float32).tolist()} for i in range(n)}


@pytest.mark.parametrize("n,d,sel,use_mean,norm", [(50, 768, None, False, True), (37, 1024, 768, True, True), (9, 520, 100, True, False),
                                                   (300, 4096, 768, False, True), (5, 64, None, True, True)])
def test_gather_matches_collate(n, d, sel, use_mean, norm):
    from dmi_b200.data import EmbeddingStore
    items = make_items(n, d, seed=n + d)
    rs = np.random.RandomState(1)
    selected = None if sel is None else np.sort(rs.choice(d, sel, replace=False))
    d_out = d if sel is None else sel
    mean = torch.from_numpy(rs.randn(d_out).astype(np.float32) * 0.1) if use_mean else None
    store = EmbeddingStore.from_items(items, selected_features=selected, mean=mean)
    keys = list(items.keys())
    pick = [keys[i] for i in rs.randint(0, n, size=23)]
    ref = O.collate_embeddings([items[k] for k in pick], selected_features=selected, emb_mean=mean, normalize=norm)
    bf = torch.full((23, d_out + 8), 7.0, dtype=torch.bfloat16, device="cuda")
    got = store.gather(store.rows(pick), normalize=norm, out_bf16=bf, check=True)
    if not use_mean and not norm:
        assert torch.equal(got.cpu(), ref)
    assert torch.allclose(got.cpu(), ref, rtol=1e-6, atol=1e-7)
    assert torch.equal(bf[:, :d_out].cpu(), got.to(torch.bfloat16).cpu())            # bf16 copy = round-to-nearest of the fp32 result
    assert bool((bf[:, d_out:] == 7.0).all())


def test_pure_gather_is_bit_exact_and_bf16_store():
    from dmi_b200.data import EmbeddingStore
    g = torch.Generator(device="cuda").manual_seed(0)
    table = torch.randn(1000, 768, device="cuda", generator=g)
    idx = torch.randint(0, 1000, (256,), device="cuda", generator=g)
    sel = np.arange(767, -1, -3)
    out = EmbeddingStore(table, selected_features=sel).gather(idx, normalize=False)
    assert torch.equal(out, table[idx][:, torch.from_numpy(sel.copy()).cuda()])
    tb = table.to(torch.bfloat16)
    outb = EmbeddingStore(tb).gather(idx, normalize=False)
    assert torch.equal(outb, tb[idx].float())


def test_out_of_range_index_is_reported():
    from dmi_b200.data import EmbeddingStore
    store = EmbeddingStore(torch.ones(10, 64, device="cuda"))
    with pytest.raises(IndexError):
        store.gather(torch.tensor([1, 10], device="cuda"), check=True)
    store.gather(torch.tensor([1, 9], device="cuda"), check=True)


def test_out_of_range_index_without_check_poisons_the_row_and_raises_later():
    """with check=False a bad sample index must not leave an uninitialised row or go unnoticed: the row is NaN and the
    deferred flag raises at the next call (or at check()), without a synchronisation on the good path."""
    from dmi_b200.data import EmbeddingStore
    store = EmbeddingStore(torch.ones(10, 64, device="cuda"))
    out = store.gather(torch.tensor([1, 10, 3], device="cuda"), normalize=False)          # no exception here: nothing is synchronised
    torch.cuda.synchronize()
    assert torch.isnan(out[1]).all() and bool((out[0] == 1).all()) and bool((out[2] == 1).all())
    with pytest.raises(IndexError):
        store.gather(torch.tensor([0], device="cuda"), normalize=False)                  # the earlier launch's flag surfaces here
    store.gather(torch.tensor([0, 9], device="cuda"), normalize=False)                     # flag was reset
    store.check()


def test_splice_out_of_range_token_id_raises_at_the_next_check():
    from dmi_b200 import ops
    from dmi_b200.model.mmmodel import splice_prefix
    table = torch.randn(50, 64, device="cuda").to(torch.bfloat16)
    proj = torch.randn(2, 64, device="cuda")
    ops.check_device_errors()
    splice_prefix(proj, table, torch.tensor([[1, 2], [3, 50]], device="cuda"))            # id 50 is outside the table
    with pytest.raises(IndexError):
        ops.check_device_errors()
    splice_prefix(proj, table, torch.tensor([[1, 2], [3, 49]], device="cuda"))
    ops.check_device_errors()
