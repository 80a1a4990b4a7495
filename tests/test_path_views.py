"""CPU: the lazy RayBundle / RayPath views over per-step records (engine.paths_from_record)
reproduce the reference's bundle structure -- compaction at every deflection, cumulative
validity rows, ray doubling and path forking at birefringent interfaces, duplicated
hand-over bundles -- without touching a GPU (the records are CPU tensors here)."""
import numpy as np
import torch

from pyrate_b200 import engine


class _Low(object):
    def __init__(self, elem_index):
        self.elem_index = elem_index


def _record(n0, flags, split, elem_index=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    rec = engine.TraceRecord()
    rec.wave = 0.5e-3
    rec.x0 = torch.rand((3, n0), generator=g, dtype=torch.float64)
    rec.k0 = torch.rand((3, n0), generator=g, dtype=torch.float64)
    rec.e0 = torch.rand((3, n0), generator=g, dtype=torch.float64)
    n = n0
    for (s, fl) in enumerate(flags):
        rec.n_in.append(n)
        rec.hit.append(torch.rand((3, n), generator=g, dtype=torch.float64))
        rec.flags.append(torch.tensor(fl, dtype=torch.uint8))
        n_out = 2 * n if split[s] else n
        rec.k.append(torch.rand((3, n_out), generator=g, dtype=torch.float64))
        rec.e.append(torch.rand((3, n_out), generator=g, dtype=torch.float64))
        rec.n_out.append(n_out)
        rec.split.append(bool(split[s]))
        n = n_out
    rec.lowered = [_Low(0 if elem_index is None else elem_index[s]) for s in range(len(flags))]
    return rec


def test_compaction_and_validity_rows():
    rec = _record(4, [[3, 3, 1, 0], [3, 2, 0, 0]], [False, False])
    (path,) = engine.paths_from_record(rec)
    b = path.raybundles
    assert len(b) == 4 and b[0] is b[1]
    assert b[0].x.shape == (2, 3, 4)
    assert torch.equal(b[0].x[0], rec.x0) and torch.equal(b[0].x[1], rec.hit[0])
    assert b[0].valid.tolist() == [[True] * 4, [True, True, True, False]]
    assert b[0].rayID.tolist() == [0, 1, 2, 3]
    # after step 0 rays 2 (TIR) and 3 (miss) are gone
    assert b[2].rayID.tolist() == [0, 1]
    assert torch.equal(b[2].x[0], rec.hit[0][:, :2]) and torch.equal(b[2].x[1], rec.hit[1][:, :2])
    assert torch.equal(b[2].k[0], rec.k[0][:, :2]) and torch.equal(b[2].k[1], rec.k[0][:, :2])
    assert b[2].valid.tolist() == [[True, True], [True, False]]
    # ray 1 has ALIVE without HIT in step 1 only in this synthetic record: kept by the flag
    assert b[3].rayID.tolist() == [0, 1] and b[3].x.shape == (1, 3, 2)
    assert torch.equal(b[3].k[0], rec.k[1][:, :2])


def test_doubling_and_forking():
    flags = [[3, 3, 3], [3] * 3, [3] * 6]
    rec = _record(3, flags, [False, True, False])
    (path,) = engine.paths_from_record(rec, splitup=False)
    b = path.raybundles
    assert [x.x.shape[2] for x in b] == [3, 3, 3, 6, 6]
    assert b[3].splitted and not b[2].splitted
    assert b[3].rayID.tolist() == [0, 1, 2, 0, 1, 2]                 # hstack order
    assert torch.equal(b[3].x[0], torch.cat((rec.hit[1], rec.hit[1]), dim=1))
    assert torch.equal(b[3].k[0], rec.k[1])
    paths = engine.paths_from_record(rec, splitup=True)
    assert len(paths) == 2
    for (p, path) in enumerate(paths):
        b = path.raybundles
        assert [x.x.shape[2] for x in b] == [3, 3, 3, 3, 3]
        assert torch.equal(b[3].k[0], rec.k[1][:, 3 * p:3 * p + 3])   # mode p of the split
        assert torch.equal(b[3].x[1], rec.hit[2][:, 3 * p:3 * p + 3])
        assert b[4].rayID.tolist() == [0, 1, 2]


def test_element_hand_over_is_duplicated():
    rec = _record(2, [[3, 3]] * 3, [False] * 3, elem_index=[0, 0, 1])
    (path,) = engine.paths_from_record(rec)
    b = path.raybundles
    assert len(b) == 6                   # input twice, 2 bundles, hand-over twice, 1 bundle
    assert b[0] is b[1] and b[3] is b[4] and b[2] is not b[3]


def test_views_are_lazy_and_numpy_export():
    rec = _record(5, [[3] * 5], [False])
    (path,) = engine.paths_from_record(rec)
    b = path.raybundles[-1]
    assert callable(b._store["x"])       # nothing materialised yet
    d = b.numpy()
    assert isinstance(d["x"], np.ndarray) and d["x"].shape == (1, 3, 5)
    assert not callable(b._store["x"])
    # Efield on demand when E was not recorded: unit and perpendicular to k
    rec.e[0] = None
    (path,) = engine.paths_from_record(rec)
    e = path.raybundles[-1].Efield[0]
    k = path.raybundles[-1].k[0]
    assert torch.allclose((e * e).sum(0), torch.ones(5, dtype=torch.float64))
    assert float((e * k).sum(0).abs().max()) < 1e-14


def test_batch_column_views():
    """engine._column_view: the per-bundle records of a wavelength batch are zero-copy
    column slices of the one batch record, and their RayPath views behave like those of
    a stand-alone trace (compaction, rayID counted from the bundle's own first ray)."""
    flags = [[3, 3, 1, 3, 3, 3, 0, 3, 3], [3, 3, 0, 3, 2, 3, 0, 3, 3]]
    rec = _record(9, flags, [False, False])
    lows = [[_Low(0), _Low(0)] for _ in range(3)]
    ends = [2, 7, 9]
    lo = 0
    for (w, hi) in enumerate(ends):
        sub = engine._column_view(rec, lo, hi, lows[w], 0.4e-3 + 0.1e-3 * w)
        assert sub.hit[1].data_ptr() == rec.hit[1].data_ptr() + 8 * lo
        assert sub.n_in == [hi - lo] * 2 and sub.wave == 0.4e-3 + 0.1e-3 * w
        (path,) = engine.paths_from_record(sub)
        b = path.raybundles
        assert len(b) == 4 and b[0] is b[1] and b[0].wave == sub.wave
        assert torch.equal(b[0].x[0], rec.x0[:, lo:hi]) and torch.equal(b[0].x[1], rec.hit[0][:, lo:hi])
        alive0 = (torch.tensor(flags[0][lo:hi]) & 2) != 0
        assert b[2].rayID.tolist() == torch.arange(hi - lo)[alive0].tolist()
        assert torch.equal(b[2].k[0], rec.k[0][:, lo:hi][:, alive0])
        alive1 = alive0 & ((torch.tensor(flags[1][lo:hi]) & 2) != 0)
        assert b[3].rayID.tolist() == torch.arange(hi - lo)[alive1].tolist()
        assert torch.equal(b[3].x[0], rec.hit[1][:, lo:hi][:, alive1])
        lo = hi


def test_device_bundle_layout_rules():
    """engine.device_bundle (device = "cpu" here: the rules are device independent):
    rows that already live on the target device with ONE common row stride are used in
    place; anything else lands in row-padded buffers (ld a multiple of 16 doubles, the
    layout of the TMA-staged input path), complex promotion included."""
    rng = np.random.default_rng(2)
    (x, k, e) = (torch.from_numpy(rng.random((3, 37))) for _ in range(3))
    (a, b, c) = engine.device_bundle(x, k, e, device="cpu")
    assert (a.data_ptr(), b.data_ptr(), c.data_ptr()) == (x.data_ptr(), k.data_ptr(), e.data_ptr())
    # one operand with a different row stride: everything is re-laid with a common ld
    wide = torch.zeros((3, 64), dtype=torch.float64)
    wide[:, :37] = k
    (a, b, c) = engine.device_bundle(x, wide[:, :37], e, device="cpu")
    assert {t.stride(0) for t in (a, b, c)} == {48} and all(t.shape == (3, 37) for t in (a, b, c))
    assert torch.equal(a, x) and torch.equal(b, k) and torch.equal(c, e)
    assert a.data_ptr() != x.data_ptr()
    # padding columns are zero (the vector / TMA paths may read them)
    full = torch.as_strided(a, (3, 48), (48, 1))
    assert float(full[:, 37:].abs().max()) == 0.0
    # complex promotion, E = None passes through, like_ld is honoured
    (a, b, c) = engine.device_bundle(x, k, None, device="cpu", complex_=True)
    assert c is None and b.dtype == torch.complex128 and a.dtype == torch.float64
    assert a.stride(0) == b.stride(0) == 48 and torch.equal(b.real, k)
    (a, b, c) = engine.device_bundle(x, k, e, device="cpu", like_ld=80)
    assert {t.stride(0) for t in (a, b, c)} == {80}
    # _padded: the per-launch view of one (3, n) array
    (buf, ld) = engine._padded(x)
    assert ld == 48 and buf.shape == (3, 48) and torch.equal(buf[:, :37], x)
    (buf, ld) = engine._padded(x, pad=False)
    assert ld == 37 and buf.data_ptr() == x.data_ptr()
    (buf, ld) = engine._padded(k, True)
    assert buf.shape == (3, 48, 2) and torch.equal(buf[:, :37, 0], k) and float(buf[..., 1].abs().max()) == 0.0
