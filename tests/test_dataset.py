"""Windowed trajectory dataset: oracle vs the reference's golden batches, host logic of
DeviceWindowDataset, and (GPU) the gather kernel through the C-ABI."""
import os

import numpy as np
import pytest
import torch

from beso_b200.dataset import DeviceWindowDataset
from oracle import ref_import
from oracle import window_oracle as WO
from oracle.make_golden import WINDOW_MODES, reference_window_dataset, window_fixture_data

GOLD = os.path.join(os.path.dirname(__file__), "golden", "windows.npz")
KEYS = ("observation", "action", "goal_observation")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_fixture_inputs_are_reproducible(gold):
    obs, act, lens = window_fixture_data()
    assert np.array_equal(obs, gold["obs"]) and np.array_equal(act, gold["act"]) and np.array_equal(lens, gold["lens"])


@pytest.mark.parametrize("mode", list(WINDOW_MODES))
def test_oracle_matches_reference_golden(gold, mode):
    kw = dict(WINDOW_MODES[mode])
    window = kw.pop("window")
    assert np.array_equal(np.array(WO.slices(gold["lens"], window)), gold[f"{mode}::slices"])
    np.random.seed(1234)
    b = WO.batch(gold["obs"], gold["act"], gold["lens"], gold[f"{mode}::order"], window, **kw)
    for k in KEYS:
        if f"{mode}::{k}" in gold:
            assert np.array_equal(b[k], gold[f"{mode}::{k}"]), (mode, k)   # bit-exact: pure indexing
    if mode == "future":   # the fixture exercises both the sampled window and the zeros placeholder
        z = (b["goal_observation"] == 0).all(axis=(1, 2))
        assert z.any() and not z.all()


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("mode", list(WINDOW_MODES))
def test_oracle_matches_live_reference(mode):
    tl = ref_import.load_trajectory_loader()
    rs = np.random.RandomState(3)
    lens = rs.randint(2, 31, size=9)
    obs, act = rs.randn(9, 30, 4).astype(np.float32), rs.randn(9, 30, 2).astype(np.float32)
    kw = dict(WINDOW_MODES[mode])
    ds = reference_window_dataset(tl, obs, act, lens, **kw)
    order = rs.permutation(len(ds))
    np.random.seed(77)
    ref = [ds[int(i)] for i in order]
    window = kw.pop("window")
    np.random.seed(77)
    b = WO.batch(obs, act, lens, order, window, **kw)
    for k in b:
        assert np.array_equal(b[k], np.stack([np.asarray(r[k]) for r in ref])), (mode, k)


@pytest.mark.parametrize("mode", list(WINDOW_MODES))
def test_host_logic_matches_golden(gold, mode):
    """Slice enumeration and the future-window starts (no kernel call: tensors stay on the CPU)."""
    kw = dict(WINDOW_MODES[mode])
    ds = DeviceWindowDataset(gold["obs"], gold["act"], gold["lens"], device="cpu", **kw)
    sl = gold[f"{mode}::slices"]
    assert len(ds) == sl.shape[0]
    assert np.array_equal(ds.slice_traj, sl[:, 0]) and np.array_equal(ds.slice_start, sl[:, 1])
    assert ds.get_seq_length(0) == kw["window"] + (kw.get("future_seq_len") or 0)
    if not ds.future_conditional:
        return
    order = gold[f"{mode}::order"]
    np.random.seed(1234)
    gs = ds.goal_starts(order)
    G = kw["future_seq_len"]
    want = gold[f"{mode}::goal_observation"]
    for n, idx in enumerate(order):
        got = np.zeros((G, gold["obs"].shape[2]), np.float32) if gs[n] < 0 else gold["obs"][sl[idx, 0], gs[n]:gs[n] + G]
        assert np.array_equal(got, want[n])
    # an explicit RandomState draws the same starts as the global generator with the same seed
    assert np.array_equal(ds.goal_starts(order, np.random.RandomState(1234)), gs)


def test_constructor_errors():
    obs, act, lens = window_fixture_data()
    with pytest.raises(AssertionError):
        DeviceWindowDataset(obs, act, lens, window=4, future_conditional=True, device="cpu")
    with pytest.raises(ValueError):
        DeviceWindowDataset(obs, act[:, :5], lens, window=4, device="cpu")
    with pytest.raises(ValueError):
        DeviceWindowDataset(obs, act, lens + 100, window=4, device="cpu")
    ds = DeviceWindowDataset(obs, act, lens, window=4, device="cpu")
    with pytest.raises(IndexError):
        ds.get_batch([len(ds)])
    with pytest.raises(ValueError):
        ds.get_batch([])
    assert len(DeviceWindowDataset(obs, act, lens, window=50, device="cpu")) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(WINDOW_MODES))
def test_gather_matches_golden_gpu(gold, mode):
    ds = DeviceWindowDataset(gold["obs"], gold["act"], gold["lens"], device="cuda", **WINDOW_MODES[mode])
    np.random.seed(1234)
    b = ds.get_batch(gold[f"{mode}::order"])
    for k in KEYS:
        if f"{mode}::{k}" in gold:
            assert np.array_equal(b[k].cpu().numpy(), gold[f"{mode}::{k}"]), (mode, k)
        else:
            assert k not in b


@pytest.mark.gpu
def test_gather_large_matches_oracle_gpu():
    """Training-sized batch (1024 windows of the kitchen shapes: 60-dim observations, 9-dim actions)."""
    rs = np.random.RandomState(0)
    N, t_max = 64, 400
    lens = rs.randint(30, t_max + 1, size=N)
    obs, act = rs.randn(N, t_max, 60).astype(np.float32), rs.randn(N, t_max, 9).astype(np.float32)
    kw = dict(window=5, future_conditional=True, min_future_sep=3, future_seq_len=2)
    ds = DeviceWindowDataset(obs, act, lens, device="cuda", **kw)
    idx = rs.randint(0, len(ds), size=1024)
    from beso_b200 import _lib
    n0 = _lib.lib().beso_kernel_launches()
    b = ds.get_batch(idx, rng=np.random.RandomState(9))
    assert _lib.lib().beso_kernel_launches() - n0 == 1
    window = kw.pop("window")
    want = WO.batch(obs, act, lens, idx, window, rng=np.random.RandomState(9), **kw)
    for k in KEYS:
        assert np.array_equal(b[k].cpu().numpy(), want[k]), k
    # epoch iterator: every window exactly once, batches of the requested size
    seen = 0
    for bb in ds.batches(4096, shuffle=True, generator=torch.Generator().manual_seed(0)):
        seen += bb["observation"].shape[0]
        assert bb["goal_observation"].shape[1:] == (2, 60)
    assert seen == len(ds)


@pytest.mark.gpu
def test_gather_rejects_bad_arguments_gpu():
    obs, act, lens = window_fixture_data()
    ds = DeviceWindowDataset(obs, act, lens, device="cuda", **WINDOW_MODES["future"])
    with pytest.raises(ValueError):
        ds.get_batch([0, 1], goal_start=[0, 19])     # 19 + 2 > t_max
    with pytest.raises(ValueError):
        ds.get_batch([0, 1], goal_start=[0])


# ---- scalers (beso/networks/scaler/scaler_class.py) ------------------------------------------------------------------
def _scaler_data(dtype=np.float32):
    from oracle.make_golden import scaler_fixture_data
    x, y = scaler_fixture_data()
    return x.astype(dtype), y.astype(dtype)


@pytest.mark.parametrize("cls", ["Scaler", "MinMaxScaler"])
@pytest.mark.parametrize("scale_data", [True, False])
def test_scaler_mirror_matches_reference_golden(cls, scale_data):
    """Fixture made with the reference's classes (python -m oracle.make_golden scalers): bit-exact statistics and
    method outputs, also where /root/reference is absent."""
    from beso_b200 import scaler as S
    from oracle.make_golden import scaler_fixture_data
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "scalers.npz"))
    x, y = scaler_fixture_data()
    sc = getattr(S, cls)(x, y, scale_data, "cpu")
    tag = f"{cls}::{int(scale_data)}::"
    for name in ("x_mean", "x_std", "x_max", "x_min", "y_min", "y_max", "y_bounds_tensor", "x_bounds_tensor"):
        assert np.array_equal(getattr(sc, name).numpy(), z[tag + name]), name
    xs, ys = torch.from_numpy(x[:4]), torch.from_numpy(y[:4])
    for fn, arg in (("scale_input", xs), ("scale_output", ys), ("inverse_scale_input", xs), ("inverse_scale_output", ys),
                    ("clip_action", ys * 3)):
        assert np.array_equal(getattr(sc, fn)(arg.clone()).numpy(), z[tag + fn]), fn


def _load_ref_scalers():
    import importlib.util
    path = os.path.join(ref_import.REF_ROOT, "beso", "networks", "scaler", "scaler_class.py")
    spec = importlib.util.spec_from_file_location("_beso_ref_scaler", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("cls", ["Scaler", "MinMaxScaler"])
@pytest.mark.parametrize("scale_data", [True, False])
def test_scaler_mirror_matches_live_reference(cls, scale_data):
    from beso_b200 import scaler as S
    ref = _load_ref_scalers()
    x, y = _scaler_data()
    a, b = getattr(S, cls)(x, y, scale_data, "cpu"), getattr(ref, cls)(x, y, scale_data, "cpu")
    for name in ("x_mean", "x_std", "x_max", "x_min", "y_min", "y_max", "y_bounds_tensor", "x_bounds_tensor", "tensor_y_bounds"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.y_bounds, b.y_bounds) and np.array_equal(a.x_bounds, b.x_bounds)
    xs, ys = torch.from_numpy(x[:4]), torch.from_numpy(y[:4])
    for fn, arg in (("scale_input", xs), ("scale_output", ys), ("inverse_scale_input", xs), ("inverse_scale_output", ys),
                    ("clip_action", ys * 3)):
        assert torch.equal(getattr(a, fn)(arg.clone()), getattr(b, fn)(arg.clone())), fn


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_scaler_special_goal_shapes_match_live_reference():
    from beso_b200 import scaler as S
    ref = _load_ref_scalers()
    rs = np.random.RandomState(2)
    for xdim, gdim in ((16, 4), (30, 7)):
        x, y = rs.randn(50, xdim).astype(np.float32), rs.randn(50, 2).astype(np.float32)
        g = torch.from_numpy(rs.randn(5, 1, gdim).astype(np.float32))
        assert torch.equal(S.Scaler(x, y, True, "cpu").scale_input(g), ref.Scaler(x, y, True, "cpu").scale_input(g))


def test_scaler_tables_reproduce_scaling_bit_for_bit():
    """The (sub, div, mul, add) rows handed to the gather kernel, evaluated step by step in fp32, equal the methods."""
    from beso_b200 import scaler as S
    x, y = _scaler_data()
    xs, ys = torch.from_numpy(x[:3]), torch.from_numpy(y[:3])
    for cls in (S.Scaler, S.MinMaxScaler):
        sc = cls(x, y, True, "cpu")
        ot, at = sc.gather_tables()
        assert torch.equal(((xs - ot[0]) / ot[1]) * ot[2] + ot[3], sc.scale_input(xs))
        assert torch.equal(((ys - at[0]) / at[1]) * at[2] + at[3], sc.scale_output(ys))
        assert cls(x, y, False, "cpu").gather_tables() == (None, None)
    with pytest.raises(TypeError):
        S.Scaler(*_scaler_data(np.float64), True, "cpu").gather_tables()
    with pytest.raises(ValueError):
        S.Scaler(np.zeros(5, np.float32), np.zeros(5, np.float32), True, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("cls", ["Scaler", "MinMaxScaler"])
def test_gather_with_fused_scaling_gpu(gold, cls):
    """The kernel's fused scaling equals scale_input / scale_output applied to the unscaled batch, bit for bit
    (including the scaled zeros placeholder)."""
    from beso_b200 import scaler as S
    sc = getattr(S, cls)(gold["obs"], gold["act"], True, "cuda")
    kw = WINDOW_MODES["future"]
    raw = DeviceWindowDataset(gold["obs"], gold["act"], gold["lens"], device="cuda", **kw)
    fused = DeviceWindowDataset(gold["obs"], gold["act"], gold["lens"], device="cuda", scaler=sc, **kw)
    idx = gold["future::order"]
    a = raw.get_batch(idx, rng=np.random.RandomState(4))
    b = fused.get_batch(idx, rng=np.random.RandomState(4))
    assert b["scaled"] is True and "scaled" not in a
    assert torch.equal(b["observation"], sc.scale_input(a["observation"]))
    assert torch.equal(b["goal_observation"], sc.scale_input(a["goal_observation"]))
    assert torch.equal(b["action"], sc.scale_output(a["action"]))


@pytest.mark.gpu
def test_train_step_on_gathered_batches_gpu():
    """The training slice end to end: window gather with fused scaling -> BesoAgent.train_step.  With the same seeds the
    losses are identical to train_step fed the unscaled batch with the scaler attached to the agent."""
    from beso_b200 import K256
    from beso_b200.denoiser import build_denoiser
    from beso_b200 import scaler as S
    from beso_b200.agent import BesoAgent
    from beso_b200.synth import synthetic_state_dict
    cfg = K256
    rs = np.random.RandomState(8)
    N, t_max = 16, 120
    lens = rs.randint(40, t_max + 1, size=N)
    obs = (rs.randn(N, t_max, cfg.obs_dim) * 2 + 1).astype(np.float32)
    act = (rs.randn(N, t_max, cfg.act_dim) * 0.5).astype(np.float32)
    sc = S.Scaler(obs, act, True, "cuda")
    kw = dict(window=cfg.window, future_conditional=True, min_future_sep=2, future_seq_len=cfg.goal_len)
    fused = DeviceWindowDataset(obs, act, lens, device="cuda", scaler=sc, **kw)
    raw = DeviceWindowDataset(obs, act, lens, device="cuda", **kw)
    sd = synthetic_state_dict(cfg, 61)
    losses = []
    for ds, agent_scaler in ((fused, None), (raw, sc)):
        m = build_denoiser(cfg, "cuda", mode="precise", state_dict=sd)
        agent = BesoAgent(m, device="cuda", sigma_min=0.005, sigma_max=1.0, window_size=cfg.window, scaler=agent_scaler)
        agent.configure_training(lr=1e-3)
        run = []
        gen = torch.Generator().manual_seed(3)
        host_rng = np.random.RandomState(3)
        for step, batch in enumerate(ds.batches(256, shuffle=True, drop_last=True, generator=gen, rng=host_rng)):
            torch.manual_seed(200 + step)
            run.append(agent.train_step(batch))
            if step == 3:
                break
        losses.append(run)
    assert len(losses[0]) == 4 and losses[0] == losses[1]
    assert all(np.isfinite(losses[0]))


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("shuffle", [True, False])
def test_epoch_order_shards_like_distributed_sampler(gold, world, shuffle):
    """Data-parallel sharding of the windows follows torch's DistributedSampler (same permutation on every rank,
    wrap-around padding, rank::world stride): disjoint equal shards that cover every window."""
    from torch.utils.data import DistributedSampler
    ds = DeviceWindowDataset(gold["obs"], gold["act"], gold["lens"], device="cpu", **WINDOW_MODES["plain"])
    n = len(ds)
    seen = []
    for rank in range(world):
        ref = DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=shuffle, seed=17)
        got = ds.epoch_order(shuffle, torch.Generator().manual_seed(17), rank, world)
        assert got.tolist() == list(iter(ref))
        seen.append(got)
    assert len({len(s) for s in seen}) == 1
    assert set(np.concatenate(seen).tolist()) == set(range(n))
    with pytest.raises(ValueError):
        ds.epoch_order(True, None, world, world)


def test_vectorised_goal_draws_equal_successive_scalar_draws(gold):
    """goal_starts makes ONE randint call with array bounds; the reference makes one scalar call per item.  Same
    values, same generator state afterwards (legacy RandomState and the np.random module functions)."""
    kw = WINDOW_MODES["future"]
    ds = DeviceWindowDataset(gold["obs"], gold["act"], gold["lens"], device="cpu", **kw)
    order = np.random.RandomState(1).randint(0, len(ds), size=500)
    want, ref_rng = [], np.random.RandomState(99)
    for idx in order:
        i, end = ds.slice_traj[idx], ds.slice_start[idx] + kw["window"]
        lo, hi = end + kw["min_future_sep"], gold["lens"][i] - kw["future_seq_len"]
        want.append(ref_rng.randint(lo, hi) if lo < hi else -1)
    rng = np.random.RandomState(99)
    assert ds.goal_starts(order, rng).tolist() == want
    assert rng.randint(0, 1 << 30) == ref_rng.randint(0, 1 << 30)
