"""CPU: the native dataset preprocessing (utils/dataloader.py:121-232 in C++ behind the C ABI) against outputs of the
unmodified reference frozen in tests/golden/dataset.npz (make_golden.py::dataset).  Host-side code: no GPU needed."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset.npz")


@pytest.fixture(scope="module")
def dl():
    import __graft_entry__ as g
    g.build()
    from eigentrajectory_b200 import dataloader
    return dataloader


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def write_split(tmp_path, gold, tag, names):
    # the reference lists files with os.listdir; recreate them in the recorded order so that a different directory
    # order on this machine cannot permute the concatenation
    d = tmp_path / tag
    d.mkdir()
    for name in names:
        (d / name).write_bytes(gold[f"{tag}_text_{name}"].tobytes())
    return str(d) + "/"


def check_dataset(ds, gold, tag):
    assert torch.equal(ds.obs_traj, torch.from_numpy(gold[f"{tag}_obs"]))            # bit-exact float32
    assert torch.equal(ds.pred_traj, torch.from_numpy(gold[f"{tag}_pred"]))
    assert torch.equal(ds.loss_mask, torch.from_numpy(gold[f"{tag}_loss_mask"]))
    assert torch.equal(ds.non_linear_ped, torch.from_numpy(gold[f"{tag}_non_linear"]))
    assert np.array_equal(ds.num_peds_in_seq, gold[f"{tag}_num_peds_in_seq"])
    assert np.array_equal(np.asarray(ds.seq_start_end), gold[f"{tag}_seq_start_end"])
    assert len(ds) == len(gold[f"{tag}_num_peds_in_seq"])
    assert ds.obs_traj.is_contiguous() and ds.obs_traj.dtype == torch.float32


def ordered_listdir(monkeypatch, order):
    real = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p: list(order) if set(real(p)) == set(order) else real(p))


def test_eth_test_split_matches_reference(dl, gold, tmp_path):
    path = write_split(tmp_path, gold, "eth", ["biwi_eth.txt"])
    ds = dl.TrajectoryDataset(path, obs_len=8, pred_len=12)
    check_dataset(ds, gold, "eth")
    # batching of the test phase and one collated batch
    batches = list(dl.TrajBatchSampler(ds, batch_size=32, shuffle=False, drop_last=False))
    assert np.array_equal(np.array([len(b) for b in batches]), gold["eth_batch_sizes"])
    assert np.array_equal(np.array(batches[0]), gold["eth_batch_first"])
    col = dl.traj_collate_fn([ds[i] for i in batches[3]])
    assert torch.equal(col[0], torch.from_numpy(gold["eth_b3_obs"])) and torch.equal(col[1], torch.from_numpy(gold["eth_b3_pred"]))
    assert torch.equal(col[4], torch.from_numpy(gold["eth_b3_scene_mask"]))
    assert torch.equal(col[5], torch.from_numpy(gold["eth_b3_seq_start_end"]))


@pytest.mark.parametrize("tag,kw", [("syn", dict(obs_len=8, pred_len=12)),
                                    ("syn_skip3", dict(obs_len=8, pred_len=12, skip=3)),
                                    ("syn_short", dict(obs_len=3, pred_len=5, min_ped=0, threshold=0.002))])
def test_synthetic_splits_match_reference(dl, gold, tmp_path, monkeypatch, tag, kw):
    order = [str(x) for x in gold[f"{tag}_file_order"]]
    path = write_split(tmp_path, gold, tag, order)
    ordered_listdir(monkeypatch, order)
    ds = dl.TrajectoryDataset(path, **kw)
    check_dataset(ds, gold, tag)
    assert 0 < float(ds.non_linear_ped.sum()) < len(ds.non_linear_ped)       # both classes occur in the fixture


def test_read_file_and_get_dataloader(dl, gold, tmp_path):
    d = tmp_path / "eth" / "test"
    d.mkdir(parents=True)
    (d / "biwi_eth.txt").write_bytes(gold["eth_text_biwi_eth.txt"].tobytes())
    rows = dl.read_file(str(d / "biwi_eth.txt"))
    text = gold["eth_text_biwi_eth.txt"].tobytes().decode()
    want = np.asarray([[float(v) for v in line.strip().split("\t")] for line in text.splitlines()])
    assert rows.dtype == np.float64 and np.array_equal(rows, want)
    loader = dl.get_dataloader(str(tmp_path / "eth"), "test", 8, 12, 1)
    first = next(iter(loader))
    assert first[0].shape[1:] == (8, 2) and first[1].shape[1:] == (12, 2)
    assert first[4].dtype == torch.bool and first[4].shape == (first[0].size(0),) * 2
    n_batches = sum(1 for _ in loader)
    assert n_batches == len(gold["eth_num_peds_in_seq"])


def test_native_nonlinear_flag_agrees_with_numpy_polyfit(dl, gold):
    """The builder's residual (orthonormal-basis projection in long double) and the mirror of the reference's
    poly_fit (np.polyfit, dataloader.py:135-151) take the same decision on every pedestrian of the fixtures."""
    for tag, pred_len, thr in (("syn", 12, 0.02), ("eth", 12, 0.02), ("syn_short", 5, 0.002)):
        full = np.concatenate([gold[f"{tag}_obs"], gold[f"{tag}_pred"]], axis=1).astype(np.float64)   # (N, T, 2)
        want = gold[f"{tag}_non_linear"]
        got = np.array([dl.poly_fit(tr.T, pred_len, thr) for tr in full[:400]])
        assert np.array_equal(got, want[:400]), tag


def test_space_delimiter_and_skip(dl, tmp_path):
    d = tmp_path / "sp"
    d.mkdir()
    rows = "".join(f"{10 * f} {p}.0 {0.1 * f + p:.3f} {0.2 * f:.3f}\n" for f in range(30) for p in (1, 2, 3))
    (d / "f.txt").write_text(rows)
    ds = dl.TrajectoryDataset(str(d) + "/", obs_len=8, pred_len=12, skip=2, delim="space")
    # 30 frames -> num_sequences = ceil(11 / 2) = 6 windows starting at frames 0, 2, .., 10 (+ the empty one at 12)
    assert len(ds) == 6 and ds.obs_traj.shape == (18, 8, 2) and ds.pred_traj.shape == (18, 12, 2)
    assert torch.allclose(ds.obs_traj[3, :, 0], torch.tensor([0.1 * f + 1 for f in range(2, 10)], dtype=torch.float32))
    assert float(ds.non_linear_ped.sum()) == 0.0 and torch.equal(ds.loss_mask, torch.ones(18, 20))


def test_oracle_restatement_matches_reference_and_native_fuzz(dl, gold):
    """The numpy restatement in oracle/ is pinned to the reference's outputs (fixtures), then serves as the checker for
    the native builder on random scenes: ragged lifetimes, missing frames, several skips and window lengths."""
    from oracle import et_oracle as O
    for tag, kw in (("eth", dict(obs_len=8, pred_len=12)), ("syn_short", dict(obs_len=3, pred_len=5, min_ped=0, threshold=0.002))):
        parts = []
        for name in [str(x) for x in gold[f"{tag}_file_order"]]:
            text = gold[f"{tag}_text_{name}"].tobytes().decode()
            parts.append(O.dataset_windows(O.dataset_parse(text), **kw))
        traj = np.concatenate([p[0] for p in parts])
        assert np.array_equal(traj[:, :kw["obs_len"]], gold[f"{tag}_obs"]) and np.array_equal(traj[:, kw["obs_len"]:], gold[f"{tag}_pred"])
        assert np.array_equal(np.concatenate([p[1] for p in parts]), gold[f"{tag}_non_linear"])
        assert np.array_equal(np.concatenate([p[2] for p in parts]), gold[f"{tag}_num_peds_in_seq"])
    rng = np.random.RandomState(5)
    for case in range(12):
        n_frames, n_peds = int(rng.randint(6, 60)), int(rng.randint(1, 25))
        frames = np.sort(rng.choice(np.arange(0, 3 * n_frames), size=n_frames, replace=False)) * 10
        rows = []
        for ped in range(n_peds):
            a = int(rng.randint(0, n_frames))
            b = int(min(n_frames, a + rng.randint(1, 40)))
            p0, v = rng.uniform(-5, 5, 2), rng.uniform(-0.5, 0.5, 2)
            for j, f in enumerate(frames[a:b]):
                pos = p0 + v * j + (0.1 * np.sin(j) if ped % 3 == 0 else 0.0) + rng.normal(0, 1e-3, 2)
                rows.append((float(f), float(ped + 1), pos[0], pos[1]))
        rng.shuffle(rows)
        rows = np.asarray(sorted(rows, key=lambda r: r[0]), dtype=np.float64)
        kw = dict(obs_len=int(rng.randint(1, 9)), pred_len=int(rng.randint(4, 13)), skip=int(rng.randint(1, 4)),
                  threshold=0.02, min_ped=int(rng.randint(0, 3)))
        if len(np.unique(rows[:, 0])) < 2:
            continue
        want = O.dataset_windows(rows, **kw)
        got = dl.build_windows(rows, **kw)
        for g, w in zip(got, want):
            assert np.array_equal(np.asarray(g), w), (case, kw)


def test_cache_round_trip(dl, gold, tmp_path):
    path = write_split(tmp_path, gold, "eth", ["biwi_eth.txt"])
    ds = dl.TrajectoryDataset(path, obs_len=8, pred_len=12)
    ds.save(str(tmp_path / "eth.etds"))
    ds2 = dl.TrajectoryDataset.load(str(tmp_path / "eth.etds"))
    check_dataset(ds2, gold, "eth")
    assert ds2[5][0].shape == ds[5][0].shape and torch.equal(ds2[5][1], ds[5][1])


def test_malformed_inputs_fail_loudly(dl, tmp_path):
    from eigentrajectory_b200 import ETLibraryError
    good = "".join(f"{10 * f}\t{p}.0\t{0.1 * f + p:.4f}\t{0.2 * f:.4f}\n" for f in range(25) for p in (1, 2, 3))

    def build(text, **kw):
        d = tmp_path / f"case{build.n}"
        build.n += 1
        d.mkdir()
        (d / "f.txt").write_text(text)
        return dl.TrajectoryDataset(str(d) + "/", **kw)
    build.n = 0
    ds = build(good)
    assert len(ds) == 6 and ds.obs_traj.shape == (18, 8, 2)         # 25 frames -> 6 windows of 3 pedestrians
    with pytest.raises(ETLibraryError):
        build(good + "\n")                                          # blank line: float('') in the reference
    with pytest.raises(ETLibraryError):
        build(good.replace("\t1.0\t", "\tabc\t", 1))                # not a number
    with pytest.raises(ETLibraryError):
        build("\n".join(ln for ln in good.splitlines() if not ln.startswith("100\t2.0")) + "\n")   # a gap inside a window
    with pytest.raises(ETLibraryError):
        build("")                                                   # empty file
    # a window needs MORE than min_ped pedestrians (dataloader.py:221)
    one = "".join(f"{10 * f}\t1.0\t{0.1 * f:.4f}\t0.0\n" for f in range(25))
    assert len(build(one + "0\t2.0\t5.0\t5.0\n", min_ped=0)) == 6
    with pytest.raises(ValueError):
        build(one)                                                  # nothing kept: np.concatenate of nothing, as the reference
