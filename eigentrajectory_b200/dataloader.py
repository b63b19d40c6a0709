"""Dataset preprocessing and batching -- drop-in for the reference's ``utils/dataloader.py``.

Same names and semantics (``get_dataloader``, ``traj_collate_fn``, ``TrajBatchSampler``, ``read_file``,
``poly_fit``, ``TrajectoryDataset`` with ``obs_traj / pred_traj / loss_mask / non_linear_ped /
num_peds_in_seq / seq_start_end``), but the per-file work -- parsing ``<frame> <ped> <x> <y>`` text and cutting
it into ``obs_len + pred_len`` frame windows (dataloader.py:121-232) -- runs in ``libet_b200.so``
(``et_dataset_parse_host`` / ``et_dataset_windows_host``, C++ on the host: two linear passes instead of the
reference's Python loops over numpy masks, which take seconds per split).

Additions that do not change the reference behaviour:
  * ``TrajectoryDataset(..., device="cuda")`` keeps the tensors resident in HBM, so every ``__getitem__`` /
    collated batch is a device view or a device-side ``cat`` and no per-batch host->device copy remains;
  * ``save(path)`` / ``TrajectoryDataset.load(path)``: a flat binary cache of the preprocessed tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np
import torch
from torch.utils.data import Dataset
from torch.utils.data.dataloader import DataLoader
from torch.utils.data.sampler import Sampler

from ._lib import check, load

_MAGIC = b"ETDS0001"


def get_dataloader(data_dir, phase, obs_len, pred_len, batch_size, device=None):
    """Loader of one split (the reference's ``get_dataloader``, dataloader.py:10-36).

    ``data_dir/phase/`` is preprocessed natively; training shuffles the scenes and drops an incomplete last batch,
    validation / test keep file order.  ``batch_size`` counts pedestrians (see :class:`TrajBatchSampler`); with
    ``batch_size <= 1`` every batch is one scene.  ``device``: keep the split resident on that device."""
    assert phase in ['train', 'val', 'test']
    training = phase == 'train'
    dataset = TrajectoryDataset(f"{data_dir}/{phase}/", obs_len=obs_len, pred_len=pred_len, device=device)
    sampler = None
    if batch_size > 1:
        sampler = TrajBatchSampler(dataset, batch_size=batch_size, shuffle=training, drop_last=training)
    # device-resident tensors need no pinned staging
    return DataLoader(dataset, collate_fn=traj_collate_fn, batch_sampler=sampler, pin_memory=not dataset.obs_traj.is_cuda)


def traj_collate_fn(data):
    """Concatenate scenes into one batch (dataloader.py:39-66).

    Returns (obs (P, obs_len, 2), pred (P, pred_len, 2), non_linear_ped (P,), loss_mask (P, obs_len + pred_len),
    scene_mask (P, P) bool -- block diagonal, True inside a scene --, seq_start_end (n_scenes, 2) int64)."""
    obs, pred, non_linear, loss_mask = (list(column) for column in list(zip(*data))[:4])
    sizes = [scene.size(0) for scene in obs]
    ends = np.cumsum(sizes)
    seq_start_end = torch.as_tensor(np.stack([ends - np.asarray(sizes), ends], axis=1).reshape(-1, 2), dtype=torch.int64)
    blocks = [torch.ones((m, m), dtype=torch.bool, device=obs[0].device) for m in sizes]
    scene_mask = torch.block_diag(*blocks)
    return (torch.cat(obs, dim=0), torch.cat(pred, dim=0), torch.cat(non_linear, dim=0), torch.cat(loss_mask, dim=0),
            scene_mask, seq_start_end)


class TrajBatchSampler(Sampler):
    """Batches of scene indices holding at least ``batch_size`` pedestrians each (dataloader.py:69-118).

    Scenes are taken in order (or in a fresh random permutation when ``shuffle``); a batch is emitted as soon as its
    pedestrian count reaches ``batch_size``; the left-over scenes form a last, smaller batch unless ``drop_last``."""

    def __init__(self, data_source, batch_size=64, shuffle=False, drop_last=False, generator=None):
        self.data_source = data_source
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.generator = generator

    def _order(self):
        n_scenes = len(self.data_source)
        if not self.shuffle:
            return list(range(n_scenes))
        gen = self.generator
        if gen is None:                      # a fresh seed per epoch, drawn from the global torch generator
            gen = torch.Generator()
            gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
        return torch.randperm(n_scenes, generator=gen).tolist()

    def __iter__(self):
        peds = self.data_source.num_peds_in_seq
        assert len(self.data_source) == len(peds)
        pending, count = [], 0
        for scene in self._order():
            pending.append(scene)
            count += peds[scene]
            if count >= self.batch_size:
                yield pending
                pending, count = [], 0
        if pending and not self.drop_last:
            yield pending

    def __len__(self):
        # an estimate: the real number depends on how the (possibly shuffled) scenes fill the batches
        total = sum(self.data_source.num_peds_in_seq)
        return total // self.batch_size if self.drop_last else -(-total // self.batch_size)


def _delim_char(delim):
    if delim == 'tab':
        delim = '\t'
    elif delim == 'space':
        delim = ' '
    assert isinstance(delim, str) and len(delim) == 1, "delimiter must be a single character"
    return delim


def _parse(text: bytes, delim):
    """(n, 4) float64 rows of a ``<frame><delim><ped><delim><x><delim><y>`` text buffer (native parser)."""
    lib = load()
    cap = text.count(b"\n") + 1
    rows = np.empty((cap, 4), dtype=np.float64)
    n = C.c_int64(0)
    check(lib.et_dataset_parse_host(text, len(text), _delim_char(delim).encode()[0], rows.ctypes.data_as(C.c_void_p), cap,
                                    C.byref(n)), "et_dataset_parse_host")
    return rows[:n.value]


def read_file(_path, delim='\t'):
    """np.ndarray (n_rows, 4) float64 of one dataset file (dataloader.py:121-132)."""
    with open(_path, 'rb') as f:
        return _parse(f.read(), delim)


def poly_fit(traj, traj_len, threshold):
    """1.0 when the last ``traj_len`` frames of ``traj`` (2, T) are not fitted by a quadratic in time to within
    ``threshold`` (sum of the squared residuals of x and y), else 0.0 -- the reference's non-linearity flag
    (dataloader.py:135-151).  Host helper for API parity; the dataset builder applies the same rule natively."""
    t = np.arange(traj_len, dtype=np.float64)
    design = np.stack([t * t, t, np.ones_like(t)], axis=1)
    total = 0.0
    for axis in (0, 1):
        y = np.asarray(traj[axis, -traj_len:], dtype=np.float64)
        coef = np.linalg.lstsq(design, y, rcond=None)[0]
        total += float(((design @ coef - y) ** 2).sum())
    return 1.0 if total >= threshold else 0.0


def build_windows(rows, obs_len=8, pred_len=12, skip=1, threshold=0.02, min_ped=1):
    """The body of ``TrajectoryDataset.__init__`` for ONE file (dataloader.py:189-226).

    rows (n, 4) float64 -> (traj (N, obs_len + pred_len, 2) float32, non_linear (N,) float32, num_peds_in_seq (n_seq,) int)."""
    lib = load()
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n_rows = rows.shape[0]
    seq_len = obs_len + pred_len
    # every kept (pedestrian, window) owns the row of its first frame, and every window owns a frame: n_rows bounds both
    cap_peds, cap_seq = max(n_rows, 1), max(n_rows, 1) + 1
    traj = np.empty((cap_peds, seq_len, 2), dtype=np.float32)
    non_linear = np.empty((cap_peds,), dtype=np.float32)
    peds_in_seq = np.empty((cap_seq,), dtype=np.int32)
    n_peds, n_seq = C.c_int64(0), C.c_int64(0)
    vp = C.c_void_p
    check(lib.et_dataset_windows_host(rows.ctypes.data_as(vp), n_rows, int(obs_len), int(pred_len), int(skip), float(threshold),
                                      int(min_ped), traj.ctypes.data_as(vp), non_linear.ctypes.data_as(vp),
                                      peds_in_seq.ctypes.data_as(vp), cap_peds, cap_seq, C.byref(n_peds), C.byref(n_seq)),
          "et_dataset_windows_host")
    return traj[:n_peds.value], non_linear[:n_peds.value], peds_in_seq[:n_seq.value].astype(np.int64)


class TrajectoryDataset(Dataset):
    """All complete pedestrian windows of a split as (N, T, 2) tensors (the reference's class of the same name,
    dataloader.py:154-241), built by the native windowing code."""

    def __init__(self, data_dir, obs_len=8, pred_len=12, skip=1, threshold=0.02, min_ped=1, delim='\t', device=None):
        """Every file of ``data_dir`` holds rows ``<frame_id> <ped_id> <x> <y>``.  A window is ``obs_len + pred_len``
        consecutive frames, windows start every ``skip`` frames; a window is kept when MORE than ``min_ped``
        pedestrians are present in all of its frames; ``threshold`` is the residual above which a future counts as
        non-linear; ``delim`` is a character, 'tab' or 'space'; ``device`` (optional) is where the tensors are kept
        (default: host memory, like the reference)."""
        super(TrajectoryDataset, self).__init__()

        self.data_dir = data_dir
        self.obs_len = obs_len
        self.pred_len = pred_len
        self.skip = skip
        self.seq_len = self.obs_len + self.pred_len
        self.delim = delim

        all_files = os.listdir(self.data_dir)
        all_files = [os.path.join(self.data_dir, _path) for _path in all_files]
        trajs, non_linear, num_peds_in_seq = [], [], []
        for path in all_files:
            t, nl, nps = build_windows(read_file(path, delim), obs_len, pred_len, skip, threshold, min_ped)
            trajs.append(t)
            non_linear.append(nl)
            num_peds_in_seq.append(nps)
        if sum(len(x) for x in num_peds_in_seq) == 0:
            # the reference concatenates an empty list here (dataloader.py:227)
            raise ValueError("need at least one array to concatenate: no window of " + str(self.data_dir) +
                             f" holds more than min_ped = {min_ped} complete pedestrians")
        self._finish(np.concatenate(trajs, axis=0), np.concatenate(non_linear, axis=0),
                     np.concatenate(num_peds_in_seq, axis=0), device)

    def _finish(self, traj, non_linear, num_peds_in_seq, device):
        self.num_seq = len(num_peds_in_seq)
        self.num_peds_in_seq = np.array(num_peds_in_seq)
        traj = torch.from_numpy(np.ascontiguousarray(traj))
        # (N, T, 2) "NTC", contiguous (the reference holds permuted views of (N, 2, T); values are identical)
        self.obs_traj = traj[:, :self.obs_len].contiguous()
        self.pred_traj = traj[:, self.obs_len:].contiguous()
        # every kept pedestrian spans the whole window (pad_front = 0, pad_end = seq_len): the mask is all ones
        self.loss_mask = torch.ones((traj.size(0), self.seq_len), dtype=torch.float)
        self.non_linear_ped = torch.from_numpy(np.ascontiguousarray(non_linear)).type(torch.float)
        cum_start_idx = [0] + np.cumsum(self.num_peds_in_seq).tolist()
        self.seq_start_end = [(start, end) for start, end in zip(cum_start_idx, cum_start_idx[1:])]
        if device is not None:
            self.to(device)

    def to(self, device):
        """Move the tensors (in place) to ``device``; batches are then device tensors."""
        for name in ("obs_traj", "pred_traj", "loss_mask", "non_linear_ped"):
            setattr(self, name, getattr(self, name).to(device))
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    # ---- flat binary cache ----------------------------------------------------------------------
    def save(self, path):
        """Write the preprocessed tensors as one flat little-endian file: magic, (obs_len, pred_len, N, n_seq) int64,
        then traj (N, seq_len, 2) f32, non_linear (N) f32, num_peds_in_seq (n_seq) int64."""
        traj = torch.cat([self.obs_traj, self.pred_traj], dim=1).cpu().numpy().astype("<f4")
        with open(path, "wb") as f:
            f.write(_MAGIC)
            f.write(struct.pack("<4q", self.obs_len, self.pred_len, traj.shape[0], self.num_seq))
            f.write(traj.tobytes())
            f.write(self.non_linear_ped.cpu().numpy().astype("<f4").tobytes())
            f.write(np.asarray(self.num_peds_in_seq, dtype="<i8").tobytes())

    @classmethod
    def load(cls, path, device=None):
        """Rebuild a dataset from a file written by :meth:`save` (no text parsing, no windowing)."""
        with open(path, "rb") as f:
            buf = f.read()
        assert buf[:8] == _MAGIC, f"{path}: not an ETDS cache"
        obs_len, pred_len, n, n_seq = struct.unpack_from("<4q", buf, 8)
        seq_len = obs_len + pred_len
        off = 8 + 32
        traj = np.frombuffer(buf, dtype="<f4", count=n * seq_len * 2, offset=off).reshape(n, seq_len, 2)
        off += traj.nbytes
        non_linear = np.frombuffer(buf, dtype="<f4", count=n, offset=off)
        off += non_linear.nbytes
        nps = np.frombuffer(buf, dtype="<i8", count=n_seq, offset=off)
        self = cls.__new__(cls)
        Dataset.__init__(self)
        self.data_dir, self.obs_len, self.pred_len, self.skip, self.delim = path, obs_len, pred_len, None, None
        self.seq_len = seq_len
        self._finish(traj.copy(), non_linear.copy(), nps.copy(), device)
        return self

    def __len__(self):
        return self.num_seq

    def __getitem__(self, index):
        start, end = self.seq_start_end[index]
        out = [self.obs_traj[start:end], self.pred_traj[start:end],
               self.non_linear_ped[start:end], self.loss_mask[start:end], None, [[0, end - start]]]
        return out
