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
    r"""Get dataloader for a specific phase (dataloader.py:10-36).

    Args:
        data_dir (str): path to the dataset directory
        phase (str): phase of the data, one of 'train', 'val', 'test'
        obs_len (int): length of observed trajectory
        pred_len (int): length of predicted trajectory
        batch_size (int): batch size
        device: optional; keep the dataset tensors on this device (no pinned-memory staging then)
    """
    assert phase in ['train', 'val', 'test']

    data_set = data_dir + '/' + phase + '/'
    shuffle = True if phase == 'train' else False
    drop_last = True if phase == 'train' else False

    dataset_phase = TrajectoryDataset(data_set, obs_len=obs_len, pred_len=pred_len, device=device)
    sampler_phase = None
    if batch_size > 1:
        sampler_phase = TrajBatchSampler(dataset_phase, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last)
    on_device = dataset_phase.obs_traj.is_cuda
    loader_phase = DataLoader(dataset_phase, collate_fn=traj_collate_fn, batch_sampler=sampler_phase,
                              pin_memory=not on_device)
    return loader_phase


def traj_collate_fn(data):
    r"""Collate function for the dataloader (dataloader.py:39-66).

    Returns obs (num_ped, obs_len, 2), pred (num_ped, pred_len, 2), non_linear_ped (num_ped,),
    loss_mask (num_ped, obs_len + pred_len), scene_mask (num_ped, num_ped) bool, seq_start_end (num_seq, 2).
    """
    obs_seq_list, pred_seq_list, non_linear_ped_list, loss_mask_list, _, _ = zip(*data)

    _len = [len(seq) for seq in obs_seq_list]
    cum_start_idx = [0] + np.cumsum(_len).tolist()
    seq_start_end = [[start, end] for start, end in zip(cum_start_idx, cum_start_idx[1:])]
    seq_start_end = torch.LongTensor(seq_start_end)
    device = obs_seq_list[0].device
    scene_mask = torch.zeros(sum(_len), sum(_len), dtype=torch.bool, device=device)
    for idx, (start, end) in enumerate(seq_start_end.tolist()):
        scene_mask[start:end, start:end] = 1

    out = [torch.cat(obs_seq_list, dim=0), torch.cat(pred_seq_list, dim=0),
           torch.cat(non_linear_ped_list, dim=0), torch.cat(loss_mask_list, dim=0), scene_mask, seq_start_end]
    return tuple(out)


class TrajBatchSampler(Sampler):
    r"""Samples batched elements by yielding a mini-batch of indices (dataloader.py:69-118).

    A batch is closed as soon as it holds at least ``batch_size`` pedestrians.
    """

    def __init__(self, data_source, batch_size=64, shuffle=False, drop_last=False, generator=None):
        self.data_source = data_source
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.generator = generator

    def __iter__(self):
        assert len(self.data_source) == len(self.data_source.num_peds_in_seq)

        if self.shuffle:
            if self.generator is None:
                generator = torch.Generator()
                generator.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            else:
                generator = self.generator
            indices = torch.randperm(len(self.data_source), generator=generator).tolist()
        else:
            indices = list(range(len(self.data_source)))
        num_peds_indices = self.data_source.num_peds_in_seq[indices]

        batch = []
        total_num_peds = 0
        for idx, num_peds in zip(indices, num_peds_indices):
            batch.append(idx)
            total_num_peds += num_peds
            if total_num_peds >= self.batch_size:
                yield batch
                batch = []
                total_num_peds = 0
        if len(batch) > 0 and not self.drop_last:
            yield batch

    def __len__(self):
        # Approximated number of batches (the order can be shuffled, so this number can vary from run to run).
        if self.drop_last:
            return sum(self.data_source.num_peds_in_seq) // self.batch_size
        else:
            return (sum(self.data_source.num_peds_in_seq) + self.batch_size - 1) // self.batch_size


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
    """1.0 if the last ``traj_len`` frames of ``traj`` (2, T) are non-linear, else 0.0 (dataloader.py:135-151).

    Host helper kept for API parity; the dataset builder evaluates the same criterion natively."""
    t = np.linspace(0, traj_len - 1, traj_len)
    res_x = np.polyfit(t, traj[0, -traj_len:], 2, full=True)[1]
    res_y = np.polyfit(t, traj[1, -traj_len:], 2, full=True)[1]
    if res_x + res_y >= threshold:
        return 1.0
    else:
        return 0.0


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
    """Dataloder for the Trajectory datasets (dataloader.py:154-241)."""

    def __init__(self, data_dir, obs_len=8, pred_len=12, skip=1, threshold=0.02, min_ped=1, delim='\t', device=None):
        """
        Args:
        - data_dir: Directory containing dataset files in the format <frame_id> <ped_id> <x> <y>
        - obs_len: Number of time-steps in input trajectories
        - pred_len: Number of time-steps in output trajectories
        - skip: Number of frames to skip while making the dataset
        - threshold: Minimum error to be considered for non-linear traj when using a linear predictor
        - min_ped: Minimum number of pedestrians that should be in a sequence
        - delim: Delimiter in the dataset files
        - device: optional device the tensors are kept on (default: host, as the reference)
        """
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
