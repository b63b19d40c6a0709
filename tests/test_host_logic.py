"""CPU: host-side logic of the package that needs no GPU and no library call."""
import torch

from eigentrajectory_b200 import ops
from eigentrajectory_b200.normalizer import TrajNorm
from eigentrajectory_b200.parallel import shard_bounds


def test_host_chunks_partition_the_batch():
    for n in (1, 100, 32768, 32769, 524288, 600_000, 1_000_000, 1_000_001, 5_000_000):
        cuts = ops._host_chunks(n)
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])) and all(b > a for a, b in cuts)
        sizes = [b - a for a, b in cuts]
        assert max(sizes) <= ops.HOST_CHUNK + ops.HOST_CHUNK // 4 + 128
        if n >= 4 * ops.HOST_CHUNK:           # short first chunk, tapered tail, tile-aligned chunk starts in between
            assert sizes[0] == ops.HOST_FIRST_CHUNK and sizes[-1] < ops.HOST_CHUNK
            assert all(a % 128 == 0 for a, _ in cuts[:-2])


def test_trajnorm_deferred_state_is_gathered_on_first_read():
    n = 10
    full = (torch.arange(n * 2, dtype=torch.float32).reshape(n, 1, 2), torch.randn(n, 2, 2), torch.ones(n, 1, 1))
    rows = torch.tensor([True, False] * 5)
    tn = TrajNorm(ori=True, rot=True, sca=False)
    tn.set_deferred(full, rows)
    assert tn._deferred is not None                      # nothing gathered yet
    assert torch.equal(tn.traj_ori, full[0][rows]) and tn._deferred is None
    assert torch.equal(tn.traj_rot, full[1][rows]) and tn.traj_sca is None      # the scale stage is off: state untouched
    # an explicit assignment after a deferred hand-over wins and is not overwritten later
    tn.set_deferred(full, ~rows)
    tn.traj_ori = torch.zeros(3, 1, 2)
    assert tn.traj_ori.shape == (3, 1, 2) and torch.equal(tn.traj_rot, full[1][~rows])
    flags_and_state = tn.get_params()
    other = TrajNorm()
    other.set_params(*flags_and_state)
    assert other.state()[2] is None and torch.equal(other.traj_rot, tn.traj_rot)


def test_shard_bounds_cover_all_rows():
    for n, world in ((10, 3), (1_000_003, 8), (5, 8), (0, 4)):
        cuts = [shard_bounds(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
