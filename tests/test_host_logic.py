"""Host-side logic on CPU: shard arithmetic, the world_size-2 gather over gloo, synthetic data determinism,
PVConv state_dict compatibility."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from ri_b200.shard import shard_range
    for total in (0, 1, 7, 32, 256, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ri_b200 import shard
    r, w, _ = shard.init_from_env(backend="gloo")
    lo, hi = shard.shard_range(total, r, w)
    full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 2, 3)
    got = shard.gather_clouds(full[lo:hi].clone(), total)
    got_t = shard.gather_clouds(full.transpose(0, 1)[:, lo:hi].contiguous(), total, dim=1)
    slow = shard.max_over_ranks(float(rank + 1), "cpu")
    q.put((rank, bool(torch.equal(got, full)), bool(torch.equal(got_t, full.transpose(0, 1))), slow))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("total", [7, 32])
def test_gather_world_size_2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok1 and ok2 for _, ok1, ok2, _ in res)
    assert all(abs(s - 2.0) < 1e-9 for *_, s in res)


def test_synth_is_seeded_and_well_formed():
    from ri_b200 import synth
    a = synth.make_clouds(6, 256, seed=3); b = synth.make_clouds(6, 256, seed=3); c = synth.make_clouds(6, 256, seed=4)
    assert a.dtype == np.float32 and a.shape == (6, 6, 256)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.abs(a[:, :3].mean(2)).max() < 1e-5                                  # centred
    assert np.linalg.norm(a[:, :3], axis=1).max() <= 1.0 + 1e-5                   # inside the unit ball
    assert np.allclose(np.linalg.norm(a[:, 3:], axis=1), 1.0, atol=1e-5)          # unit normals
    src, tgt, R, t = synth.make_pairs(4, 128, seed=1)
    assert np.allclose(np.einsum('pij,pkj->pik', R, R), np.eye(3)[None], atol=1e-5)
    assert np.abs(tgt[:, :3] - (np.einsum('pij,pjn->pin', R, src[:, :3]) + t[:, :, None])).max() <= 0.05 + 1e-6
    s = synth.make_scan(5000, seed=2)
    assert s.shape == (6, 5000) and np.isfinite(s).all()


def test_pvconv_state_dict_keys_match_reference_layout():
    """Keys a reference checkpoint holds for one PVConv (pvconv.py:27-43): voxel_layers.{0,1,3,4}.*, SE3d fc,
    point_layers.layers.{0,1}.*, coefficient."""
    import ri_b200
    conv = ri_b200.modules.PVConv(8, 16, 'dgcnn_kernel', 'spherical', 3, 8, with_coeff=True, with_se=True)
    keys = set(conv.state_dict().keys())
    want = {"coefficient", "voxel_layers.0.weight", "voxel_layers.0.bias", "voxel_layers.1.weight",
            "voxel_layers.1.running_mean", "voxel_layers.3.weight", "voxel_layers.4.running_var",
            "voxel_layers.6.fc.0.weight", "voxel_layers.6.fc.2.weight", "point_layers.layers.0.weight",
            "point_layers.layers.1.running_mean"}
    assert want <= keys
    assert conv.point_layers.layers[0].in_channels == 16      # dgcnn kernel doubles the point-branch input
