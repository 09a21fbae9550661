"""N>1 host logic on CPU: world_size-2 gloo run of the frame sharding and the NaN-aware plane reduction
(the only collective of the path)."""
import os
import subprocess
import sys
import numpy as np
from helpers import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from wass_b200 import launcher
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 7
planes_all = np.array([[0.1 * i, 0.2, 0.9, -3.0 - i] for i in range(n)])
planes_all[2] = np.nan            # a frame whose RANSAC failed (plane.txt = "nan nan nan nan")
owned = launcher.shard(n, rank, world)
assert owned == list(range(rank, n, world))
mean, allp = launcher.reduce_planes([planes_all[i] for i in owned], n, owned, dist)
ref = np.nanmean(planes_all, axis=0)          # what wassgridsurface computes from planes.txt
assert np.allclose(mean, ref, rtol=1e-14), (mean, ref)
assert np.array_equal(np.isnan(allp), np.isnan(planes_all)) and np.allclose(np.nan_to_num(allp), np.nan_to_num(planes_all))
os.write(1, ("rank %%d ok\n" %% rank).encode())      # one write: atomic on the shared pipe
dist.destroy_process_group()
'''


def test_two_rank_gloo_plane_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", GLOO_SOCKET_IFNAME="lo")
    import socket
    for attempt in range(3):               # the rendezvous itself can lose a race for the port on a busy host: retry
        with socket.socket() as sk:        # a free rendezvous port (a fixed one can still be in TIME_WAIT)
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                           capture_output=True, text=True, env=env, timeout=300)
        if r.returncode == 0 or "AssertionError" in r.stderr:
            break
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_shard_is_a_partition():
    from wass_b200 import launcher
    for n in (0, 1, 5, 32, 1024):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in launcher.shard(n, r, w))
            assert got == list(range(n))
            sizes = [len(launcher.shard(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_single_rank_reduce_matches_nanmean():
    from wass_b200 import launcher
    planes = np.array([[0, 0.6, 0.8, -2.0], [np.nan] * 4, [0.1, 0.5, 0.86, -2.5]])
    mean, allp = launcher.reduce_planes(list(planes), 3, [0, 1, 2], None)
    assert np.allclose(mean, np.nanmean(planes, axis=0))
