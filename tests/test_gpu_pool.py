"""The field-buffer pool behind gx_malloc / gx_free (genetic_b200/csrc/mempool.cu): exact-size reuse, no reuse of a block
whose last user is still running, release through gx_workspace_release, and results unchanged by pooling."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _malloc(lib, nbytes):
    p = C.c_void_p()
    lib.gx_malloc(C.byref(p), nbytes)
    return p


def test_freed_blocks_come_back_by_exact_size():
    from genetic_b200 import lib as L
    lib = L.load()
    lib.gx_set_device(0)
    lib.gx_workspace_release()
    a = _malloc(lib, 3 << 20)
    addr = a.value
    lib.gx_free(a)
    lib.gx_synchronize(None)
    b = _malloc(lib, 3 << 20)
    assert b.value == addr                       # same size: the pooled block
    c = _malloc(lib, (3 << 20) + 256)
    assert c.value != addr                       # another size: a new mapping
    lib.gx_free(b)
    lib.gx_free(c)
    lib.gx_workspace_release()                   # drops the idle blocks; the library still works afterwards
    d = _malloc(lib, 3 << 20)
    lib.gx_memset_zero(d, 3 << 20, None)
    lib.gx_synchronize(None)
    lib.gx_free(d)


def test_a_block_in_use_is_not_handed_out():
    """Free a field while transforms on it are still queued: the next allocation of that size must not get the same block
    until the queued work has finished (the old cudaFree synchronised the device instead)."""
    from genetic_b200 import lib as L
    lib = L.load()
    lib.gx_set_device(0)
    lib.gx_workspace_release()
    n = 256
    nbytes = lib.gx_field_bytes(n, L.F64)
    a = _malloc(lib, nbytes)
    lib.gx_memset_zero(a, nbytes, None)
    for _ in range(120):                         # ~40 ms of queued device work on the block
        lib.gx_fft_forward(a, n, L.F64, None)
        lib.gx_fft_inverse(a, n, L.F64, None)
    addr = a.value
    import torch
    done = torch.cuda.Event()
    done.record()                                # torch's default stream is the legacy stream the transforms were queued on
    lib.gx_free(a)
    b = _malloc(lib, nbytes)                     # immediately: the transforms are still running
    if not done.query():                         # (queried after the allocation: still running now => running then)
        assert b.value != addr
    lib.gx_synchronize(None)
    c = _malloc(lib, nbytes)                     # now the first block is idle
    assert c.value == addr or b.value == addr
    lib.gx_free(b)
    lib.gx_free(c)


def test_pooling_does_not_change_results(golden_dir, tmp_path):
    """The two-level overdensity + variance golden (test_02i) with GX_POOL=0 and with the pool: identical dumped grids."""
    import shutil
    out = {}
    for pool in ("0", "1"):
        d = tmp_path / pool
        shutil.copytree(os.path.join(golden_dir, "test_02i_multilevel_algorithm"), d / "t")
        shutil.copy(os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat"), d)
        env = dict(os.environ, GX_POOL=pool, PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-m", "genetic_b200", str(d / "t" / "paramfile.txt")], env=env, cwd=str(d / "t"),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        out[pool] = [np.load(str(d / "t" / ("grid-%d.npy" % l))) for l in (0, 1)]
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b)
