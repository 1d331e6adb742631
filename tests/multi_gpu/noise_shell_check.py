"""Multi-GPU parity check of the staged, sharded noise draw at the bench sizes (run under torchrun on the GPU box; not
collected by pytest). Every rank draws its share of the shells of a parallel Fourier draw and stores the modes into their
owners' slabs over NVLink (gx_slab_noise_staged); rank r then compares, bit for bit, the planes kz = ks of its own ky-slab
for the outermost shells with values formed from the oracle's sequential gaussian stream of those shells (one stream per
shell, seed + ks: randomfieldgenerator.hpp:226-243), using the reference's loop order
    call = ((k1 + ks) * 2ks + (k2 + ks)) * 3 + 2  ->  mode (k1, k2, ks) = (g[2 call] + i g[2 call + 1]) / sqrt(2).

  torchrun --nproc-per-node 8 tests/multi_gpu/noise_shell_check.py --n 4096 --dtype f32 --shells 2
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--shells", type=int, default=2)
    ap.add_argument("--seed", type=int, default=8896131)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from genetic_b200 import lib as L, slab as SL
    from oracle import crng
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = L.load()
    lib.gx_set_device(local)
    n = a.size
    dtype = np.float32 if a.dtype == "f32" else np.float64
    cdt = np.complex64 if a.dtype == "f32" else np.complex128
    geo = SL.SlabGeometry(n, world, rank, np.dtype(dtype).itemsize)
    ops = SL.CudaSlabOps(geo, dtype)
    comm = SL.TorchComm()
    slab = ops.alloc_shared(geo.slab_bytes)
    scratch = ops.alloc(geo.slab_bytes)
    slab.fill_(0x7f)
    peers = comm.peer_pointers(slab)
    comm.stream_barrier()
    ops.noise_staged(slab, peers, a.seed, 0, None, 1.0, scratch)
    comm.stream_barrier()
    torch.cuda.synchronize()
    W = slab.view(torch.float32 if a.dtype == "f32" else torch.float64).view(n, geo.P, geo.nc, 2)
    sigma = dtype(1.0 / np.sqrt(2.0))
    bad = 0
    checked = 0
    for ks in range(n // 2 - a.shells, n // 2):
        g = crng.Stream(a.seed + ks).gaussians(24 * ks * ks)
        plane = W[:, :, ks, :].contiguous().cpu().numpy().view(cdt).reshape(n, geo.P)     # [kx index][ky_local]
        # signed k1 (= kx) and k2 (= ky) strictly inside (-ks, ks): never overwritten by the conjugate partner
        k1 = np.arange(-ks + 1, ks)
        iy = np.arange(geo.iy0, geo.iy0 + geo.P)
        k2 = np.where(iy <= n // 2, iy, iy - n)
        k2 = k2[(k2 > -ks) & (k2 < ks) & (iy != n // 2)]
        call = ((k1[:, None] + ks) * (2 * ks) + (k2[None, :] + ks)) * 3 + 2
        re = (g[2 * call] * sigma).astype(dtype) if dtype == np.float64 else (np.float64(sigma) * g[2 * call]).astype(dtype)
        im = (g[2 * call + 1] * sigma).astype(dtype) if dtype == np.float64 else (np.float64(sigma) * g[2 * call + 1]).astype(dtype)
        ix = np.where(k1 < 0, k1 + n, k1)
        jy = np.where(k2 < 0, k2 + n, k2) - geo.iy0
        got = plane[ix[:, None], jy[None, :]]
        bad += int(np.count_nonzero((got.real != re) | (got.imag != im)))
        checked += got.size
    t = torch.tensor([bad, checked], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    if rank == 0:
        print("noise_shell_check n=%d %s world=%d: %d modes of the outermost %d shells compared with the oracle stream, %d differ"
              % (n, a.dtype, world, int(t[1]), a.shells, int(t[0])))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 1 if int(t[0]) else 0


if __name__ == "__main__":
    sys.exit(main())
