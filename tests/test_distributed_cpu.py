"""
CPU test of the N > 1 host path (world_size 2 and 3, gloo): scanline partition + the single gather + reassembly.
The per-rank "renderer" here is the oracle restricted to the rank's rows (the CUDA kernels need a GPU; their
partition is tested on one device in tests/test_gpu_parity.py::test_scanline_partition_reassembles_bit_exact).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracing_b200.distributed import RadianceGather, SharedHostImage, local_rows, rows_max


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, w, h, mb, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.orcbind import Oracle
    from raytracing_b200 import scene_io
    from raytracing_b200.camera import default_camera
    sc = scene_io.load_scene("CornellBox")
    cam = default_camera(w, h)
    full, _, _ = Oracle(sc).render(cam, w, h, mb, row_first=rank, row_step=world, want_hits=False)
    slab = torch.from_numpy(np.ascontiguousarray(full[rank::world]).reshape(-1, 4))      # what a rank's device buffer holds
    g = RadianceGather(w, h, rank, world, torch.device("cpu"))
    assert slab.shape[0] == g.n_local
    slabs = g.gather(slab)
    if rank == 0:
        np.save(out_path, g.reassemble(slabs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_gather_reassemble(tmp_path, world):
    w, h, mb = 64, 37, 3          # 37 rows: ranks own unequal row counts
    assert sum(local_rows(h, r, world) for r in range(world)) == h and rows_max(h, world) == local_rows(h, 0, world)
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, mb, out), nprocs=world, join=True)
    from oracle.orcbind import Oracle
    from raytracing_b200 import scene_io
    from raytracing_b200.camera import default_camera
    ref, _, _ = Oracle(scene_io.load_scene("CornellBox")).render(default_camera(w, h), w, h, mb, want_hits=False)
    got = np.load(out)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def _shared_worker(rank, world, port, w, h, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = SharedHostImage(w, h, rank, world, register_cuda=False)
    # what rt_resolve does with a full-image destination: this rank's rows only (row y belongs to rank y % world)
    rows = np.arange(rank, h, world)
    img.image[rows] = (rows[:, None, None] * 1000 + np.arange(w)[None, :, None] + np.arange(4)[None, None, :] * 0.25).astype(np.float32)
    img.complete()
    if rank == 0:
        np.save(out_path, np.array(img.image))
    img.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_host_image_collects_every_ranks_rows(tmp_path, world):
    """The parallel read-back path: every rank writes its scanlines into one shared host image; after one barrier rank 0 sees the whole frame."""
    w, h = 48, 29
    out = str(tmp_path / "shared.npy")
    mp.spawn(_shared_worker, args=(world, _free_port(), w, h, out), nprocs=world, join=True)
    y = np.arange(h)
    want = (y[:, None, None] * 1000 + np.arange(w)[None, :, None] + np.arange(4)[None, None, :] * 0.25).astype(np.float32)
    assert np.array_equal(np.load(out), want)
