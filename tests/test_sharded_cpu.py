"""World-size-2 gloo tests (CPU) of the N>1 host logic: band assignment, the gradient all-reduce closure the sharded
rasterizer installs, and image gathering.  The CUDA kernels themselves are covered on the GPU box
(tests/test_parity_gpu.py::test_sharded_backward_sums_to_whole, bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, layout, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import street_gaussians_b200 as sgb
        from street_gaussians_b200.sharded import ShardedGaussianRasterizer, band_of_rows
        H, W = 200, 64  # 13 tile rows (last one partial)
        st = sgb.GaussianRasterizationSettings(H, W, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
        r = ShardedGaussianRasterizer(st, layout=layout)
        assert r.world == world and r.rank == rank and r.band is not None and r.grad_reduce is not None
        # every tile row is owned by exactly one rank
        rows = (H + 15) // 16
        mine = torch.tensor([1 if (r.band.begin <= t < r.band.end and (t - r.band.begin) % r.band.step == 0) else 0 for t in range(rows)])
        tot = mine.clone()
        dist.all_reduce(tot)
        assert (tot == 1).all(), tot
        assert int(band_of_rows(H, rank, world, layout).sum()) == int(min(H, 1 << 30) and sum(min(16, H - 16 * t) for t in range(rows) if mine[t]))
        # the reduce closure sums the partial per-Gaussian sums (and tolerates S == 0)
        g2d = torch.full((5, 12), float(rank + 1))
        gsem = torch.zeros(5, 0)
        out, outsem = r.grad_reduce(g2d, gsem)
        assert torch.allclose(out, torch.full((5, 12), float(sum(range(1, world + 1))))) and outsem.shape == (5, 0)
        gsem = torch.full((5, 3), float(rank))
        _, outsem = r.grad_reduce(torch.zeros(5, 12), gsem)
        assert torch.allclose(outsem, torch.full((5, 3), float(sum(range(world)))))
        # gather_images: zero-padded band images add up to the full frame
        img = torch.zeros(1, H, W)
        img[:, band_of_rows(H, rank, world, layout)] = 1.0
        (full,) = r.gather_images(img)
        assert torch.allclose(full, torch.ones(1, H, W))
        # Gaussian-sharded module: same cyclic band, collective agreement on the per-rank slot count
        from street_gaussians_b200.sharded import GaussianShardedRasterizer, cyclic_band
        if layout == "cyclic":
            g = GaussianShardedRasterizer(st, exchange="p2p")
            assert g.band == cyclic_band(H, rank, world) and g.world == world and g.rank == rank and g.chunk is None
            assert g.repartition(10 if rank == 0 else 17) == 17 and g.chunk_for(3) == 17
        else:
            g = GaussianShardedRasterizer(st, layout=layout, chunk=9)
            assert g.chunk_for(5) == 9 and g.exchange == "nccl"
            try:
                GaussianShardedRasterizer(st, layout=layout, exchange="p2p")
                raise AssertionError("p2p with a contiguous layout must be rejected")
            except ValueError:
                pass
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout", ["cyclic", "contiguous"])
def test_sharded_host_logic_gloo_world2(layout):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, layout, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_is_plain_rasterizer():
    import street_gaussians_b200 as sgb
    from street_gaussians_b200.sharded import ShardedGaussianRasterizer
    st = sgb.GaussianRasterizationSettings(32, 32, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    r = ShardedGaussianRasterizer(st)
    assert r.world == 1 and r.band is None and r.grad_reduce is None
