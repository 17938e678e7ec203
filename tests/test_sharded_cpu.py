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


# ---------------------------------------------------------------------------------------------------------------------
# Gaussian-sharded data flow under gloo: the six C-ABI steps are replaced by a linear toy renderer on CPU tensors, so the
# test checks what the HOST does at N > 1 — padding to the common chunk, rank-major global indexing, the all-gather /
# reduce-scatter shapes, band ownership, slicing of radii and gradients back to the local Gaussians — against a closed
# form.  The CUDA steps themselves are checked on the GPU (tests/test_parity_gpu.py, tools/check_gaussian_sharded.py).
# ---------------------------------------------------------------------------------------------------------------------
def _toy_weights(P_total, H):
    i = torch.arange(P_total).view(-1, 1)
    y = torch.arange(H).view(1, -1)
    return (((i * 7 + y * 3) % 5) + 1).float() / 5.0  # [P_total, H]


def _install_toy_steps(SH, H, W, rank, world):
    from street_gaussians_b200.rasterizer import _ForwardState

    def local_tensors(means3D, sh, colors, semantics, opac, scales, rots, cov):
        S = int(semantics.shape[1]) if (semantics is not None and semantics.dim() == 2) else 0
        return dict(means3D=means3D.float(), opacities=opac.float(), sh=None, colors_precomp=None, scales=None, rotations=None,
                    cov3Ds_precomp=None, semantics=semantics.float() if S > 0 else None)

    def project(tensors, settings, chunk):
        P = tensors["means3D"].shape[0]
        assert P <= chunk
        rec = torch.full((chunk, 12), float("nan"))  # padding slots carry garbage: only radii == 0 may protect them
        rec[:P] = 0
        rec[:P, :3] = tensors["means3D"]
        rec[:P, 3] = tensors["opacities"][:, 0]
        radii = torch.zeros(chunk, dtype=torch.int32)
        radii[:P] = 1
        return rec, radii

    def alloc(settings, P_total, S, device):
        st = _ForwardState()
        st.geom, st.img, st.binning, st.num_instances = torch.empty(P_total, 12), None, None, 0
        return st, st.geom, 0, 0

    def own_rows(band):
        return SH.band_of_rows(H, rank, world).float() if band is not None else torch.ones(H)

    def forward(settings, band, st, sizes, radii_all, sem_all, capacity):
        rec, vis = torch.nan_to_num(st.geom, nan=1e9), (radii_all > 0).float().view(-1, 1)
        Wt = _toy_weights(rec.shape[0], H) * vis * own_rows(band).view(1, -1)
        img = lambda v: (v.t() @ Wt).view(-1, H, 1).expand(-1, H, W).contiguous()
        S = sem_all.shape[1] if sem_all is not None else 0
        sem = img(sem_all) if S else torch.zeros(0, H, W)
        return img(rec[:, :3] * vis), img(rec[:, 3:4] * vis), img(vis), sem

    def backward_blend(settings, band, st, P_total, sem_all, alpha, gc, gd, ga, gs, grad2d_out=None):
        Wt = _toy_weights(P_total, H) * own_rows(band).view(1, -1)
        g2 = torch.zeros(P_total, 12)
        g2[:, :3] = Wt @ gc.sum(dim=2).t()
        g2[:, 3] = Wt @ gd.sum(dim=2)[0]
        S = sem_all.shape[1] if sem_all is not None else 0
        return g2, (Wt @ gs.sum(dim=2).t() if S else torch.zeros(P_total, 0))

    def geom_local(settings, tensors, rec_local, radii_local, g2_local):
        P = tensors["means3D"].shape[0]
        return g2_local[:P, :3].clone(), torch.zeros(P, 3), None, None, g2_local[:P, 3:4].clone(), None, None, None

    SH._local_tensors, SH.project_records, SH.alloc_gathered = local_tensors, project, alloc
    SH.forward_records, SH.backward_blend_records, SH.backward_geom_local = forward, backward_blend, geom_local


def _flow_worker(rank, world, port, P, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import street_gaussians_b200 as sgb
        from street_gaussians_b200 import sharded as SH
        H, W = 72, 8  # 5 tile rows, the last one partial
        _install_toy_steps(SH, H, W, rank, world)
        gen = torch.Generator().manual_seed(5)
        means, opac, sem = torch.randn(P, 3, generator=gen), torch.rand(P, 1, generator=gen), torch.rand(P, S, generator=gen)
        gcf, gdf, gsf = (torch.randn(c, H, W, generator=gen) for c in (3, 1, S))
        st = sgb.GaussianRasterizationSettings(H, W, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
        chunk = (P + world - 1) // world
        lo, hi = min(P, rank * chunk), min(P, (rank + 1) * chunk)
        rast = SH.GaussianShardedRasterizer(st)
        m = means[lo:hi].clone().requires_grad_(True)
        o = opac[lo:hi].clone().requires_grad_(True)
        s_loc = sem[lo:hi].clone().requires_grad_(True) if S else None
        n = hi - lo
        color, radii, depth, alpha, semantic = rast(means3D=m, means2D=torch.zeros(n, 3, requires_grad=True), opacities=o,
                                                    shs=torch.zeros(n, 1, 3), scales=torch.ones(n, 3), rotations=torch.ones(n, 4), semantics=s_loc)
        assert rast.chunk == chunk and radii.shape == (n,) and bool((radii == 1).all())
        rows = SH.band_of_rows(H, rank, world)
        Wt = _toy_weights(chunk * world, H)[:P]
        full = lambda v: (v.t() @ Wt).view(-1, H, 1).expand(-1, H, W)
        for got, want in ((color, full(means)), (depth, full(opac)), (alpha, full(torch.ones(P, 1)))) + (((semantic, full(sem)),) if S else ()):
            assert torch.allclose(got[:, rows], want[:, rows], atol=1e-5) and float(got[:, ~rows].abs().max()) == 0.0
        mask = rows.view(1, H, 1).float()
        outs, ups = [color, depth], [gcf * mask, gdf * mask]
        if S:
            outs.append(semantic); ups.append(gsf * mask)
        torch.autograd.backward(outs, ups)
        assert torch.allclose(m.grad, (Wt @ gcf.sum(dim=2).t())[lo:hi], atol=1e-4)      # summed over ALL ranks' rows
        assert torch.allclose(o.grad[:, 0], (Wt @ gdf.sum(dim=2)[0])[lo:hi], atol=1e-4)
        if S:
            assert torch.allclose(s_loc.grad, (Wt @ gsf.sum(dim=2).t())[lo:hi], atol=1e-4)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,S", [(7, 2), (1, 0), (10, 0)])
def test_gaussian_sharded_data_flow_gloo_world2(P, S):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flow_worker, args=(r, world, port, P, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
