"""Probe (torchrun, N >= 2): does torch's CUDA symmetric memory work on this box, and what do a peer copy and a
device-side barrier cost?  Decides whether the Gaussian-sharded exchange can write records straight into peer memory
over NVLink instead of going through NCCL collectives."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        import torch.distributed._symmetric_memory as symm
        n = 64 << 20
        t = symm.empty(n // 4, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        t.fill_(float(rank + 1))
        hdl.barrier(channel=0)
        peer = (rank + 1) % world
        remote = hdl.get_buffer(peer, (n // 4,), torch.float32)
        ok = float(remote[12345].item()) == float(peer + 1)
        print(f"[rank {rank}] symm mem OK={ok} ptrs={[hex(p) for p in hdl.buffer_ptrs]} multicast={hdl.has_multicast_support} "
              f"mc_ptr={hex(hdl.multicast_ptr) if hdl.has_multicast_support else None} signal_pad={hdl.signal_pad_size}", flush=True)
        dst = torch.empty_like(t)
        for name, fn in (("pull (local <- peer)", lambda: dst.copy_(remote)), ("push (peer <- local)", lambda: remote.copy_(dst))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print(f"[rank {rank}] {name}: 64 MiB in {ms:.3f} ms = {n / ms / 1e6:.0f} GB/s", flush=True)
        hdl.barrier(channel=0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            hdl.barrier(channel=0)
        b.record()
        torch.cuda.synchronize()
        print(f"[rank {rank}] symm barrier: {a.elapsed_time(b) * 10:.1f} us each", flush=True)
        x = torch.zeros(1, device=dev)
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        a.record()
        for _ in range(100):
            dist.all_reduce(x)
        b.record()
        torch.cuda.synchronize()
        print(f"[rank {rank}] nccl 4-byte all-reduce: {a.elapsed_time(b) * 10:.1f} us each", flush=True)
        for mb in (11, 45, 91):
            src = torch.empty(mb * (1 << 20) // 4 // world, device=dev)
            out = torch.empty(src.numel() * world, device=dev)
            for _ in range(3):
                dist.all_gather_into_tensor(out, src)
            torch.cuda.synchronize()
            a.record()
            for _ in range(10):
                dist.all_gather_into_tensor(out, src)
            b.record()
            torch.cuda.synchronize()
            ag = a.elapsed_time(b) / 10
            a.record()
            for _ in range(10):
                dist.reduce_scatter_tensor(src, out)
            b.record()
            torch.cuda.synchronize()
            rs = a.elapsed_time(b) / 10
            if rank == 0:
                print(f"nccl total {mb} MiB: all_gather {ag:.3f} ms, reduce_scatter {rs:.3f} ms", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[rank {rank}] symmetric memory probe FAILED: {e!r}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
