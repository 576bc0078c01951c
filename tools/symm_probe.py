"""Probe: torch symmetric memory on this box (peer-mapped buffers + device-side barrier) under torchrun."""
import os, sys, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((1024, 1024), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, t.shape, t.dtype)
    print(rank, "peer value", float(peer[0, 0]), "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
    # P2P write into the peer's buffer from a kernel (a copy kernel = generic stores over NVLink)
    hdl.barrier(channel=0)
    peer[1].fill_(100 + rank)
    torch.cuda.synchronize(); hdl.barrier(channel=0)
    print(rank, "row 1 now", float(t[1, 0]), flush=True)
    # timing: barrier cost
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100):
        hdl.barrier(channel=0)
    torch.cuda.synchronize()
    print(rank, "barrier us", (time.perf_counter() - t0) / 100 * 1e6, flush=True)
except Exception as e:  # noqa: BLE001
    import traceback; traceback.print_exc()
    print(rank, "symm_mem FAILED:", repr(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
