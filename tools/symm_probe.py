"""Probe torch symmetric memory under torchrun: rendezvous, peer pointers, device barrier, a P2P read of the peer's buffer."""
import os
import torch
import torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
import torch.distributed._symmetric_memory as symm
t = symm.empty((1024,), dtype=torch.float32, device="cuda")
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs][:world], hdl.rank, hdl.world_size, flush=True)
t.fill_(float(rank + 1))
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
g = torch.cuda.CUDAGraph()
out = torch.empty(1024, device="cuda")
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        hdl.barrier()
        out.copy_(peer)
        hdl.barrier()
    g.replay(); torch.cuda.synchronize()
    print(rank, "graph capture of barrier + peer copy ok", float(out[0]), flush=True)
except Exception as e:  # noqa: BLE001
    print(rank, "graph capture failed:", repr(e)[:200], flush=True)
dist.destroy_process_group()
