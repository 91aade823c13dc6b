"""Probe: does torch's symmetric memory (peer-mapped buffers over NVLink) work in this environment?  torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
t = symm.empty(1 << 16, dtype=torch.float32, device=torch.device("cuda", local))
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
      "multicast", hex(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else None, "signal_pad_size", hdl.signal_pad_size, flush=True)
t.fill_(rank + 1.0)
torch.cuda.synchronize(); dist.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 16,), torch.float32)
print(rank, "peer value", float(peer[0]), "p2p can access", torch.cuda.can_device_access_peer(local, (local + 1) % world), flush=True)
dist.barrier()
dist.destroy_process_group()
