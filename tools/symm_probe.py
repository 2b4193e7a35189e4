"""Probe: does torch's symmetric memory (CUDA peer mappings over NVLink) work on this box?  torchrun, 2+ ranks."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((1024,), dtype=torch.float64, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; world", hdl.world_size, "rank", hdl.rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float64)
torch.cuda.synchronize()
print(rank, "peer value", float(peer[0]), "multicast_ptr", getattr(hdl, "multicast_ptr", None), flush=True)
hdl.barrier()
torch.cuda.synchronize()
# barrier latency
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    hdl.barrier()
e1.record(); torch.cuda.synchronize()
print(rank, "barrier us", 10 * e0.elapsed_time(e1), flush=True)
dist.destroy_process_group()
