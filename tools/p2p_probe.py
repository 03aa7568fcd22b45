"""Probe (2+ GPUs, torchrun): does torch symmetric memory give peer-mapped device pointers on this box, and what
does its device-side barrier cost?"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1024 * world
t = symm_mem.empty(N, dtype=torch.float64, device="cuda")
t.zero_()
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
hdl.barrier()
# every rank writes its slice into every peer's buffer
for r in range(world):
    buf = hdl.get_buffer(r, (N,), torch.float64)
    buf[rank * 1024:(rank + 1) * 1024] = float(rank + 1)
hdl.barrier()
torch.cuda.synchronize()
ok = all(bool((t[r * 1024:(r + 1) * 1024] == float(r + 1)).all()) for r in range(world))
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
for _ in range(10): hdl.barrier()
torch.cuda.synchronize(); dist.barrier()
ev0.record()
for _ in range(100): hdl.barrier()
ev1.record(); torch.cuda.synchronize()
tb = ev0.elapsed_time(ev1) / 100
x = torch.zeros(N, dtype=torch.float64, device="cuda")
for _ in range(10): dist.all_gather_into_tensor(x, x[rank * 1024:(rank + 1) * 1024])
torch.cuda.synchronize(); dist.barrier()
ev0.record()
for _ in range(100): dist.all_gather_into_tensor(x, x[rank * 1024:(rank + 1) * 1024])
ev1.record(); torch.cuda.synchronize()
ta = ev0.elapsed_time(ev1) / 100
print("rank %d ok=%s barrier_ms=%.4f allgather_ms=%.4f" % (rank, ok, tb, ta), flush=True)
dist.barrier(); dist.destroy_process_group()
