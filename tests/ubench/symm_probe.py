"""Feasibility probe (2 GPUs): torch symmetric memory rendezvous, peer pointers, multicast support, device barrier."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; world", hdl.world_size, "multicast", hdl.has_multicast_support, "mc_ptr", hex(hdl.multicast_ptr) if hdl.has_multicast_support else None,
      "ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
t.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[123]), flush=True)
hdl.barrier(channel=0)
dist.destroy_process_group()
