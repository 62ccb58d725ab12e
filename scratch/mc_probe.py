import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((1024, 128), dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
if rank == 0:
    print("attrs", [a for a in dir(h) if not a.startswith("_")])
    print("multicast_ptr", getattr(h, "multicast_ptr", None), "buffer_ptrs", h.buffer_ptrs)
dist.destroy_process_group()
