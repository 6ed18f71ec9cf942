"""Does this box give us NVLink-switch multicast memory?  torchrun worker (2+ GPUs): allocates a symmetric buffer through
torch.distributed._symmetric_memory (plumbing: cuMemCreate + cuMulticastCreate + handle exchange), prints the multicast / peer
pointers, and checks the library's own multimem all-reduce on it.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nvls_probe.py
"""
import os
import sys
import traceback

import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    print(f"[{rank}] backend", symm_mem.get_backend(dev), "nvshmem", symm_mem.is_nvshmem_available(), flush=True)
    n = 1 << 22
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(f"[{rank}] rendezvous ok: world {hdl.world_size} rank {hdl.rank} multicast_ptr {hdl.multicast_ptr:#x} "
          f"buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} signal_pad_ptrs {[hex(p) for p in hdl.signal_pad_ptrs]} "
          f"buffer_size {hdl.buffer_size} signal_pad_size {hdl.signal_pad_size}", flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    if hdl.multicast_ptr:
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        want = float(sum(range(1, world + 1)))
        print(f"[{rank}] multimem_all_reduce_: got {t[0].item()} / {t[-1].item()}, want {want}", flush=True)
    print(f"[{rank}] NVLS_PROBE_OK" if hdl.multicast_ptr else f"[{rank}] NVLS_PROBE_NO_MULTICAST", flush=True)
except Exception:
    traceback.print_exc()
    print(f"[{rank}] NVLS_PROBE_FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
