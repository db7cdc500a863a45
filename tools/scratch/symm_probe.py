import os, torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1024, dtype=torch.float64, device='cuda')
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, 'symm ok', [hex(p) for p in hdl.buffer_ptrs], [hex(p) for p in hdl.signal_pad_ptrs], hdl.signal_pad_size if hasattr(hdl,'signal_pad_size') else None, flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float64)
    print(rank, 'peer value', float(peer[0]), flush=True)
    hdl.barrier()
except Exception as e:
    print(rank, 'symm FAILED', repr(e)[:500], flush=True)
dist.barrier()
dist.destroy_process_group()
