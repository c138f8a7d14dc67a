# all ranks copy 7.6 MB device -> pinned host at the same time (torchrun): aggregate read-back bandwidth of the box
import os, time, torch, torch.distributed as dist
dist.init_process_group("nccl")
r = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(r)
n = int(7.6e6 / 4)
d = torch.empty(n, device="cuda"); h = torch.empty(n).pin_memory()
for reps in (1, 1):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(100): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if r == 0:
        ws = dist.get_world_size()
        print(f"{ws} ranks x 7.6 MB D2H: {t.item()/100*1e3:.3f} ms per copy, aggregate {ws*7.6e6*100/t.item()/1e9:.1f} GB/s", flush=True)
dist.destroy_process_group()
