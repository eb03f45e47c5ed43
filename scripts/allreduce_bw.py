"""All-reduce bandwidth on this box for the gradient-bucket sizes of the training step (fp32), one process per GPU:
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/allreduce_bw.py"""
import os
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for mb in (0.6, 12.6, 16.8, 119.0):
    for dtype in (torch.float32, torch.bfloat16):
        n = int(mb * 1e6 / 4)                       # the same element count in both dtypes
        x = torch.ones(n, dtype=dtype, device=dev)
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            by = n * x.element_size()
            print(f"{mb:6.1f} M-elem-equivalent MB fp32, {str(dtype):15s} {by / 1e6:7.1f} MB  {t.item() * 1e3:8.1f} us  algbw {by / t.item() / 1e6:7.1f} GB/s", flush=True)
dist.destroy_process_group()
