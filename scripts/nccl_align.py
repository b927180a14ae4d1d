"""Does ncclAllReduce(AVG) accept 4-byte aligned fp32 slices? (torchrun, 2 ranks)"""
import os
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
flat = torch.ones(4_000_000, device="cuda") * (rank + 1)
dist.all_reduce(flat)
torch.cuda.synchronize()
for lo, hi in ((0, 1_000_003), (1_000_003, 2_500_001), (2_500_001, 4_000_000)):
    try:
        dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        print("rank", rank, "slice", lo, hi, "ok", float(flat[lo]), flush=True)
    except Exception as e:  # noqa: BLE001
        print("rank", rank, "slice", lo, hi, "FAILED", str(e)[:200], flush=True)
        break
dist.destroy_process_group()
