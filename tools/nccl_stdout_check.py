"""stdout must stay clean when NCCL initialises at NCCL_DEBUG=VERSION (bench.py prints ONE JSON line): world-size-1 NCCL group."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["NCCL_DEBUG"] = "VERSION"
from bench import quiet_nccl_stdout
quiet_nccl_stdout()
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
import torch
import torch.distributed as dist
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
t = torch.ones(4, device="cuda")
dist.all_reduce(t)
torch.cuda.synchronize()
dist.destroy_process_group()
print("STDOUT_ONLY_LINE")
