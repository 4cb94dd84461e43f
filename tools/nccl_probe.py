import os, sys
mode = sys.argv[1]
os.environ["NCCL_DEBUG"] = "INFO"
os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
if mode == "stderr":
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
print("done", mode, float(t[0]))
dist.destroy_process_group()
