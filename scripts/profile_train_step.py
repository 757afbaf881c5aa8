import sys, os, runpy, torch
sys.argv = ["train_step_bench.py", "floor2", "2"]
from torch.profiler import profile, ProfilerActivity
src = open("scripts/train_step_bench.py").read()
# run setup + warm-up, then profile two recorded steps
head, tail = src.split("for _ in range(6):\n    step()\n")
g = {"__name__": "__main__", "__file__": os.path.abspath("scripts/train_step_bench.py")}
exec(compile(head + "for _ in range(6):\n    step()\n", "tsb", "exec"), g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g["step"](); g["step"]()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
