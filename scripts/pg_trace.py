"""Per-launch device times of one PoseGuider forward at (1,3,24,768,576) (hv_set_profiling / hv_dump_profile on its handle)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanvid_b200 as hv
from humanvid_b200 import _native as N
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_init_
dev = torch.device("cuda", 0)
pg = hv.PoseGuider(320, block_out_channels=(16, 32, 96, 256)).to(dev, torch.float16)
synthetic_init_(pg, 11, dev); pg.refresh_native()
x = torch.rand(1, 3, 24, 768, 576, device=dev).half()
for _ in range(2): pg(x)
torch.cuda.synchronize()
N.lib().hv_set_profiling(pg._handle, 1)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); pg(x); e1.record(); torch.cuda.synchronize()
N.lib().hv_dump_profile(pg._handle, b"gpurun_out/pg_trace.csv")
print("pose guider forward", e0.elapsed_time(e1), "ms"); print(open("gpurun_out/pg_trace.csv").read())
