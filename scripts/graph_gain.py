"""Does replaying one captured CUDA graph per timestep beat eager launches of the same ~810 kernels?  Config-2/3 shape, DeviceDenoiseLoop."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanvid_b200 as hv
from bench import CH, XDIM, build_native
from humanvid_b200.device_loop import DeviceDenoiseLoop

dev = torch.device("cuda", 0)
unet = build_native(dev)
g = torch.Generator(device=dev).manual_seed(1)
F, H, W = 24, 96, 72
lat = torch.randn(1, 4, F, H, W, generator=g, device=dev).half()
ehs = torch.randn(2, 1, XDIM, generator=g, device=dev).half(); ehs[:1] = 0
cond = (torch.randn(1, CH[0], F, H, W, generator=g, device=dev) * 0.5).half().repeat(2, 1, 1, 1, 1)
sched = hv.DDIMScheduler(); sched.set_timesteps(50)
loop = DeviceDenoiseLoop(unet, sched, lat, [list(range(F))], ehs, [cond], 3.5, True)
lat0 = loop.latents.clone()
for mode in ("eager", "graph", "eager", "graph"):
    loop.latents.copy_(lat0); loop.step_index.zero_()
    loop.run(3, use_graph=(mode == "graph"))
    loop.latents.copy_(lat0); loop.step_index.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); loop.run(20, use_graph=(mode == "graph")); e1.record(); torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / 20:.3f} ms per step", flush=True)
