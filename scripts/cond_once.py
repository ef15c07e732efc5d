"""One warm-up + one forward of PoseGuider and CameraPoseEncoder.forward_cameras at (1, ., 24, 768, 576): the process ncu captures the
small-channel / Pluecker / PixelUnshuffle kernels from (scripts/gpu_r2_ncu_rest.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanvid_b200 as hv  # noqa: E402
from bench import synthetic_init_  # noqa: E402

dev = torch.device("cuda", 0)
F_, H_, W_ = 24, 768, 576
pg = hv.PoseGuider(320, block_out_channels=(16, 32, 96, 256)).to(dev, torch.float16)
synthetic_init_(pg, 11, dev)
pg.refresh_native()
cam = hv.CameraPoseEncoder(downscale_factor=8, channels=[320], nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False, compression_factor=1,
                           temporal_attention_nhead=8, attention_block_types=["Temporal_Self"], temporal_position_encoding=True,
                           temporal_position_encoding_max_len=24).to(dev, torch.float16)
synthetic_init_(cam, 13, dev)
cam.refresh_native()
img = torch.rand(1, 3, F_, H_, W_, device=dev).half()
K = torch.tensor([[[1.788079 * H_, 1.788079 * H_, 0.5 * W_, 0.5 * H_]]], device=dev).repeat(1, F_, 1)
c2w = torch.eye(4, device=dev).repeat(1, F_, 1, 1)
for _ in range(2):
    y = pg(img)
    z = cam.forward_cameras(K, c2w, H_, W_)
torch.cuda.synchronize()
print("ok", tuple(y.shape), [tuple(t.shape) for t in z], bool(torch.isfinite(y).all()), bool(torch.isfinite(z[0]).all()))
