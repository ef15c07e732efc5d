"""Multi-GPU check of humanvid_b200.distributed on real GPUs (launch under torchrun, >= 2 ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py

ONE clip of 48 frames = three 24-frame context windows, CFG, reference banks: the 6 (window x CFG-half) units of every timestep are
split over the ranks (HV_FLAG_UNCOND_ONLY / HV_FLAG_COND_ONLY forwards), exchanged with one NCCL all-gather per step, and every rank
runs the accumulate / CFG / DDIM kernel.  Checked against the single-GPU loop (CFG-doubled batch per window) on rank 0, and that the
latents stay bitwise replicated across ranks.  Then config 4's pattern: one clip per rank, one final all-gather.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humanvid_b200 as hv  # noqa: E402
from humanvid_b200.device_loop import DeviceDenoiseLoop  # noqa: E402
from humanvid_b200.distributed import UnitExchange, gather_clip_latents, unit_list  # noqa: E402
from humanvid_b200.pipeline import uniform  # noqa: E402
from oracle import hv_oracle as O  # noqa: E402

MM_KW = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
             temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    chs, xdim, F, H, W = (64, 128, 256, 256), 64, 48, 16, 16
    ora = O.synthetic_init(O.UNet3DConditionModel(block_out_channels=chs, cross_attention_dim=xdim).eval(), seed=7)
    unet = hv.UNet3DConditionModel(block_out_channels=chs, cross_attention_dim=xdim, use_motion_module=True, use_inflated_groupnorm=True,
                                   motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=True, motion_module_type="Vanilla",
                                   motion_module_kwargs=MM_KW, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    unet.load_state_dict(ora.state_dict())
    unet = unet.to(dev, torch.float16)
    g = torch.Generator(device=dev).manual_seed(42)          # the same clip on every rank
    lat = torch.randn(1, 4, F, H, W, generator=g, device=dev).half()
    ehs = torch.randn(2, 1, xdim, generator=g, device=dev).half()
    ehs[:1] = 0
    windows = list(uniform(0, 2, F, 24, 1, 4))
    conds = [(torch.randn(1, chs[0], 24, H, W, generator=g, device=dev) * 0.5).half().repeat(2, 1, 1, 1, 1) for _ in windows]
    hv.ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
    for blk, (l, c) in zip(unet.reader_blocks(), O.bank_shapes(ora, H, W)):
        blk.bank = [torch.randn(2, l, c, generator=g, device=dev).half()]
    sched = hv.DDIMScheduler()
    sched.set_timesteps(2)
    ex = UnitExchange(unit_list(len(windows), True))
    loop = DeviceDenoiseLoop(unet, sched, lat, windows, ehs, conds, 3.5, True, exchange=ex, my_units=ex.my_units, all_units=ex.units)
    out = loop.run(2).clone()
    loop.close()
    # replicated: every rank holds the same latents, bit for bit
    allr = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(allr, out)
    same = all(torch.equal(allr[0], t) for t in allr)
    # single-GPU loop (CFG-doubled batch per window, CUDA graph) on every rank
    loop1 = DeviceDenoiseLoop(unet, sched, lat, windows, ehs, conds, 3.5, True)
    ref = loop1.run(2).clone()
    loop1.close()
    e = rel(out, ref)
    clips = gather_clip_latents(out[:, :, :2] * (rank + 1))
    ok_clips = all(torch.equal(clips[k], out[0, :, :2] * (k + 1)) for k in range(world))
    if rank == 0:
        print(f"dist_check world={world}: units per rank {[len(a) for a in ex.assignment]}; replicated={same}; unit split vs single-GPU loop rel {e:.2e}; "
              f"clip gather ok={ok_clips}", flush=True)
    assert same and ok_clips and e < 2e-3 and torch.isfinite(out).all()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
