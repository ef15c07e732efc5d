"""BASELINE config 2/3 at full size on a B200: (2,4,24,96,72) fp16 latents, full-width random-init UNet (1.31 B params).

* parity proper: native vs the oracle run on the same GPU in fp32 (reference arithmetic) and in fp16 eager (the
  reference's own deployment mode, `pipe.to("cuda", torch.float16)`);
* size-independent properties: determinism, batch-item independence, CFG/bank isolation of the unconditional half.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import report

if torch.cuda.is_available():
    import humanvid_b200 as hv
    from oracle import hv_oracle as O

MM_KW = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
             temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)
CH, XDIM, F, H, W = (320, 640, 1280, 1280), 768, 24, 96, 72


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(scope="module")
def full():
    ora = O.UNet3DConditionModel(block_out_channels=CH, cross_attention_dim=XDIM).eval()
    O.synthetic_init(ora, seed=7)
    ora = ora.half().cuda()
    nat = hv.UNet3DConditionModel(block_out_channels=CH, cross_attention_dim=XDIM, use_motion_module=True, use_inflated_groupnorm=True,
                                  motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=True, motion_module_type="Vanilla",
                                  motion_module_kwargs=MM_KW, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    nat.load_state_dict(ora.state_dict())
    nat = nat.to("cuda", torch.float16)
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn(2, 4, F, H, W, generator=g, device="cuda").half()
    ehs = torch.randn(2, 1, XDIM, generator=g, device="cuda").half()
    ehs[:1] = 0
    pose = (torch.randn(2, CH[0], F, H, W, generator=g, device="cuda") * 0.5).half()
    return ora, nat, x, ehs, pose


def test_config2_parity_full_size(full):
    ora, nat, x, ehs, pose = full
    t = 519
    with torch.no_grad():
        yn = nat(x, t, ehs, pose_cond_fea=pose, return_dict=False)[0]
        y16 = ora(x, torch.tensor(t, device="cuda"), ehs, pose_cond_fea=pose)[0]
        ora.float()
        y32 = ora(x.float(), torch.tensor(t, device="cuda"), ehs.float(), pose_cond_fea=pose.float())[0]
        ora.half()
    torch.cuda.synchronize()
    e_ref, e_nat, e_pair = rel(y16, y32), rel(yn, y32), rel(yn, y16)
    report(f"config2 full size (2,4,24,96,72): fp16-eager vs fp32 {e_ref:.2e}; native vs fp32 {e_nat:.2e}; native vs fp16-eager {e_pair:.2e}")
    assert torch.isfinite(yn).all()
    # north_star: <= 1e-3 per tensor.  Per block on identical inputs it is 3e-4 (tests/test_ladder_gpu.py, profiles/r02_error_ladder_config2.txt);
    # accumulated over the ~65 sequential fp16 tensors of a forward the storage roundings alone reach 1.4e-3 (the reference's own fp16 path:
    # 1.7e-3), so the network-level bar is "closer to fp32 than the reference's deployment, and below 1.6e-3" (measured 1.40e-3 / 1.70e-3 / 1.98e-3)
    assert e_nat <= e_ref and e_nat <= 1.6e-3
    assert e_pair <= 2.5e-3
    # per-frame breakdown: no single frame may be an outlier
    per_frame = [(rel(yn[:, :, f], y32[:, :, f])) for f in range(F)]
    assert max(per_frame) <= 3 * (sum(per_frame) / F)


def test_config2_determinism_and_batch_independence(full):
    _, nat, x, ehs, pose = full
    with torch.no_grad():
        a = nat(x, 999, ehs, pose_cond_fea=pose, return_dict=False)[0].clone()
        b = nat(x, 999, ehs, pose_cond_fea=pose, return_dict=False)[0].clone()
        x2 = x.clone()
        x2[1] = x2[1].flip(-1)  # change only batch item 1
        c = nat(x2, 999, ehs, pose_cond_fea=pose, return_dict=False)[0]
    assert torch.equal(a, b)
    assert torch.equal(a[0], c[0])
    assert not torch.equal(a[1], c[1])


def test_config3_banks_full_size(full):
    ora, nat, x, ehs, pose = full
    g = torch.Generator(device="cuda").manual_seed(5)
    banks = [torch.randn(2, l, c, generator=g, device="cuda").half() for (l, c) in O.bank_shapes(ora, H, W)]
    assert [tuple(b.shape[1:]) for b in banks][:6] == [(432, 1280)] * 5 + [(108, 1280)]
    with torch.no_grad():
        plain = nat(x, 39, ehs, pose_cond_fea=pose, return_dict=False)[0].clone()
        ctl = hv.ReferenceAttentionControl(nat, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
        for blk, bk in zip(nat.reader_blocks(), banks):
            blk.bank = [bk]
        O.set_reference_banks(ora, banks, cfg=True)
        yn = nat(x, 39, ehs, pose_cond_fea=pose, return_dict=False)[0].clone()
        y16 = ora(x, torch.tensor(39, device="cuda"), ehs, pose_cond_fea=pose)[0]
        O.set_reference_banks(ora, None)
        ctl.clear()
    torch.cuda.synchronize()
    report(f"config3 full size (16 banks): native vs fp16-eager oracle {rel(yn, y16):.2e}")
    assert rel(yn, y16) <= 2.5e-3
    assert torch.equal(yn[:1], plain[:1])          # unconditional half never sees the bank
    assert rel(yn[1:], plain[1:]) > 1e-2           # conditional half does
