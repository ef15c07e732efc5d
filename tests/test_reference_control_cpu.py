"""The reference's own ReferenceAttentionControl finds its blocks with isinstance() against src/models/attention.py's classes
(mutual_self_attention.py:284-300, 321-330).  A native UNet built inside the reference tree must therefore present blocks that
pass that test, or reader.update(writer) would silently zip nothing.  Both tests run in a subprocess
because they plant modules named ``src.models.attention`` / ``diffusers`` in sys.modules."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

COMMON = """
import sys, types
sys.path.insert(0, %r)
import torch, torch.nn as nn
MM = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"], temporal_position_encoding=True,
          temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)
def build(hv):
    unet = hv.UNet3DConditionModel(block_out_channels=(32, 64, 64, 64), cross_attention_dim=32, use_motion_module=True, use_inflated_groupnorm=True,
                                   motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=True, motion_module_type="Vanilla", motion_module_kwargs=MM)
    wr = hv.UNet2DConditionModel(block_out_channels=(32, 64, 64, 64), cross_attention_dim=32)
    return unet, wr
""" % ROOT


def _run(body):
    r = subprocess.run([sys.executable, "-c", COMMON + textwrap.dedent(body)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_blocks_adopt_a_loaded_reference_class():
    out = _run("""
        fake = types.ModuleType("src.models.attention")
        class TemporalBasicTransformerBlock(nn.Module):
            def __init__(self, dim, heads, head_dim):      # a constructor the shells could never call
                super().__init__()
        class BasicTransformerBlock(nn.Module):
            def __init__(self, dim, heads, head_dim):
                super().__init__()
        fake.TemporalBasicTransformerBlock, fake.BasicTransformerBlock = TemporalBasicTransformerBlock, BasicTransformerBlock
        for n in ("src", "src.models"):
            sys.modules.setdefault(n, types.ModuleType(n))
        sys.modules["src.models.attention"] = fake
        import humanvid_b200 as hv
        unet, wr = build(hv)
        def dfs(m):
            out = [m]
            for c in m.children():
                out += dfs(c)
            return out
        readers = [m for m in dfs(unet) if isinstance(m, fake.TemporalBasicTransformerBlock)]
        writers = [m for m in dfs(wr) if isinstance(m, fake.BasicTransformerBlock)]
        assert len(readers) == 16 and len(writers) == 16, (len(readers), len(writers))
        assert sorted(readers, key=lambda m: -m.norm1.normalized_shape[0]) == unet.reader_blocks()
        assert all(isinstance(m, hv.TemporalBasicTransformerBlock) for m in readers)
        assert set(unet.state_dict()) == set(build(hv)[0].state_dict())
        print("adopted", len(readers))
    """)
    assert "adopted 16" in out


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_the_references_own_control_drives_the_native_unets():
    out = _run("""
        from oracle import pin_against_reference as P
        P.install_stubs()
        P.install_stubs_2d()
        from src.models.mutual_self_attention import ReferenceAttentionControl as RefControl   # the reference's class, unmodified
        import src.models.attention as ref_attn
        import humanvid_b200 as hv
        unet, wr = build(hv)
        assert all(isinstance(b, ref_attn.TemporalBasicTransformerBlock) for b in unet.reader_blocks())
        writer = RefControl(wr, do_classifier_free_guidance=True, mode="write", batch_size=1, fusion_blocks="full")
        reader = RefControl(unet, do_classifier_free_guidance=True, mode="read", batch_size=1, fusion_blocks="full")
        # stand in for the writer forward: block i of the writer's order leaves a recognisable bank
        for i, (blk, _) in enumerate(wr.writer_blocks()):
            blk.bank.append(torch.full((2, 3, blk.norm1.normalized_shape[0]), float(i)))
        reader.update(writer)
        for i, blk in enumerate(unet.reader_blocks()):
            assert len(blk.bank) == 1 and blk.bank[0].dtype == torch.float16 and float(blk.bank[0][0, 0, 0]) == i, i
        reader.clear()
        writer.clear()
        assert all(len(b.bank) == 0 for b in unet.reader_blocks())
        print("reference control ok")
    """)
    assert "reference control ok" in out
