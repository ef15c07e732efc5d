"""Reference-facing modules of the B200-native denoising path.

``UNet3DConditionModel``, ``PoseGuider`` and ``CameraPoseEncoder`` keep the reference's constructor arguments,
``forward`` signatures, attribute names and ``state_dict`` keys, so
``scripts/pose2vid.py``-style drivers build, load and call them unchanged:

  * ``UNet3DConditionModel.forward``  <->  src/models/unet_3d.py:397-577
  * ``PoseGuider.forward``            <->  src/models/pose_guider.py:51-61
  * ``CameraPoseEncoder.forward``     <->  src/cameractrl/pose_adaptor.py:232-248
  * ``UNet2DConditionModel.forward``  <->  src/models/unet_2d_condition.py:872-1308 (the reference / "writer" UNet, whose
    LayerNorm-1 outputs become the denoising UNet's K/V banks)

The ``nn.Module`` tree below only *holds parameters* under the reference's names (torch is plumbing: device memory,
``load_state_dict``, ``.to()``); none of its sub-modules' ``forward`` is ever executed.  All arithmetic happens in
``libhv_b200.so`` (hand-written sm_100a CUDA) through the C ABI in ``include/hv_b200.h``.  There is no PyTorch or
CPU fallback: without the library, or on a non-CUDA tensor, ``forward`` raises.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import warnings
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import _native as N


# --------------------------------------------------------------------------------------------- parameter shells
class _Shell(nn.Module):
    """Parameter container; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("humanvid_b200 parameter shells are not executable; call the owning network's forward")


def _conv(cin, cout, k, **kw):
    return nn.Conv2d(cin, cout, k, **kw)


class _Attention(_Shell):
    """diffusers Attention parameter names: to_q/to_k/to_v (no bias), to_out.0 (bias)."""

    def __init__(self, dim, cross_dim=None):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(cross_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(cross_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])


class _GEGLU(_Shell):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Linear(dim, dim * 8)


class _FeedForward(_Shell):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])


class TemporalBasicTransformerBlock(_Shell):
    """Parameter shell of src/models/attention.py:298-379; ``bank`` is what ReferenceAttentionControl.update fills."""

    def __init__(self, dim, cross_attention_dim):
        nn.Module.__init__(self)   # not super(): an adopted reference class (see _adopt_reference_class) sits later in the MRO
        self.attn1 = _Attention(dim)
        self.norm1 = nn.LayerNorm(dim)
        self.attn2 = _Attention(dim, cross_attention_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.ff = _FeedForward(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.bank: List[torch.Tensor] = []


class BasicTransformerBlock(TemporalBasicTransformerBlock):
    """Writer-side block of the 2-D reference UNet (src/models/attention.py BasicTransformerBlock): same parameters."""


_ADOPTED = {}


def _adopt_reference_class(cls):
    """Inside the reference tree its own ``ReferenceAttentionControl`` finds the blocks to hook with
    ``isinstance(module, src.models.attention.TemporalBasicTransformerBlock / BasicTransformerBlock)``
    (mutual_self_attention.py:284-300, 321-330).  A native UNet built there must not silently present zero blocks (banks
    would never arrive): when ``src.models.attention`` is loaded, the shells are instantiated from a subclass that also derives
    from the reference's class of the same name, so the reference's control walks, sorts, hooks and fills ``.bank`` on them
    exactly as on its own modules.  (The hooked ``forward`` is never called; the native forward reads ``.bank``.)"""
    import sys

    ref_mod = sys.modules.get("src.models.attention")
    ref = getattr(ref_mod, cls.__name__, None) if ref_mod is not None else None
    if not (isinstance(ref, type) and issubclass(ref, nn.Module)) or issubclass(cls, ref):
        return cls
    key = (cls, ref)
    if key not in _ADOPTED:
        _ADOPTED[key] = type(cls.__name__, (cls, ref), {"__module__": cls.__module__, "__doc__": cls.__doc__})
    return _ADOPTED[key]


class _Transformer3D(_Shell):
    def __init__(self, ch, cross_attention_dim, groups, block_cls=None):
        super().__init__()
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = _conv(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList([_adopt_reference_class(block_cls or TemporalBasicTransformerBlock)(ch, cross_attention_dim)])
        self.proj_out = _conv(ch, ch, 1)


def _sinusoid(max_len, dim):
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2) * (-math.log(10000.0) / dim))
    pe = torch.zeros(1, max_len, dim)
    pe[0, :, 0::2] = torch.sin(pos * div)
    pe[0, :, 1::2] = torch.cos(pos * div)
    return pe


class _PosEnc(_Shell):
    def __init__(self, dim, max_len):
        super().__init__()
        self.register_buffer("pe", _sinusoid(max_len, dim))


class _TemporalAttention(_Attention):
    def __init__(self, dim, max_len):
        super().__init__(dim)
        self.pos_encoder = _PosEnc(dim, max_len)


class _TemporalBlock(_Shell):
    def __init__(self, dim, n_attn, max_len):
        super().__init__()
        self.attention_blocks = nn.ModuleList([_TemporalAttention(dim, max_len) for _ in range(n_attn)])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in range(n_attn)])
        self.ff = _FeedForward(dim)
        self.ff_norm = nn.LayerNorm(dim)


class _TemporalTransformer3D(_Shell):
    def __init__(self, ch, max_len, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = nn.Linear(ch, ch)
        self.transformer_blocks = nn.ModuleList([_TemporalBlock(ch, 2, max_len)])
        self.proj_out = nn.Linear(ch, ch)
        nn.init.zeros_(self.proj_out.weight)   # zero_module (motion_module.py:72-75): an unloaded motion module is the identity
        nn.init.zeros_(self.proj_out.bias)


class _MotionModule(_Shell):
    def __init__(self, ch, max_len, groups):
        super().__init__()
        self.temporal_transformer = _TemporalTransformer3D(ch, max_len, groups)


class _Resnet(_Shell):
    def __init__(self, cin, cout, temb, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = _conv(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = _conv(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = _conv(cin, cout, 1)


class _Sampler(_Shell):
    def __init__(self, ch, stride):
        super().__init__()
        self.conv = _conv(ch, ch, 3, stride=stride, padding=1)


class _Block(_Shell):
    def __init__(self, res_io, ch, attn, motion, xdim, temb, groups, max_len, sampler=None, block_cls=None):
        super().__init__()
        if attn:
            self.attentions = nn.ModuleList([_Transformer3D(ch, xdim, groups, block_cls) for _ in range(attn)])
        self.resnets = nn.ModuleList([_Resnet(i, o, temb, groups) for i, o in res_io])
        if block_cls is None:   # the 2-D reference UNet's blocks (block_cls given) have no motion_modules attribute
            self.motion_modules = nn.ModuleList([_MotionModule(ch, max_len, groups) if motion else None for _ in range(max(attn, len(res_io)) if sampler != "mid" else 1)])
        if sampler == "down":
            self.downsamplers = nn.ModuleList([_Sampler(ch, 2)])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([_Sampler(ch, 1)])


class _TimestepEmbedding(_Shell):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor


# --------------------------------------------------------------------------------------------- native base
class _NativeNet(nn.Module):
    """Owns one ``hv_handle``; pushes the module's state_dict into it whenever parameters changed."""

    _kind = 0

    def __init__(self):
        super().__init__()
        self._handle = None
        self._pushed_versions = None
        self._epoch = 0
        self._bank_fp = None
        self._bank_refs = None
        self._reserved = set()

    def _hv_config(self) -> "HvConfig":  # pragma: no cover - overridden
        raise NotImplementedError

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def _versions(self):
        # fingerprint checked on every forward: in-place edits of ANY parameter / buffer (sum of their version counters), a
        # re-allocation or dtype change of the first one, and load_state_dict()/.to()/.half() (the epoch) all force a re-push.
        # The flat tensor list is cached per epoch: walking the module tree (~600 modules, 1270 tensors) cost 3 ms of host time per
        # forward, which the end-to-end step (copy in -> forward -> copy out, nothing to hide behind) paid in full.  A Parameter OBJECT
        # replaced by attribute assignment is not seen by the cache: call refresh_native() after such surgery (in-place edits are seen).
        cache = self.__dict__.get("_ver_cache")
        if cache is None or cache[0] != self._epoch:
            cache = (self._epoch, list(self.parameters()) + list(self.buffers()))
            self.__dict__["_ver_cache"] = cache
        tensors = cache[1]
        p = tensors[0]
        ver = 0
        for t in tensors:
            ver += t._version
        return (p.data_ptr(), ver, p.dtype, self._epoch)

    def _apply(self, fn, *a, **k):
        self._epoch = getattr(self, "_epoch", 0) + 1
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._epoch = getattr(self, "_epoch", 0) + 1
        return super().load_state_dict(*a, **k)

    def refresh_native(self):
        """Call after modifying parameters in place: the packed device copy is rebuilt on the next forward."""
        self._epoch = getattr(self, "_epoch", 0) + 1

    def _destroy(self):
        if self._handle is not None:
            N.lib().hv_destroy(self._handle)
            self._handle = None
        # a new handle may be created at the same address: nothing cached against the old one may survive it
        self._bank_fp = None
        self._bank_refs = None
        self._reserved = set()

    def _reserve(self, h, B, F, H, W):
        """The C forwards never allocate: grow the handle's private workspace here, once per new shape."""
        key = (B, F, H, W)
        if key not in self._reserved:
            N.check(N.lib().hv_reserve_workspace(h, N.i32(B), N.i32(F), N.i32(H), N.i32(W)), h)
            self._reserved.add(key)

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _sync_native(self):
        if self.device.type != "cuda":
            raise RuntimeError("humanvid_b200 runs on CUDA (B200) only: move the module with .to('cuda', torch.float16); there is no CPU path")
        ver = self._versions()
        if self._handle is not None and ver == self._pushed_versions:
            return self._handle
        self._destroy()
        lib = N.lib()
        cfg = self._hv_config()
        h = C.c_void_p()
        N.check(lib.hv_create(C.byref(cfg), C.byref(h)))
        st = N.stream()
        for k, v in self.state_dict().items():
            if v.dtype == torch.float16:
                dt = 0
            elif v.dtype == torch.float32:
                dt = 1
            else:
                raise RuntimeError(f"parameter {k} has dtype {v.dtype}; humanvid_b200 ingests fp16 or fp32 weights")
            t = v.detach().contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
            N.check(lib.hv_set_weight(h, k.encode(), N.ptr(t), shape, N.i32(max(t.dim(), 1)), N.i32(dt), st), h)
        N.check(lib.hv_finalize(h, st), h)
        torch.cuda.current_stream().synchronize()
        self._handle = h
        self._pushed_versions = ver
        return h

    @staticmethod
    def _as_half(x: torch.Tensor, what: str) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError(f"{what} must be a CUDA tensor (humanvid_b200 has no CPU path)")
        return x.to(torch.float16).contiguous()

    @property
    def last_launch_count(self) -> int:
        if self._handle is None:
            return 0
        f = N.lib().hv_last_launch_count
        f.restype = C.c_int64
        return int(f(self._handle))


class HvConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("in_channels", C.c_int32),
        ("out_channels", C.c_int32),
        ("block_out_channels", C.c_int32 * 4),
        ("heads", C.c_int32),
        ("cross_attention_dim", C.c_int32),
        ("norm_groups", C.c_int32),
        ("use_motion_module", C.c_int32),
        ("motion_max_len", C.c_int32),
        ("pg_cond_channels", C.c_int32),
        ("pg_block_channels", C.c_int32 * 4),
        ("pg_out_channels", C.c_int32),
        ("cam_downscale", C.c_int32),
        ("cam_cin", C.c_int32),
        ("cam_channels", C.c_int32),
        ("cam_nums_rb", C.c_int32),
        ("cam_heads", C.c_int32),
        ("cam_max_len", C.c_int32),
    ]


# --------------------------------------------------------------------------------------------- UNet3DConditionModel
class UNet3DConditionModel(_NativeNet):
    """Drop-in for src/models/unet_3d.py:30-577 (inference): same ctor keywords, forward signature, state_dict keys."""

    _kind = 0

    def __init__(
        self,
        sample_size: Optional[int] = None,
        in_channels: int = 4,
        out_channels: int = 4,
        center_input_sample: bool = False,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
        mid_block_type: str = "UNetMidBlock3DCrossAttn",
        up_block_types: Tuple[str, ...] = ("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
        only_cross_attention=False,
        block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280),
        layers_per_block: int = 2,
        downsample_padding: int = 1,
        mid_block_scale_factor: float = 1,
        act_fn: str = "silu",
        norm_num_groups: int = 32,
        norm_eps: float = 1e-5,
        cross_attention_dim: int = 1280,
        attention_head_dim=8,
        dual_cross_attention: bool = False,
        use_linear_projection: bool = False,
        class_embed_type=None,
        num_class_embeds=None,
        upcast_attention: bool = False,
        resnet_time_scale_shift: str = "default",
        use_inflated_groupnorm=False,
        use_motion_module=False,
        motion_module_resolutions=(1, 2, 4, 8),
        motion_module_mid_block=False,
        motion_module_decoder_only=False,
        motion_module_type=None,
        motion_module_kwargs=None,
        unet_use_cross_frame_attention=None,
        unet_use_temporal_attention=None,
    ):
        super().__init__()
        mmk = dict(motion_module_kwargs or {})
        unsupported = []
        if tuple(down_block_types) != ("CrossAttnDownBlock3D",) * 3 + ("DownBlock3D",) or tuple(up_block_types) != ("UpBlock3D",) + ("CrossAttnUpBlock3D",) * 3:
            unsupported.append("block types other than the SD1.5 layout")
        if layers_per_block != 2 or len(block_out_channels) != 4:
            unsupported.append("layers_per_block != 2 / depth != 4")
        if center_input_sample or not flip_sin_to_cos or freq_shift != 0 or use_linear_projection or dual_cross_attention:
            unsupported.append("center_input_sample / flip_sin_to_cos=False / freq_shift / use_linear_projection / dual_cross_attention")
        if class_embed_type is not None or num_class_embeds is not None or unet_use_temporal_attention or unet_use_cross_frame_attention:
            unsupported.append("class embeddings / unet_use_temporal_attention / cross_frame_attention")
        if resnet_time_scale_shift != "default" or act_fn not in ("silu", "swish") or mid_block_scale_factor != 1 or not isinstance(attention_head_dim, int):
            unsupported.append("non-default resnet/act options")
        if use_motion_module and (
            tuple(motion_module_resolutions) != (1, 2, 4, 8) or not motion_module_mid_block or motion_module_decoder_only
            or motion_module_type != "Vanilla" or mmk.get("num_transformer_block", 1) != 1
            or tuple(mmk.get("attention_block_types", ("Temporal_Self", "Temporal_Self"))) != ("Temporal_Self", "Temporal_Self")
            or not mmk.get("temporal_position_encoding", False) or mmk.get("temporal_attention_dim_div", 1) != 1
        ):
            unsupported.append("motion-module layout other than configs/inference/inference_v2.yaml")
        if unsupported:
            raise NotImplementedError("humanvid_b200.UNet3DConditionModel implements the inference_v2 hot path only; unsupported: " + "; ".join(unsupported))
        self.config = SimpleNamespace(**{k: v for k, v in locals().items() if k not in ("self", "mmk", "unsupported", "__class__")})
        self.sample_size = sample_size
        self.in_channels = in_channels
        ch = list(block_out_channels)
        heads, xdim, g = attention_head_dim, cross_attention_dim, norm_num_groups
        temb = ch[0] * 4
        mm = bool(use_motion_module)
        max_len = mmk.get("temporal_position_encoding_max_len", 24)
        self._heads, self._max_len, self._mm = heads, max_len, mm
        self.conv_in = _conv(in_channels, ch[0], 3, padding=1)
        self.time_embedding = _TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList([])
        self.mid_block = None
        self.up_blocks = nn.ModuleList([])
        prev = ch[0]
        for i, c in enumerate(ch):
            last = i == 3
            self.down_blocks.append(_Block([(prev, c), (c, c)], c, 0 if last else 2, mm, xdim, temb, g, max_len, None if last else "down"))
            prev = c
        self.mid_block = _Block([(ch[3], ch[3]), (ch[3], ch[3])], ch[3], 1, mm, xdim, temb, g, max_len, "mid")
        rev = ch[::-1]
        prev = rev[0]
        for i, c in enumerate(rev):
            cin = rev[min(i + 1, 3)]
            io = [((prev if j == 0 else c) + (cin if j == 2 else c), c) for j in range(3)]
            self.up_blocks.append(_Block(io, c, 0 if i == 0 else 3, mm, xdim, temb, g, max_len, "up" if i < 3 else None))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=norm_eps)
        self.conv_out = _conv(ch[0], out_channels, 3, padding=1)
        self._ref_cfg = True
        self._forward_flags = None   # humanvid_b200.distributed sets HV_FLAG_UNCOND_ONLY / HV_FLAG_COND_ONLY for one-half units

    # ---- reference banks (ReferenceAttentionControl, read side) -----------------------------------------------
    def reader_blocks(self) -> List[TemporalBasicTransformerBlock]:
        """fusion_blocks='full' order of mutual_self_attention.py:284-300."""

        def dfs(m):
            out = [m]
            for c in m.children():
                out += dfs(c)
            return out

        cached = self.__dict__.get("_reader_blocks_cache")
        if cached is None:      # the module tree is fixed after __init__: walk it once (1 ms per forward otherwise)
            mods = [m for m in dfs(self) if isinstance(m, TemporalBasicTransformerBlock)]
            cached = sorted(mods, key=lambda m: -m.norm1.normalized_shape[0])
            self.__dict__["_reader_blocks_cache"] = cached
        return list(cached)

    def _push_banks(self, h):
        lib = N.lib()
        blocks = self.reader_blocks()
        any_bank = any(len(b.bank) > 0 for b in blocks)
        # the banks are step-invariant (written once per clip): skip the 16 device copies when nothing changed since the last push
        # (the pushed tensors are kept referenced so that their addresses cannot be recycled under the fingerprint)
        tensors = [t for b in blocks for t in b.bank]
        fp = (h.value if hasattr(h, "value") else h, self._epoch) + tuple((t.data_ptr(), t._version, tuple(t.shape), t.dtype) for t in tensors)
        if getattr(self, "_bank_fp", None) == fp:
            return
        self._bank_fp, self._bank_refs = fp, tensors
        N.check(lib.hv_clear_ref_banks(h), h)
        if not any_bank:
            # A reader control is registered (ours, or the reference's: its register_reference_hooks leaves an instance-level `forward` on every
            # block, mutual_self_attention.py:302-330) but no bank has been written: the reference would then attend over the clip's own tokens
            # only, and so does this forward -- legal, but almost always a missing `reader.update(writer)`.  Say so once.
            hooked = getattr(self, "_reader_registered", False) or any("forward" in b.__dict__ for b in blocks)
            if hooked and not getattr(self, "_warned_no_bank", False):
                self._warned_no_bank = True
                warnings.warn("humanvid_b200.UNet3DConditionModel: a ReferenceAttentionControl reader is registered but every reference bank is "
                              "empty; this forward runs without reference attention (call reader.update(writer) after the writer forward)",
                              RuntimeWarning, stacklevel=3)
            return
        st = N.stream()
        for i, b in enumerate(blocks):
            if len(b.bank) == 0:
                continue
            if len(b.bank) != 1:
                raise NotImplementedError("one reference bank per block is supported (the pipeline writes exactly one)")
            t = self._as_half(b.bank[0], "reference bank")
            N.check(lib.hv_set_ref_bank(h, N.i32(i), N.ptr(t), N.i64(t.shape[0]), N.i64(t.shape[1]), N.i64(t.shape[2]), st), h)

    def _hv_config(self):
        c = self.config
        cfg = HvConfig()
        cfg.kind = 0
        cfg.in_channels, cfg.out_channels = c.in_channels, c.out_channels
        cfg.block_out_channels = (C.c_int32 * 4)(*c.block_out_channels)
        cfg.heads, cfg.cross_attention_dim, cfg.norm_groups = c.attention_head_dim, c.cross_attention_dim, c.norm_num_groups
        cfg.use_motion_module, cfg.motion_max_len = int(self._mm), self._max_len
        return cfg

    # ---- debug taps (per-layer error ladder; scripts/error_ladder.py) ---------------------------------------------------
    def debug_tap_plan(self, B, F, h, w):
        """[(reference module path, (NF, H, W, C))] of the activations the native forward can copy out, in execution order."""
        hd = self._sync_native()
        lib = N.lib()
        n = lib.hv_debug_tap_count(hd, N.i32(B), N.i32(F), N.i32(h), N.i32(w))
        if n < 0:
            N.check(n, hd)
        plan = []
        for i in range(n):
            name = C.create_string_buffer(128)
            dims = (C.c_int64 * 4)()
            N.check(lib.hv_debug_tap_info(hd, N.i32(i), name, N.i32(128), dims), hd)
            plan.append((name.value.decode(), tuple(int(d) for d in dims)))
        return plan

    def debug_set_taps(self, tensors):
        """tensors[i] (or None) receives tap i as channels-last fp16 (NF, H, W, C) on the next forwards; [] disables."""
        hd = self._sync_native()
        self._tap_refs = list(tensors)
        arr = (C.c_void_p * max(len(tensors), 1))(*[None if t is None else t.data_ptr() for t in tensors])
        N.check(N.lib().hv_debug_set_taps(hd, arr if tensors else None, N.i32(len(tensors))), hd)

    def workspace_bytes(self, B, F, h, w) -> int:
        hd = self._sync_native()
        f = N.lib().hv_workspace_bytes
        f.restype = C.c_size_t
        return int(f(hd, N.i32(B), N.i32(F), N.i32(h), N.i32(w)))

    @torch.no_grad()
    def forward(
        self,
        sample: torch.Tensor,
        timestep: Union[torch.Tensor, float, int],
        encoder_hidden_states: torch.Tensor,
        class_labels: Optional[torch.Tensor] = None,
        pose_cond_fea: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
        mid_block_additional_residual: Optional[torch.Tensor] = None,
        return_dict: bool = True,
    ):
        if class_labels is not None or attention_mask is not None or down_block_additional_residuals is not None or mid_block_additional_residual is not None:
            raise NotImplementedError("class_labels / attention_mask / additional residuals are not part of the CamAnimate inference path")
        if sample.dim() != 5:
            raise ValueError(f"Expected sample to have ndim=5 (b c f h w), got {sample.dim()}")
        B, Cc, F, H, W = sample.shape
        if Cc != self.in_channels:
            raise ValueError(f"sample has {Cc} channels, model expects {self.in_channels}")
        if F > 1 and not self.config.use_inflated_groupnorm:
            # resnet.py:158-161: with use_inflated_groupnorm=False the reference applies nn.GroupNorm to the 5-D tensor, pooling the
            # statistics over frames; the native path normalises per frame (InflatedGroupNorm), identical only when F == 1
            raise NotImplementedError("use_inflated_groupnorm=False with more than one frame (cross-frame GroupNorm statistics) is not implemented; "
                                      "the inference_v2 configuration sets use_inflated_groupnorm=True")
        h = self._sync_native()
        self._push_banks(h)
        x = self._as_half(sample, "sample")
        ehs = self._as_half(encoder_hidden_states, "encoder_hidden_states")
        if ehs.shape[0] != B or ehs.shape[1] != 1:
            raise ValueError(f"encoder_hidden_states must be (batch, 1, dim) with batch {B}; got {tuple(ehs.shape)}")
        pose = None if pose_cond_fea is None else self._as_half(pose_cond_fea, "pose_cond_fea")
        if pose is not None and tuple(pose.shape) != (B, self.config.block_out_channels[0], F, H, W):
            raise ValueError(f"pose_cond_fea must be {(B, self.config.block_out_channels[0], F, H, W)}, got {tuple(pose.shape)}")
        t = int(timestep.reshape(-1)[0].item()) if torch.is_tensor(timestep) else int(timestep)
        out = torch.empty((B, self.config.out_channels, F, H, W), device=x.device, dtype=torch.float16)
        flags = self._forward_flags if self._forward_flags is not None else (1 if self._ref_cfg else 0)
        self._reserve(h, B, F, H, W)
        N.check(N.lib().hv_unet3d_forward(h, N.ptr(x), N.i64(t), N.ptr(ehs), N.ptr(pose), N.ptr(out), N.i32(B), N.i32(F), N.i32(H), N.i32(W),
                                          C.c_uint32(flags), None, C.c_size_t(0), N.stream()), h)
        out = out.to(sample.dtype) if sample.dtype != torch.float16 else out
        if not return_dict:
            return (out,)
        return UNet3DConditionOutput(sample=out)

    # ---- loading (unet_3d.py:579-670) ---------------------------------------------------------------------------
    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, motion_module_path, subfolder=None, unet_additional_kwargs=None, mm_zero_proj_out=False):
        path = os.path.join(str(pretrained_model_path), subfolder) if subfolder is not None else str(pretrained_model_path)
        config_file = os.path.join(path, "config.json")
        if not (os.path.exists(config_file) and os.path.isfile(config_file)):
            raise RuntimeError(f"{config_file} does not exist or is not a file")
        with open(config_file) as f:
            cfg = json.load(f)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        cfg["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
        cfg["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
        cfg["mid_block_type"] = "UNetMidBlock3DCrossAttn"
        import inspect

        allowed = set(inspect.signature(cls.__init__).parameters)
        kw = {k: v for k, v in cfg.items() if k in allowed}
        kw.update(unet_additional_kwargs or {})
        model = cls(**kw)
        st_path = os.path.join(path, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file

            state = load_file(st_path, device="cpu")
        else:
            bin_path = os.path.join(path, "diffusion_pytorch_model.bin")
            if not os.path.isfile(bin_path):
                raise FileNotFoundError(f"no weights file found in {path}")
            state = torch.load(bin_path, map_location="cpu", weights_only=True)
        mp = str(motion_module_path)
        if os.path.exists(mp) and os.path.isfile(mp):
            suffix = os.path.splitext(mp)[1].lower()
            if suffix in (".pth", ".pt", ".ckpt"):
                mm_state = torch.load(mp, map_location="cpu", weights_only=True)
            elif suffix == ".safetensors":
                from safetensors.torch import load_file

                mm_state = load_file(mp, device="cpu")
            else:
                raise RuntimeError(f"unknown file format for motion module weights: {suffix}")
            if mm_zero_proj_out:
                mm_state = {k: v for k, v in mm_state.items() if "proj_out" not in k}
            state.update(mm_state)   # every key, like the reference (unet_3d.py:657-658); load_state_dict(strict=False) drops what the model lacks
        model.load_state_dict(state, strict=False)
        return model

# --------------------------------------------------------------------------------------------- reference ("writer") UNet
@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor


class UNet2DConditionModel(_NativeNet):
    """Drop-in for the reference ("writer") UNet, src/models/unet_2d_condition.py (SD1.5 layout), inference only.

    ``forward(sample (b,4,h,w), timestep, encoder_hidden_states (b,1,dim))`` returns what the reference returns -- the last
    up block's output (b, 320, h, w); the post-process is removed there (unet_2d_condition.py:1295-1299).  Under
    ``ReferenceAttentionControl(self, mode="write", fusion_blocks="full")`` every transformer block's LayerNorm-1 output
    is left in ``block.bank`` (mutual_self_attention.py:137-146), ready for ``reader.update(writer)``.  The native forward
    (``hv_unet2d_reference_forward``) is the denoising UNet's own kernels at one frame without motion modules.
    """

    _kind = 3

    def __init__(
        self,
        sample_size: Optional[int] = None,
        in_channels: int = 4,
        out_channels: int = 4,
        center_input_sample: bool = False,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
        mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn",
        up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
        only_cross_attention=False,
        block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280),
        layers_per_block=2,
        downsample_padding: int = 1,
        mid_block_scale_factor: float = 1,
        act_fn: str = "silu",
        norm_num_groups: Optional[int] = 32,
        norm_eps: float = 1e-5,
        cross_attention_dim=1280,
        attention_head_dim=8,
        dual_cross_attention: bool = False,
        use_linear_projection: bool = False,
        class_embed_type: Optional[str] = None,
        num_class_embeds: Optional[int] = None,
        upcast_attention: bool = False,
        resnet_time_scale_shift: str = "default",
        **unused,
    ):
        super().__init__()
        unsupported = []
        if tuple(down_block_types) != ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",) or tuple(up_block_types) != ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3 \
                or mid_block_type != "UNetMidBlock2DCrossAttn":
            unsupported.append("block types other than the SD1.5 layout")
        if layers_per_block != 2 or len(block_out_channels) != 4:
            unsupported.append("layers_per_block != 2 / depth != 4")
        if center_input_sample or not flip_sin_to_cos or freq_shift != 0 or use_linear_projection or dual_cross_attention or only_cross_attention:
            unsupported.append("center_input_sample / flip_sin_to_cos=False / freq_shift / use_linear_projection / dual / only_cross_attention")
        if class_embed_type is not None or num_class_embeds is not None or resnet_time_scale_shift != "default" or act_fn not in ("silu", "swish") \
                or mid_block_scale_factor != 1 or not isinstance(attention_head_dim, int) or norm_num_groups is None:
            unsupported.append("class embeddings / non-default resnet, activation or norm options")
        if unsupported:
            raise NotImplementedError("humanvid_b200.UNet2DConditionModel implements the SD1.5 reference-UNet layout only; unsupported: " + "; ".join(unsupported))
        self.config = SimpleNamespace(**{k: v for k, v in locals().items() if k not in ("self", "unsupported", "unused", "__class__")})
        self.sample_size = sample_size
        self.in_channels = in_channels
        ch = list(block_out_channels)
        heads, xdim, g = attention_head_dim, cross_attention_dim, norm_num_groups
        temb = ch[0] * 4
        self._heads = heads
        self.conv_in = _conv(in_channels, ch[0], 3, padding=1)
        self.time_embedding = _TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        prev = ch[0]
        for i, c in enumerate(ch):
            last = i == 3
            self.down_blocks.append(_Block([(prev, c), (c, c)], c, 0 if last else 2, False, xdim, temb, g, 0, None if last else "down", block_cls=BasicTransformerBlock))
            prev = c
        self.mid_block = _Block([(ch[3], ch[3]), (ch[3], ch[3])], ch[3], 1, False, xdim, temb, g, 0, "mid", block_cls=BasicTransformerBlock)
        rev = ch[::-1]
        prev = rev[0]
        for i, c in enumerate(rev):
            cin = rev[min(i + 1, 3)]
            io = [((prev if j == 0 else c) + (cin if j == 2 else c), c) for j in range(3)]
            self.up_blocks.append(_Block(io, c, 0 if i == 0 else 3, False, xdim, temb, g, 0, "up" if i < 3 else None, block_cls=BasicTransformerBlock))
            prev = c
        self._ref_write = False

    # ---- banks (ReferenceAttentionControl, write side) ---------------------------------------------------------
    def writer_blocks(self) -> List[Tuple[BasicTransformerBlock, int]]:
        """(block, level) in ``update()`` order: DFS(down_blocks, up_blocks, mid_block), stable sort by descending width."""
        out = []
        for i, b in enumerate(self.down_blocks):
            for a in getattr(b, "attentions", []):
                out.append((a.transformer_blocks[0], i))
        for i, b in enumerate(self.up_blocks):
            for a in getattr(b, "attentions", []):
                out.append((a.transformer_blocks[0], 3 - i))
        out.append((self.mid_block.attentions[0].transformer_blocks[0], 3))
        return sorted(out, key=lambda t: -t[0].norm1.normalized_shape[0])

    def _hv_config(self):
        c = self.config
        cfg = HvConfig()
        cfg.kind = 3
        cfg.in_channels, cfg.out_channels = c.in_channels, c.out_channels
        cfg.block_out_channels = (C.c_int32 * 4)(*c.block_out_channels)
        cfg.heads, cfg.cross_attention_dim, cfg.norm_groups = c.attention_head_dim, c.cross_attention_dim, c.norm_num_groups
        cfg.use_motion_module, cfg.motion_max_len = 0, 0
        return cfg

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int], encoder_hidden_states: torch.Tensor,
                class_labels: Optional[torch.Tensor] = None, timestep_cond=None, attention_mask=None, cross_attention_kwargs=None,
                added_cond_kwargs=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
                down_intrablock_additional_residuals=None, encoder_attention_mask=None, return_dict: bool = True):
        if any(v is not None for v in (class_labels, timestep_cond, attention_mask, cross_attention_kwargs, added_cond_kwargs, down_block_additional_residuals,
                                       mid_block_additional_residual, down_intrablock_additional_residuals, encoder_attention_mask)):
            raise NotImplementedError("only (sample, timestep, encoder_hidden_states) are part of the CamAnimate reference-UNet call")
        if sample.dim() != 4:
            raise ValueError(f"Expected sample to have ndim=4 (b c h w), got {sample.dim()}")
        B, Cc, H, W = sample.shape
        if Cc != self.in_channels:
            raise ValueError(f"sample has {Cc} channels, model expects {self.in_channels}")
        h = self._sync_native()
        x = self._as_half(sample, "sample")
        ehs = self._as_half(encoder_hidden_states, "encoder_hidden_states")
        if ehs.shape[0] != B or ehs.shape[1] != 1:
            raise ValueError(f"encoder_hidden_states must be (batch, 1, dim) with batch {B}; got {tuple(ehs.shape)}")
        t = int(timestep.reshape(-1)[0].item()) if torch.is_tensor(timestep) else int(timestep)
        blocks = self.writer_blocks()
        banks = [torch.empty((B, (H >> lvl) * (W >> lvl), blk.norm1.normalized_shape[0]), device=x.device, dtype=torch.float16) for blk, lvl in blocks]
        hidden = torch.empty((B, self.config.block_out_channels[0], H, W), device=x.device, dtype=torch.float16)
        ptrs = (C.c_void_p * len(banks))(*[b.data_ptr() for b in banks])
        self._reserve(h, B, 1, H, W)
        N.check(N.lib().hv_unet2d_reference_forward(h, N.ptr(x), N.i64(t), N.ptr(ehs), N.ptr(hidden), ptrs, N.i32(len(banks)), N.i32(B), N.i32(H), N.i32(W),
                                                    None, C.c_size_t(0), N.stream()), h)
        if self._ref_write:
            for (blk, _), bank in zip(blocks, banks):
                blk.bank.append(bank)
        hidden = hidden.to(sample.dtype) if sample.dtype != torch.float16 else hidden
        if not return_dict:
            return (hidden,)
        return UNet2DConditionOutput(sample=hidden)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, **kwargs):
        """diffusers-style loader (scripts/pose2vid.py:72-75): ``<path>/<subfolder>/config.json`` + ``diffusion_pytorch_model.{safetensors,bin}``."""
        path = os.path.join(str(pretrained_model_path), subfolder) if subfolder is not None else str(pretrained_model_path)
        with open(os.path.join(path, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        import inspect

        allowed = set(inspect.signature(cls.__init__).parameters) - {"self", "unused"}
        model = cls(**{k: v for k, v in cfg.items() if k in allowed})
        st_path = os.path.join(path, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file

            state = load_file(st_path, device="cpu")
        else:
            state = torch.load(os.path.join(path, "diffusion_pytorch_model.bin"), map_location="cpu", weights_only=True)
        model.load_state_dict(state, strict=False)   # SD checkpoints also carry conv_norm_out / conv_out, unused by the reference UNet
        return model


# --------------------------------------------------------------------------------------------- PoseGuider
class PoseGuider(_NativeNet):
    """Drop-in for src/models/pose_guider.py:16-61."""

    def __init__(self, conditioning_embedding_channels: int, conditioning_channels: int = 3, block_out_channels: Tuple[int, ...] = (16, 32, 64, 128)):
        super().__init__()
        if len(block_out_channels) != 4:
            raise NotImplementedError("PoseGuider with 4 resolution levels only")
        ch = list(block_out_channels)
        self._cfg = (conditioning_embedding_channels, conditioning_channels, ch)
        self.conv_in = _conv(conditioning_channels, ch[0], 3, padding=1)
        self.blocks = nn.ModuleList([])
        for a, b in zip(ch[:-1], ch[1:]):
            self.blocks.append(_conv(a, a, 3, padding=1))
            self.blocks.append(_conv(a, b, 3, padding=1, stride=2))
        self.conv_out = _conv(ch[-1], conditioning_embedding_channels, 3, padding=1)
        nn.init.zeros_(self.conv_out.weight)  # zero_module (pose_guider.py:42)
        nn.init.zeros_(self.conv_out.bias)

    def _hv_config(self):
        out, cin, ch = self._cfg
        cfg = HvConfig()
        cfg.kind = 1
        cfg.pg_cond_channels, cfg.pg_out_channels = cin, out
        cfg.pg_block_channels = (C.c_int32 * 4)(*ch)
        cfg.norm_groups = 32
        return cfg

    @torch.no_grad()
    def forward(self, conditioning: torch.Tensor) -> torch.Tensor:
        if conditioning.dim() != 5:
            raise ValueError("conditioning must be (b, c, f, h, w)")
        B, Cc, F, H, W = conditioning.shape
        h = self._sync_native()
        x = self._as_half(conditioning, "conditioning")
        out = torch.empty((B, self._cfg[0], F, H // 8, W // 8), device=x.device, dtype=torch.float16)
        self._reserve(h, B, F, H, W)
        N.check(N.lib().hv_pose_guider_forward(h, N.ptr(x), N.ptr(out), N.i32(B), N.i32(F), N.i32(H), N.i32(W), None, C.c_size_t(0), N.stream()), h)
        return out.to(conditioning.dtype) if conditioning.dtype != torch.float16 else out


# --------------------------------------------------------------------------------------------- CameraPoseEncoder
class _CamResnet(_Shell):
    def __init__(self, ch, ksize):
        super().__init__()
        self.block1 = _conv(ch, ch, 3, padding=1)
        self.block2 = _conv(ch, ch, ksize, padding=ksize // 2)


class CameraPoseEncoder(_NativeNet):
    """Drop-in for src/cameractrl/pose_adaptor.py:160-248 with configs/inference/inference_v2.yaml pose_encoder_kwargs."""

    def __init__(self, downscale_factor, channels=(320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False, use_conv=True,
                 compression_factor=1, temporal_attention_nhead=8, attention_block_types=("Temporal_Self",), temporal_position_encoding=False,
                 temporal_position_encoding_max_len=16, rescale_output_factor=1.0):
        super().__init__()
        channels = list(channels)
        if len(channels) != 1 or ksize != 1 or not sk or compression_factor != 1 or tuple(attention_block_types) != ("Temporal_Self",) \
                or not temporal_position_encoding or rescale_output_factor != 1.0:
            raise NotImplementedError("CameraPoseEncoder: only the inference_v2.yaml layout (one level, ksize=1, sk=True, Temporal_Self + PE) is implemented")
        c = channels[0]
        self._cfg = dict(downscale=downscale_factor, cin=cin, c=c, nums_rb=nums_rb, heads=temporal_attention_nhead, max_len=temporal_position_encoding_max_len)
        self.channels = channels
        self.nums_rb = nums_rb
        self.encoder_down_conv_blocks = nn.ModuleList([nn.ModuleList([_CamResnet(c, ksize) for _ in range(nums_rb)])])
        att = []
        for _ in range(nums_rb):
            blk = _Shell()
            blk.attention_blocks = nn.ModuleList([_TemporalAttention(c, temporal_position_encoding_max_len)])
            blk.norms = nn.ModuleList([nn.LayerNorm(c)])
            blk.ff = _FeedForward(c)
            blk.ff_norm = nn.LayerNorm(c)
            att.append(blk)
        self.encoder_down_attention_blocks = nn.ModuleList([nn.ModuleList(att)])
        zc = _conv(c, c, 1, bias=False)
        nn.init.zeros_(zc.weight)
        self.zero_conv_layers = nn.ModuleList([zc])
        self.encoder_conv_in = _conv(cin, c, 3, padding=1)

    def _hv_config(self):
        d = self._cfg
        cfg = HvConfig()
        cfg.kind = 2
        cfg.cam_downscale, cfg.cam_cin, cfg.cam_channels = d["downscale"], d["cin"], d["c"]
        cfg.cam_nums_rb, cfg.cam_heads, cfg.cam_max_len = d["nums_rb"], d["heads"], d["max_len"]
        cfg.heads = d["heads"]
        cfg.norm_groups = 32
        return cfg

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        if x.dim() != 5:
            raise ValueError("plucker embedding must be (b, c, f, h, w)")
        B, Cc, F, H, W = x.shape
        r = self._cfg["downscale"]
        if Cc * r * r != self._cfg["cin"]:
            raise ValueError(f"plucker embedding has {Cc} channels; encoder expects {self._cfg['cin'] // (r * r)}")
        h = self._sync_native()
        xin = self._as_half(x, "plucker embedding")
        out = torch.empty((B * F, self._cfg["c"], H // r, W // r), device=xin.device, dtype=torch.float16)
        self._reserve(h, B, F, H, W)
        N.check(N.lib().hv_camera_encoder_forward(h, N.ptr(xin), N.ptr(out), N.i32(B), N.i32(F), N.i32(H), N.i32(W), None, C.c_size_t(0), N.stream()), h)
        return [out.to(x.dtype) if x.dtype != torch.float16 else out]


    @torch.no_grad()
    def forward_cameras(self, intrinsics: torch.Tensor, c2w: torch.Tensor, height: int, width: int) -> List[torch.Tensor]:
        """``forward(ray_condition(intrinsics, c2w, height, width) as (b, 6, f, h, w))`` without building the embedding
        (SURVEY 8f-3): intrinsics (b, f, 4) in pixels, c2w (b, f, 4, 4), see humanvid_b200.camera.relative_cameras."""
        if intrinsics.dim() != 3 or intrinsics.shape[-1] != 4 or c2w.shape != (*intrinsics.shape[:2], 4, 4):
            raise ValueError("intrinsics must be (b, f, 4) and c2w (b, f, 4, 4)")
        B, F = intrinsics.shape[:2]
        r = self._cfg["downscale"]
        if self._cfg["cin"] != 6 * r * r:
            raise ValueError("the Plucker producer feeds a 6-channel encoder")
        h = self._sync_native()
        dev = self.device
        K = intrinsics.to(dev, torch.float32).contiguous()
        M = c2w.to(dev, torch.float32).contiguous()
        out = torch.empty((B * F, self._cfg["c"], height // r, width // r), device=dev, dtype=torch.float16)
        self._reserve(h, B, F, height, width)
        N.check(N.lib().hv_camera_encoder_forward_rays(h, N.ptr(K), N.ptr(M), N.ptr(out), N.i32(B), N.i32(F), N.i32(height), N.i32(width), None,
                                                       C.c_size_t(0), N.stream()), h)
        return [out.to(self.dtype) if self.dtype != torch.float16 else out]


# --------------------------------------------------------------------------------------------- ReferenceAttentionControl (reader)
class ReferenceAttentionControl:
    """src/models/mutual_self_attention.py for the native UNets: same constructor keywords / ``update(writer)`` / ``clear()``.

    ``mode="read"`` wraps a native ``UNet3DConditionModel`` (the denoising UNet): ``update`` hands the banks to the native
    attention kernel, which reads them as a second key/value segment (the reference monkey-patches every block's forward
    and concatenates).  ``mode="write"`` wraps a native ``UNet2DConditionModel`` (the reference UNet): its next forward
    leaves every block's LayerNorm-1 output in ``block.bank``.  ``update`` also accepts the reference's own PyTorch writer
    (anything exposing ``.unet`` whose transformer blocks carry ``norm1`` and a ``bank`` list); banks are matched in the
    same sorted order (mutual_self_attention.py:302-339).
    """

    def __init__(self, unet, mode="read", do_classifier_free_guidance=False, attention_auto_machine_weight=float("inf"),
                 gn_auto_machine_weight=1.0, style_fidelity=1.0, reference_attn=True, reference_adain=False, fusion_blocks="midup",
                 batch_size=1):
        if mode not in ("read", "write"):
            raise ValueError("mode must be 'read' or 'write'")
        if fusion_blocks != "full" or reference_adain or not reference_attn:
            raise NotImplementedError("only fusion_blocks='full', reference_attn=True (what pipeline_pose2vid_long.py uses)")
        self.unet = unet
        self.mode = mode
        if mode == "read":
            if not isinstance(unet, UNet3DConditionModel):
                raise TypeError("mode='read' needs a humanvid_b200.UNet3DConditionModel")
            unet._ref_cfg = bool(do_classifier_free_guidance)
            unet._reader_registered = True
            for m in unet.reader_blocks():
                m.bank = []
        else:
            if not isinstance(unet, UNet2DConditionModel):
                raise TypeError("mode='write' needs a humanvid_b200.UNet2DConditionModel (or use the reference's own writer around its PyTorch UNet)")
            unet._ref_write = True
            for m, _ in unet.writer_blocks():
                m.bank = []

    @staticmethod
    def _writer_blocks(writer):
        root = writer.unet if hasattr(writer, "unet") else writer
        if isinstance(root, UNet2DConditionModel):
            return [m for m, _ in root.writer_blocks()]

        def dfs(m):
            out = [m]
            for c in m.children():
                out += dfs(c)
            return out

        mods = [m for m in dfs(root) if hasattr(m, "bank") and hasattr(m, "norm1") and hasattr(m, "attn1")]
        return sorted(mods, key=lambda m: -m.norm1.normalized_shape[0])

    def update(self, writer, dtype=torch.float16):
        if self.mode != "read":
            raise RuntimeError("update() is called on the reader")
        readers = self.unet.reader_blocks()
        writers = self._writer_blocks(writer)
        for r, w in zip(readers, writers):
            r.bank = [v.clone().to(dtype) for v in w.bank]

    def clear(self):
        if self.mode == "read":
            for r in self.unet.reader_blocks():
                r.bank.clear()
        else:
            for m, _ in self.unet.writer_blocks():
                m.bank.clear()
