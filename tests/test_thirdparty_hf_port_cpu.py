"""Third-party cross-check of the diffusers-owned primitives the oracle restates (diffusers 0.24.0 is not installed, "parity unpinned").

TVM's ``relax.frontend.nn`` ships its own port of the HuggingFace layers (``Timesteps`` / ``get_timestep_embedding``, ``TimestepEmbedding``,
``Attention`` -- written by the TVM authors for their Stable-Diffusion support, vendored in this image under tilelang/3rdparty/tvm).  It is NOT
diffusers, but it is a restatement by a different hand: the test exports those modules to Relax IR (no code generation: this TVM build has no
LLVM), evaluates the dataflow graph with a dozen torch ops, and requires the oracle's ``timestep_sincos`` / ``TimestepEmbedding`` /
``Attention`` (self- and one-key cross-attention) to agree on the same weights, and the parameter names / shapes to be identical.
Skipped when the vendored TVM is not importable.
"""
import math

import pytest
import torch

from oracle import hv_oracle as O


def _tvm():
    """The vendored TVM, imported only when a test of this file actually runs (collection under `-m gpu` on the GPU box must not load it)."""
    try:
        import tilelang  # noqa: F401  (registers the vendored tvm package)
        from tilelang import tvm  # noqa: F401
        from tvm import relax
        from tvm.relax.frontend import nn as tnn
        from tvm.relax.frontend.nn import modules as TM
        from tvm.relax.frontend.nn import spec as tspec
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vendored TVM not importable: {e}")
    return relax, tnn, TM, tspec


def _eval_relax(func, inputs):
    """Evaluate a Relax function whose body is one dataflow block of the ops below; inputs: list of torch tensors in parameter order."""
    relax = _tvm()[0]
    env = {p: v for p, v in zip(func.params, inputs)}
    assert len(func.params) == len(inputs)

    def val(e):
        if isinstance(e, relax.Var):
            return env[e]
        if isinstance(e, relax.Constant):
            return torch.from_numpy(e.data.numpy())
        if isinstance(e, relax.PrimValue):
            return int(e.value)
        if isinstance(e, relax.ShapeExpr):
            return [int(v) for v in e.values]
        if isinstance(e, relax.Tuple):
            return [val(f) for f in e.fields]
        raise NotImplementedError(type(e))

    def call(c):
        name, a = c.op.name, [val(x) for x in c.args]
        at = c.attrs
        if name == "relax.permute_dims":
            axes = list(range(a[0].dim()))[::-1] if at.axes is None else [int(v) for v in at.axes]
            return a[0].permute(axes)
        if name == "relax.matmul":
            return a[0] @ a[1]
        if name == "relax.add":
            return a[0] + a[1]
        if name == "relax.multiply":
            return a[0] * a[1]
        if name == "relax.divide":
            return a[0] / a[1]
        if name == "relax.reshape":
            return a[0].reshape(a[1])
        if name == "relax.astype":
            assert str(at.dtype) == "float32"
            return a[0].float()
        if name == "relax.expand_dims":
            out = a[0]
            for ax in [int(v) for v in at.axis]:
                out = out.unsqueeze(ax)
            return out
        if name == "relax.arange":
            return torch.arange(a[0], a[1], a[2], dtype=torch.float32)
        if name == "relax.exp":
            return torch.exp(a[0])
        if name == "relax.cos":
            return torch.cos(a[0])
        if name == "relax.sin":
            return torch.sin(a[0])
        if name == "relax.concat":
            return torch.cat(a[0], dim=int(at.axis))
        if name == "relax.nn.silu":
            return torch.nn.functional.silu(a[0])
        if name == "relax.nn.attention":      # (batch, seq, heads, dim) layout; scale None = 1 / sqrt(dim); no mask
            assert at.scale is None and at.causal_mask is None and len(a) == 3
            q, k, v = (t.transpose(1, 2) for t in a)
            s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
            return (torch.softmax(s, dim=-1) @ v).transpose(1, 2)
        raise NotImplementedError(name)

    body = func.body
    assert len(body.blocks) == 1
    for b in body.blocks[0].bindings:
        env[b.var] = call(b.value) if isinstance(b.value, relax.Call) else val(b.value)
    return env[body.body]


@pytest.fixture(scope="module")
def exported():
    _, tnn, TM, tspec = _tvm()

    class _HFPort(tnn.Module):
        def __init__(self):
            self.ts = TM.Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)     # unet_3d.py:93
            self.emb = TM.TimestepEmbedding(320, 1280)                                     # unet_3d.py:96
            self.self_attn = TM.Attention(query_dim=320, heads=8, dim_head=40)             # attention.py:321-329
            self.cross_attn = TM.Attention(query_dim=320, cross_attention_dim=768, heads=8, dim_head=40)   # attention.py:337-345

        def sincos(self, t):
            return self.ts(t)

        def embed(self, x):
            return self.emb(x)

        def attend_self(self, x):
            return self.self_attn(x)

        def attend_cross(self, x, e):
            return self.cross_attn(x, e)

    m = _HFPort()
    mod, params = m.export_tvm(spec={
        "sincos": {"t": tspec.Tensor([3], "float32")},
        "embed": {"x": tspec.Tensor([3, 320], "float32")},
        "attend_self": {"x": tspec.Tensor([2, 24, 320], "float32")},
        "attend_cross": {"x": tspec.Tensor([2, 24, 320], "float32"), "e": tspec.Tensor([2, 1, 768], "float32")},
    })
    return mod, [(n, tuple(int(s) for s in p.shape)) for n, p in params]


def _weights(names_shapes, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {n: torch.randn(s, generator=g) * (0.5 / math.sqrt(s[-1])) for n, s in names_shapes}


def test_parameter_names_and_shapes_match_the_independent_port(exported):
    _, ps = exported
    port = dict(ps)
    ora = {}
    for prefix, mod in (("emb", O.TimestepEmbedding(320, 1280)), ("self_attn", O.Attention(320, heads=8, dim_head=40)),
                        ("cross_attn", O.Attention(320, cross_attention_dim=768, heads=8, dim_head=40))):
        ora.update({f"{prefix}.{k}": tuple(v.shape) for k, v in mod.state_dict().items()})
    assert ora == port        # to_q / to_k / to_v without bias, to_out.0 with bias, linear_1 / linear_2
    assert _tvm()[2].Attention(query_dim=320, heads=8, dim_head=40).scale == 40 ** -0.5


def test_timestep_sincos_matches_the_independent_port(exported):
    mod, ps = exported
    w = _weights(ps)
    t = torch.tensor([999.0, 519.0, 39.0])
    got = _eval_relax(mod["sincos"], [t] + [w[n] for n, _ in ps])
    ref = O.timestep_sincos(t, 320)
    assert got.shape == (3, 320)
    assert torch.allclose(got, ref, rtol=0, atol=2e-4)      # f32 sin / cos of arguments up to 999: the exponent is evaluated in a different order
    assert torch.allclose(got[:, :160] ** 2 + got[:, 160:] ** 2, torch.ones(3, 160), atol=1e-5)   # [cos | sin] halves of the same angles
    assert float(got[0, 0]) == pytest.approx(math.cos(999.0), abs=1e-4) and float(got[0, 160]) == pytest.approx(math.sin(999.0), abs=1e-4)


def test_timestep_embedding_and_attention_match_the_independent_port(exported):
    mod, ps = exported
    w = _weights(ps, seed=1)
    params = [w[n] for n, _ in ps]
    g = torch.Generator().manual_seed(2)

    emb = O.TimestepEmbedding(320, 1280)
    emb.load_state_dict({k[len("emb."):]: v for k, v in w.items() if k.startswith("emb.")})
    x = torch.randn(3, 320, generator=g)
    assert torch.allclose(_eval_relax(mod["embed"], [x] + params), emb(x), rtol=1e-5, atol=1e-5)

    sa = O.Attention(320, heads=8, dim_head=40)
    sa.load_state_dict({k[len("self_attn."):]: v for k, v in w.items() if k.startswith("self_attn.")})
    h = torch.randn(2, 24, 320, generator=g)
    assert torch.allclose(_eval_relax(mod["attend_self"], [h] + params), sa(h), rtol=1e-5, atol=1e-5)

    ca = O.Attention(320, cross_attention_dim=768, heads=8, dim_head=40)
    ca.load_state_dict({k[len("cross_attn."):]: v for k, v in w.items() if k.startswith("cross_attn.")})
    e = torch.randn(2, 1, 768, generator=g)
    got = _eval_relax(mod["attend_cross"], [h, e] + params)
    assert torch.allclose(got, ca(h, e), rtol=1e-5, atol=1e-5)
    # one key: softmax == 1, the output is to_out(to_v(e)) for every query (the collapse the native path uses, misc.cu small_linear)
    collapsed = ca.to_out[0](ca.to_v(e)).expand(-1, 24, -1)
    assert torch.allclose(got, collapsed, rtol=1e-5, atol=1e-5)
