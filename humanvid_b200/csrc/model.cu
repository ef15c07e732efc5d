// Network runtime behind include/hv_b200.h: weight ingestion under the reference's state_dict keys, packing into
// kernel layouts, a stack-allocated activation workspace, and the three forwards (UNet3DConditionModel, PoseGuider,
// CameraPoseEncoder) expressed as sequences of the operators in ops.h / kernels.h.
//
// Activations are channels-last fp16 (N = B*F frames, H, W, C) == a [tokens][C] matrix, so the reference's
// "b c f h w -> (b f) c h w", "(bf) c h w -> (bf) (hw) c" and "(b f) d c -> (b d) f c" rearranges never touch memory:
// spatial tokens are rows, frames of one pixel are rows HW apart.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hv_b200.h"
#include "gemm.cuh"
#include "kernels.h"
#include "ops.h"
#include "tma.h"

namespace hv {

struct Err {
  int code;
  std::string msg;
};
[[noreturn]] static void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Err{code, buf};
}
static void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) fail(HV_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
static void ckop(int status, const char* what) {
  if (status != HV_OK) fail(status, "%s: %s", what, last_error());
}
static inline int64_t rup(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------ tensors / weights
struct Raw {  // one state_dict entry, fp16 on device
  __half* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct Mat {  // packed [rows][cols] fp16
  __half* p = nullptr;
  int64_t rows = 0, cols = 0;
};
struct Norm {
  __half* g = nullptr;
  __half* b = nullptr;
  int C = 0;
  float eps = 1e-5f;
};
struct Lin {
  Mat w;
  __half* bias = nullptr;
};
struct Conv3 {       // implicit-GEMM packed 3x3 conv
  Mat w;             // [cout_pad][9 * cin_pad]
  __half* bias = nullptr;  // [cout_pad]
  int cin = 0, cout = 0, cin_pad = 0, cout_pad = 0;
};
struct ConvDirect {  // reference layout (Cout, Cin, 3, 3)
  __half* w = nullptr;
  __half* bias = nullptr;
  int cin = 0, cout = 0;
};

struct Tens {  // channels-last activation
  __half* p = nullptr;
  int NF = 0, H = 0, W = 0, C = 0;
  int64_t rows() const { return static_cast<int64_t>(NF) * H * W; }
  int64_t numel() const { return rows() * C; }
};

struct ResnetW {
  Norm n1, n2;
  Conv3 c1, c2;
  Lin temb;
  bool has_sc = false;
  Lin sc;
  int cin = 0, cout = 0;
};
struct SpatialW {
  Norm gn, ln1, ln3;
  Lin proj_in, proj_out;
  Mat wqk;   // [2*heads*dpad][C]  (q rows then k rows, each head zero-padded to dpad)
  Mat wk;    // [heads*dpad][C]    view into wqk (bank keys)
  Mat wv;    // [heads*dv][C]: to_v rows of head h at h*dv.., zero rows elsewhere (dv = (d+1) rounded up to 16)
  __half* vones = nullptr;  // [heads*dv]: 1 at row h*dv + d (the softmax-denominator row of V^T), else 0
  int dv = 0;
  Lin out1;
  Lin v2, out2;  // cross-attention collapse: to_v (C x xdim), to_out
  Lin ff1, ff2;  // ff1 geglu-packed
  int C = 0, heads = 0, d = 0, dpad = 0;
  int reader_idx = -1;
  // reference bank (B_ref, L, C): `bank` is non-null while a bank is set; its storage (bank_store, bank_cap elements) is kept
  // across hv_clear_ref_banks so that the per-forward clear / set of the Python reader costs no allocation
  __half* bank_store = nullptr;
  int64_t bank_cap = 0;
  __half* bank = nullptr;
  int64_t bank_B = 0, bank_L = 0;
  std::string name;
};
struct TAttnW {
  Norm ln;
  Mat wqkv;  // [3C][C]
  Lin out;
  __half* pe = nullptr;  // [max_len][C]
  int max_len = 0;
};
struct MotionW {
  Norm gn, ffn;
  Lin proj_in, proj_out, ff1, ff2;
  std::vector<TAttnW> attn;
  int C = 0, heads = 0, d = 0;
};
struct CamResW {
  Conv3 block1;
  Lin block2;
};

// ------------------------------------------------------------------------------------------ workspace arena
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;  // measuring pass: hand out offsets only, launch nothing
  void* alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    if (!dry && off + bytes > cap) fail(HV_ERR_INVALID, "workspace too small: need > %zu bytes, have %zu", off + bytes, cap);
    void* p = base + off;
    off += bytes;
    peak = std::max(peak, off);
    return p;
  }
};
struct Scope {
  Arena& a;
  size_t mark;
  explicit Scope(Arena& ar) : a(ar), mark(ar.off) {}
  ~Scope() { a.off = mark; }
};

}  // namespace hv

using namespace hv;

// ------------------------------------------------------------------------------------------ the model object
struct hv_model {
  hv_config cfg{};
  std::string err;
  std::unordered_map<std::string, Raw> raw;
  std::vector<void*> owned;  // cudaMalloc'ed blocks (weights, banks)
  bool finalized = false;
  int sms = 0;
  cudaStream_t st = nullptr;
  Arena ar;
  void* private_ws = nullptr;
  size_t private_ws_bytes = 0;
  int64_t launches = 0;
  // device-resident timestep (CUDA-graph replay of a denoising loop): t = ts_table[*ts_index] when set
  const long long* ts_table = nullptr;
  const int* ts_index = nullptr;
  // debug taps (error ladder): the activation leaving block i of the forward is copied to tap_dst[i] (channels-last fp16)
  struct TapInfo { std::string name; int NF, H, W, C; };
  std::vector<TapInfo> tap_list;
  std::vector<void*> tap_dst;
  void tap(const std::string& name, const Tens& t) {
    const size_t i = tap_list.size();
    tap_list.push_back({name, t.NF, t.H, t.W, t.C});
    if (!ar.dry && i < tap_dst.size() && tap_dst[i] != nullptr) {
      launches += 1;
      ck(cudaMemcpyAsync(tap_dst[i], t.p, static_cast<size_t>(t.numel()) * 2, cudaMemcpyDeviceToDevice, st), "tap copy");
    }
  }
  // optional per-category device timing (bench.py's roofline): events bracket every operator launch
  enum Cat { CAT_GEMM = 0, CAT_CONV = 1, CAT_ATTN = 2, CAT_TATTN = 3, CAT_NORM = 4, CAT_MISC = 5, CAT_N = 6 };
  bool profiling = false;
  struct Rec { int cat; cudaEvent_t a, b; double flops; long long m, n, k; const char* label; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  cudaEvent_t ev() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      ck(cudaEventCreate(&e), "cudaEventCreate");
      ev_pool.push_back(e);
    }
    return ev_pool[ev_used++];
  }
  struct Timed {  // RAII bracket around one operator
    hv_model* m; int cat; double flops; long long M, N, K; const char* label; cudaEvent_t a{};
    Timed(hv_model* mm, int c, double f, const char* lb = "", long long M_ = 0, long long N_ = 0, long long K_ = 0)
        : m(mm), cat(c), flops(f), M(M_), N(N_), K(K_), label(lb) {
      if (m->profiling && !m->ar.dry) { a = m->ev(); cudaEventRecord(a, m->st); }
    }
    ~Timed() {
      if (m->profiling && !m->ar.dry) { cudaEvent_t b = m->ev(); cudaEventRecord(b, m->st); m->recs.push_back({cat, a, b, flops, M, N, K, label}); }
    }
  };

  // ---- UNet
  Conv3 conv_in;  // 4 input channels zero-padded to 64: runs on the implicit-GEMM kernel
  Lin te1, te2;
  struct DownBlk { std::vector<ResnetW> res; std::vector<SpatialW> attn; std::vector<MotionW> mm; bool has_down = false; Conv3 down; };
  struct UpBlk { std::vector<ResnetW> res; std::vector<SpatialW> attn; std::vector<MotionW> mm; bool has_up = false; Conv3 up; };
  std::vector<DownBlk> down;
  struct { std::vector<ResnetW> res; std::vector<SpatialW> attn; std::vector<MotionW> mm; } mid;
  std::vector<UpBlk> up;
  Norm norm_out;
  Conv3 conv_out;
  std::vector<SpatialW*> readers;  // reference-bank order
  __half* const* bank_out = nullptr;  // writer forward: reader-order destinations of every block's LayerNorm-1 output (the "bank")
  // ---- PoseGuider
  ConvDirect pg_in;
  std::vector<Conv3> pg_convs;  // blocks.0..5, conv_out
  std::vector<int> pg_strides;
  std::vector<bool> pg_small;   // layer runs on the small-channel mma.sync kernel at its true channel counts (smallconv.cu)
  bool pg_fast_in = false;      // conv_in reads the (B,3,F,H,W) image directly and writes 16 channels (no layout pass, no padding)
  // ---- CameraPoseEncoder
  Conv3 cam_in;
  std::vector<CamResW> cam_res;
  std::vector<MotionW> cam_att;  // reuse MotionW: attn[0] + ff (no gn / proj)
  Lin cam_zero;

  ~hv_model() {
    for (void* p : owned) cudaFree(p);
    if (private_ws) cudaFree(private_ws);
    for (auto e : ev_pool) cudaEventDestroy(e);
  }

  // ================================================================== weight helpers
  __half* dmalloc(int64_t halves) {
    void* p = nullptr;
    ck(cudaMalloc(&p, static_cast<size_t>(halves) * 2), "cudaMalloc(weights)");
    owned.push_back(p);
    return static_cast<__half*>(p);
  }
  const Raw& get(const std::string& key) {
    auto it = raw.find(key);
    if (it == raw.end()) fail(HV_ERR_MISSING, "missing weight '%s'", key.c_str());
    return it->second;
  }
  bool has(const std::string& key) const { return raw.count(key) != 0; }

  __half* vec(const std::string& key, int64_t n, int64_t pad_to = 0) {
    const Raw& r = get(key);
    if (r.numel() != n) fail(HV_ERR_INVALID, "weight '%s' has %lld elements, expected %lld", key.c_str(), (long long)r.numel(), (long long)n);
    if (pad_to <= n) return r.p;
    __half* p = dmalloc(pad_to);
    ck(cudaMemsetAsync(p, 0, pad_to * 2, st), "memset");
    ck(cudaMemcpyAsync(p, r.p, n * 2, cudaMemcpyDeviceToDevice, st), "memcpy");
    return p;
  }
  Norm norm(const std::string& pfx, int C, float eps) {
    Norm n;
    n.g = vec(pfx + ".weight", C);
    n.b = vec(pfx + ".bias", C);
    n.C = C;
    n.eps = eps;
    return n;
  }
  Mat mat(const std::string& key, int64_t rows, int64_t cols) {  // 2-D (or 4-D 1x1 conv) weight used as is
    const Raw& r = get(key);
    if (r.numel() != rows * cols) fail(HV_ERR_INVALID, "weight '%s': %lld elements, expected %lld x %lld", key.c_str(), (long long)r.numel(), (long long)rows, (long long)cols);
    return Mat{r.p, rows, cols};
  }
  Lin lin(const std::string& pfx, int64_t out, int64_t in, bool bias = true) {
    Lin l;
    l.w = mat(pfx + ".weight", out, in);
    if (bias) l.bias = vec(pfx + ".bias", out);
    return l;
  }
  Conv3 conv3(const std::string& pfx, int cout, int cin, bool bias = true, bool pad_out_64 = false) {
    const Raw& r = get(pfx + ".weight");
    if (r.shape.size() != 4 || r.shape[0] != cout || r.shape[1] != cin || r.shape[2] != 3 || r.shape[3] != 3)
      fail(HV_ERR_INVALID, "weight '%s.weight' is not (%d,%d,3,3)", pfx.c_str(), cout, cin);
    Conv3 c;
    c.cin = cin;
    c.cout = cout;
    c.cin_pad = static_cast<int>(rup(cin, 64));
    c.cout_pad = static_cast<int>(pad_out_64 ? rup(cout, 64) : rup(cout, 8));  // pad_out_64: the output feeds another implicit-GEMM conv
    c.w.rows = c.cout_pad;
    c.w.cols = 9LL * c.cin_pad;
    c.w.p = dmalloc(c.w.rows * c.w.cols);
    ck(launch_pack_conv3x3(r.p, c.w.p, cout, cin, c.cout_pad, c.cin_pad, sms, st), "pack conv3x3");
    if (bias) c.bias = vec(pfx + ".bias", cout, c.cout_pad);
    return c;
  }
  // Upsample3D's conv packed for op_upconv2x2: [4 parities][cout][4 * cin]
  Conv3 upconv(const std::string& pfx, int cout, int cin) {
    const Raw& r = get(pfx + ".weight");
    if (r.shape.size() != 4 || r.shape[0] != cout || r.shape[1] != cin || r.shape[2] != 3 || r.shape[3] != 3)
      fail(HV_ERR_INVALID, "weight '%s.weight' is not (%d,%d,3,3)", pfx.c_str(), cout, cin);
    if ((cin % 64) || (cout % 8)) fail(HV_ERR_INVALID, "upsampler conv '%s': %d -> %d channels must be multiples of 64 / 8", pfx.c_str(), cin, cout);
    Conv3 c;
    c.cin = c.cin_pad = cin;
    c.cout = c.cout_pad = cout;
    c.w.rows = 4LL * cout;
    c.w.cols = 4LL * cin;
    c.w.p = dmalloc(c.w.rows * c.w.cols);
    ck(launch_pack_upconv2x2(r.p, c.w.p, cout, cin, sms, st), "pack upconv2x2");
    c.bias = vec(pfx + ".bias", cout);
    return c;
  }
  ConvDirect conv_direct(const std::string& pfx, int cout, int cin) {
    ConvDirect c;
    c.w = get(pfx + ".weight").p;
    c.bias = vec(pfx + ".bias", cout);
    c.cin = cin;
    c.cout = cout;
    return c;
  }
  Lin ff1_geglu(const std::string& pfx, int C) {  // net.0.proj: [8C][C] + bias -> 128-row hidden/gate interleave
    const Raw& w = get(pfx + ".weight");
    const Raw& b = get(pfx + ".bias");
    if (w.numel() != 8LL * C * C || b.numel() != 8LL * C) fail(HV_ERR_INVALID, "'%s' is not a GEGLU proj of dim %d", pfx.c_str(), C);
    if ((4 * C) % 128) fail(HV_ERR_INVALID, "GEGLU inner dim %d must be a multiple of 128", 4 * C);
    Lin l;
    l.w = Mat{dmalloc(8LL * C * C), 8LL * C, C};
    ck(launch_pack_geglu(w.p, l.w.p, 8 * C, C, sms, st), "pack geglu");
    l.bias = dmalloc(8LL * C);
    ck(launch_pack_geglu(b.p, l.bias, 8 * C, 1, sms, st), "pack geglu bias");
    return l;
  }

  ResnetW resnet(const std::string& pfx, int cin, int cout, int temb) {
    ResnetW r;
    r.cin = cin;
    r.cout = cout;
    r.n1 = norm(pfx + ".norm1", cin, 1e-5f);
    r.c1 = conv3(pfx + ".conv1", cout, cin);
    r.temb = lin(pfx + ".time_emb_proj", cout, temb);
    r.n2 = norm(pfx + ".norm2", cout, 1e-5f);
    r.c2 = conv3(pfx + ".conv2", cout, cout);
    r.has_sc = cin != cout;
    if (r.has_sc) r.sc = lin(pfx + ".conv_shortcut", cout, cin);
    return r;
  }
  SpatialW spatial(const std::string& pfx, int C) {
    SpatialW s;
    s.name = pfx;
    s.C = C;
    s.heads = cfg.heads;
    s.d = C / cfg.heads;
    s.dpad = static_cast<int>(rup(s.d, 16));
    if (s.d % 8) fail(HV_ERR_INVALID, "head dim %d (C=%d / %d heads) must be a multiple of 8", s.d, C, cfg.heads);
    s.gn = norm(pfx + ".norm", C, 1e-6f);
    s.proj_in = lin(pfx + ".proj_in", C, C);
    s.proj_out = lin(pfx + ".proj_out", C, C);
    const std::string b = pfx + ".transformer_blocks.0";
    s.ln1 = norm(b + ".norm1", C, 1e-5f);
    s.ln3 = norm(b + ".norm3", C, 1e-5f);
    const int64_t hp = static_cast<int64_t>(s.heads) * s.dpad;
    s.wqk = Mat{dmalloc(2 * hp * C), 2 * hp, C};
    ck(launch_pack_heads(mat(b + ".attn1.to_q.weight", C, C).p, s.wqk.p, s.heads, s.d, s.dpad, C, sms, st), "pack q");
    ck(launch_pack_heads(mat(b + ".attn1.to_k.weight", C, C).p, s.wqk.p + hp * C, s.heads, s.d, s.dpad, C, sms, st), "pack k");
    s.wk = Mat{s.wqk.p + hp * C, hp, C};
    s.dv = static_cast<int>(rup(s.d + 1, 16));
    s.wv = Mat{dmalloc(static_cast<int64_t>(s.heads) * s.dv * C), static_cast<int64_t>(s.heads) * s.dv, C};
    ck(launch_pack_heads(mat(b + ".attn1.to_v.weight", C, C).p, s.wv.p, s.heads, s.d, s.dv, C, sms, st), "pack v");
    {
      std::vector<__half> ones(static_cast<size_t>(s.heads) * s.dv, __float2half(0.f));
      for (int hh = 0; hh < s.heads; ++hh) ones[static_cast<size_t>(hh) * s.dv + s.d] = __float2half(1.f);
      s.vones = dmalloc(static_cast<int64_t>(ones.size()));
      ck(cudaMemcpyAsync(s.vones, ones.data(), ones.size() * 2, cudaMemcpyHostToDevice, st), "vones upload");
      ck(cudaStreamSynchronize(st), "vones sync");  // `ones` is a host temporary
    }
    s.out1 = lin(b + ".attn1.to_out.0", C, C);
    s.v2 = lin(b + ".attn2.to_v", C, cfg.cross_attention_dim, false);
    s.out2 = lin(b + ".attn2.to_out.0", C, C);
    s.ff1 = ff1_geglu(b + ".ff.net.0.proj", C);
    s.ff2 = lin(b + ".ff.net.2", C, 4 * C);
    return s;
  }
  TAttnW tattn(const std::string& attn_pfx, const std::string& norm_pfx, int C, int max_len) {
    TAttnW t;
    t.ln = norm(norm_pfx, C, 1e-5f);
    t.wqkv = Mat{dmalloc(3LL * C * C), 3LL * C, C};
    const char* names[3] = {".to_q.weight", ".to_k.weight", ".to_v.weight"};
    for (int i = 0; i < 3; ++i)
      ck(cudaMemcpyAsync(t.wqkv.p + static_cast<int64_t>(i) * C * C, mat(attn_pfx + names[i], C, C).p, static_cast<size_t>(C) * C * 2,
                         cudaMemcpyDeviceToDevice, st), "qkv concat");
    t.out = lin(attn_pfx + ".to_out.0", C, C);
    const Raw& pe = get(attn_pfx + ".pos_encoder.pe");
    if (pe.numel() % C) fail(HV_ERR_INVALID, "'%s.pos_encoder.pe' width != %d", attn_pfx.c_str(), C);
    t.pe = pe.p;
    t.max_len = static_cast<int>(pe.numel() / C);
    (void)max_len;
    return t;
  }
  MotionW motion(const std::string& pfx, int C) {
    MotionW m;
    m.C = C;
    m.heads = cfg.heads;
    m.d = C / cfg.heads;
    const std::string t = pfx + ".temporal_transformer";
    m.gn = norm(t + ".norm", C, 1e-6f);
    m.proj_in = lin(t + ".proj_in", C, C);
    m.proj_out = lin(t + ".proj_out", C, C);
    const std::string b = t + ".transformer_blocks.0";
    for (int i = 0; i < 2; ++i)
      m.attn.push_back(tattn(b + ".attention_blocks." + std::to_string(i), b + ".norms." + std::to_string(i), C, cfg.motion_max_len));
    m.ffn = norm(b + ".ff_norm", C, 1e-5f);
    m.ff1 = ff1_geglu(b + ".ff.net.0.proj", C);
    m.ff2 = lin(b + ".ff.net.2", C, 4 * C);
    return m;
  }

  void build_unet() {
    const int* ch = cfg.block_out_channels;
    const int temb = ch[0] * 4;
    const bool mm = cfg.use_motion_module != 0;
    conv_in = conv3("conv_in", ch[0], cfg.in_channels);
    te1 = lin("time_embedding.linear_1", temb, ch[0]);
    te2 = lin("time_embedding.linear_2", temb, temb);
    int prev = ch[0];
    down.resize(4);
    for (int i = 0; i < 4; ++i) {
      auto& d = down[i];
      const std::string p = "down_blocks." + std::to_string(i);
      for (int j = 0; j < 2; ++j) {
        d.res.push_back(resnet(p + ".resnets." + std::to_string(j), j == 0 ? prev : ch[i], ch[i], temb));
        if (i < 3) d.attn.push_back(spatial(p + ".attentions." + std::to_string(j), ch[i]));
        if (mm) d.mm.push_back(motion(p + ".motion_modules." + std::to_string(j), ch[i]));
      }
      d.has_down = i < 3;
      if (d.has_down) d.down = conv3(p + ".downsamplers.0.conv", ch[i], ch[i]);
      prev = ch[i];
    }
    mid.res.push_back(resnet("mid_block.resnets.0", ch[3], ch[3], temb));
    mid.res.push_back(resnet("mid_block.resnets.1", ch[3], ch[3], temb));
    mid.attn.push_back(spatial("mid_block.attentions.0", ch[3]));
    if (mm) mid.mm.push_back(motion("mid_block.motion_modules.0", ch[3]));
    const int rev[4] = {ch[3], ch[2], ch[1], ch[0]};
    up.resize(4);
    prev = rev[0];
    for (int i = 0; i < 4; ++i) {
      auto& u = up[i];
      const std::string p = "up_blocks." + std::to_string(i);
      const int c = rev[i], cin = rev[std::min(i + 1, 3)];
      for (int j = 0; j < 3; ++j) {
        const int rin = (j == 0 ? prev : c) + (j == 2 ? cin : c);
        u.res.push_back(resnet(p + ".resnets." + std::to_string(j), rin, c, temb));
        if (i > 0) u.attn.push_back(spatial(p + ".attentions." + std::to_string(j), c));
        if (mm) u.mm.push_back(motion(p + ".motion_modules." + std::to_string(j), c));
      }
      u.has_up = i < 3;
      if (u.has_up) u.up = upconv(p + ".upsamplers.0.conv", c, c);
      prev = c;
    }
    if (cfg.kind == HV_KIND_UNET3D) {   // the reference ("writer") 2-D UNet has its post-process removed (unet_2d_condition.py:1295-1299)
      norm_out = norm("conv_norm_out", ch[0], 1e-5f);
      conv_out = conv3("conv_out", cfg.out_channels, ch[0]);
    }
    // reader order: DFS(down_blocks, up_blocks, mid_block), stable sort by descending width
    readers.clear();
    for (auto& d : down) for (auto& a : d.attn) readers.push_back(&a);
    for (auto& u : up) for (auto& a : u.attn) readers.push_back(&a);
    for (auto& a : mid.attn) readers.push_back(&a);
    std::stable_sort(readers.begin(), readers.end(), [](const SpatialW* x, const SpatialW* y) { return x->C > y->C; });
    for (size_t i = 0; i < readers.size(); ++i) readers[i]->reader_idx = static_cast<int>(i);
  }

  // unpadded [cout][9 * cin] packing for the small-channel kernel
  Conv3 conv3_small(const std::string& pfx, int cout, int cin) {
    const Raw& r = get(pfx + ".weight");
    if (r.shape.size() != 4 || r.shape[0] != cout || r.shape[1] != cin || r.shape[2] != 3 || r.shape[3] != 3)
      fail(HV_ERR_INVALID, "weight '%s.weight' is not (%d,%d,3,3)", pfx.c_str(), cout, cin);
    Conv3 c;
    c.cin = c.cin_pad = cin;
    c.cout = c.cout_pad = cout;
    c.w.rows = cout;
    c.w.cols = 9LL * cin;
    c.w.p = dmalloc(c.w.rows * c.w.cols);
    ck(launch_pack_conv3x3(r.p, c.w.p, cout, cin, cout, cin, sms, st), "pack small conv3x3");
    c.bias = vec(pfx + ".bias", cout);
    return c;
  }

  void build_pose_guider() {
    const int* bc = cfg.pg_block_channels;
    pg_in = conv_direct("conv_in", bc[0], cfg.pg_cond_channels);
    pg_convs.clear();
    pg_strides.clear();
    pg_small.clear();
    // the 16 / 32-channel layers at (up to) full image resolution are HBM-bound: they run at their true channel counts on the small-channel
    // kernel as long as the chain from conv_in is unbroken; from the first layer it does not cover, the tcgen05 implicit GEMM takes over
    pg_fast_in = cfg.pg_cond_channels == 3 && bc[0] == 16 && smallconv_supported(bc[0], bc[0], 1);
    bool chain = pg_fast_in;
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 2; ++k) {
        const int cin = bc[i], cout = k == 0 ? bc[i] : bc[i + 1], stride = k == 0 ? 1 : 2;
        const std::string name = "blocks." + std::to_string(2 * i + k);
        chain = chain && smallconv_supported(cin, cout, stride);
        pg_small.push_back(chain);
        pg_convs.push_back(chain ? conv3_small(name, cout, cin) : conv3(name, cout, cin, true, true));
        pg_strides.push_back(stride);
      }
    }
    pg_convs.push_back(conv3("conv_out", cfg.pg_out_channels, bc[3]));
    pg_strides.push_back(1);
    pg_small.push_back(false);
  }

  void build_camera() {
    const int C = cfg.cam_channels;
    cam_in = conv3("encoder_conv_in", C, cfg.cam_cin);
    for (int j = 0; j < cfg.cam_nums_rb; ++j) {
      const std::string p = "encoder_down_conv_blocks.0." + std::to_string(j);
      CamResW r;
      r.block1 = conv3(p + ".block1", C, C);
      r.block2 = lin(p + ".block2", C, C);
      cam_res.push_back(r);
      const std::string a = "encoder_down_attention_blocks.0." + std::to_string(j);
      MotionW m;
      m.C = C;
      m.heads = cfg.cam_heads;
      m.d = C / cfg.cam_heads;
      m.attn.push_back(tattn(a + ".attention_blocks.0", a + ".norms.0", C, cfg.cam_max_len));
      m.ffn = norm(a + ".ff_norm", C, 1e-5f);
      m.ff1 = ff1_geglu(a + ".ff.net.0.proj", C);
      m.ff2 = lin(a + ".ff.net.2", C, 4 * C);
      cam_att.push_back(m);
    }
    cam_zero = lin("zero_conv_layers.0", C, C, false);
  }

  // ================================================================== op wrappers (no-ops while measuring)
  Tens alloc_act(int NF, int H, int W, int C) {
    Tens a;
    a.NF = NF; a.H = H; a.W = W; a.C = C;
    a.p = static_cast<__half*>(ar.alloc(static_cast<size_t>(a.numel()) * 2));
    return a;
  }
  __half* alloc_h(int64_t n) { return static_cast<__half*>(ar.alloc(static_cast<size_t>(n) * 2)); }

  Tens op_gn(const Tens& x, const Tens* x2, const Norm& n, bool silu) {
    const int C2 = x2 ? x2->C : 0;
    if (n.C != x.C + C2) fail(HV_ERR_INVALID, "groupnorm width %d != %d + %d", n.C, x.C, C2);
    Tens out = alloc_act(x.NF, x.H, x.W, x.C + C2);
    float* stats = static_cast<float*>(ar.alloc(sizeof(float) * groupnorm_scratch_floats(x.C + C2, x.NF, x.H * x.W, cfg.norm_groups, sms)));
    if (ar.dry) return out;
    launches += groupnorm_num_launches(x.C + C2, x.NF, x.H * x.W, sms);
    Timed tm(this, CAT_NORM, 0, "groupnorm", x.rows(), x.C + C2, 0);
    ck(launch_groupnorm(x.p, x.C, x2 ? x2->p : nullptr, C2, n.g, n.b, out.p, x.NF, x.H * x.W, cfg.norm_groups, n.eps, silu ? 1 : 0, stats, sms, st),
       "groupnorm");
    return out;
  }
  // out = LN(x (+ pre_add per batch item)) (+ pe per frame); returns the normalised tensor, x_new receives x + pre_add
  Tens op_ln(const Tens& x, const Norm& n, const __half* pre_add, int64_t rows_per_b, Tens* x_new, const __half* pe, int F) {
    Tens out = alloc_act(x.NF, x.H, x.W, x.C);
    if (pre_add && x_new) *x_new = alloc_act(x.NF, x.H, x.W, x.C);
    if (ar.dry) return out;
    launches += 1;
    Timed tm(this, CAT_NORM, 0, "layernorm", x.rows(), x.C, 0);
    ck(launch_layernorm(x.p, n.g, n.b, out.p, x.rows(), x.C, n.eps, pre_add, rows_per_b, (pre_add && x_new) ? x_new->p : nullptr, pe, x.H * x.W, F, st),
       "layernorm");
    return out;
  }
  void gemm(const __half* A, int64_t lda, const __half* A2, int64_t lda2, int64_t K1, const Mat& w, __half* out, int64_t ldc, int64_t M,
            const hv_epilogue* ep) {
    if (ar.dry) return;
    launches += 1;
    Timed tm(this, CAT_GEMM, 2.0 * M * w.rows * w.cols, (ep && ep->geglu) ? "gemm_geglu" : ((ep && ep->residual) ? "gemm_res" : "gemm"), M, w.rows, w.cols);
    ckop(op_gemm(A, lda, A2, lda2, K1, w.p, out, ldc, M, w.rows, w.cols, ep, st), "gemm");
  }
  // y = x W^T + b (+ residual)
  Tens op_linear(const Tens& x, const Lin& l, const Tens* residual, int act = HV_ACT_NONE, bool geglu = false) {
    const int N = static_cast<int>(geglu ? l.w.rows / 2 : l.w.rows);
    if (l.w.cols != x.C) fail(HV_ERR_INVALID, "linear: input width %d != weight cols %lld", x.C, (long long)l.w.cols);
    Tens out = alloc_act(x.NF, x.H, x.W, N);
    hv_epilogue ep{};
    ep.bias = l.bias;
    ep.act = act;
    ep.geglu = geglu ? 1 : 0;
    if (residual) { ep.residual = residual->p; ep.ldr = residual->C; }
    gemm(x.p, x.C, nullptr, 0, 0, l.w, out.p, N, x.rows(), &ep);
    return out;
  }
  Tens op_conv3(const Tens& x, const Conv3& c, int stride, const __half* rowvec, int64_t rows_per_group, int act, const Tens* residual) {
    if (x.C != c.cin_pad) fail(HV_ERR_INVALID, "conv3x3: input has %d channels, weight packed for %d", x.C, c.cin_pad);
    const int Ho = stride == 1 ? x.H : x.H / 2, Wo = stride == 1 ? x.W : x.W / 2;
    Tens out = alloc_act(x.NF, Ho, Wo, c.cout_pad);
    if (ar.dry) return out;
    hv_epilogue ep{};
    ep.bias = c.bias;
    ep.rowvec = rowvec;
    ep.rowvec_ld = c.cout_pad;
    ep.rows_per_group = static_cast<int32_t>(rows_per_group);
    ep.act = act;
    if (residual) { ep.residual = residual->p; ep.ldr = residual->C; }
    launches += 1;
    Timed tm(this, CAT_CONV, 2.0 * out.rows() * c.cout * 9.0 * c.cin, stride == 1 ? "conv3" : "conv3_s2", out.rows(), c.cout_pad, 9LL * c.cin_pad);
    ckop(op_conv3x3(x.p, c.w.p, out.p, c.cout_pad, x.NF, x.H, x.W, c.cin_pad, c.cout_pad, stride, &ep, st), "conv3x3");
    return out;
  }
  __half* op_small_linear(const __half* x, const Lin& l, int M, int act_in) {
    __half* out = alloc_h(static_cast<int64_t>(M) * l.w.rows);
    if (ar.dry) return out;
    launches += 1;
    Timed tm(this, CAT_MISC, 2.0 * M * l.w.rows * l.w.cols);
    ck(launch_small_linear(x, l.w.p, l.bias, out, M, static_cast<int>(l.w.rows), static_cast<int>(l.w.cols), act_in, st), "small_linear");
    return out;
  }

  // ================================================================== blocks
  // ResnetBlock3D (resnet.py:215-245); x2 = skip tensor concatenated on the channel axis (never materialised raw)
  Tens resnet_fwd(const ResnetW& r, const Tens& x, const Tens* x2, const __half* emb, int B, int F) {
    Tens out;
    {
      const int Cin = x.C + (x2 ? x2->C : 0);
      if (Cin != r.cin) fail(HV_ERR_INVALID, "resnet expects %d channels, got %d", r.cin, Cin);
      out = alloc_act(x.NF, x.H, x.W, r.cout);
    }
    Scope s(ar);
    __half* tproj = op_small_linear(emb, r.temb, B, HV_ACT_SILU);  // time_emb_proj(silu(emb)) : [B][cout]
    Tens h = op_gn(x, x2, r.n1, true);
    Tens h1 = op_conv3(h, r.c1, 1, tproj, static_cast<int64_t>(F) * x.H * x.W, HV_ACT_NONE, nullptr);
    Tens h2 = op_gn(h1, nullptr, r.n2, true);
    Tens res = x;
    if (r.has_sc) {
      res = alloc_act(x.NF, x.H, x.W, r.cout);
      hv_epilogue ep{};
      ep.bias = r.sc.bias;
      gemm(x.p, x.C, x2 ? x2->p : nullptr, x2 ? x2->C : 0, x2 ? x.C : 0, r.sc.w, res.p, r.cout, x.rows(), &ep);
    } else if (x2) {
      fail(HV_ERR_INVALID, "resnet with concatenated input must have a shortcut conv");
    }
    // conv2 + bias, + residual, written straight into `out`
    if (!ar.dry) {
      hv_epilogue ep{};
      ep.bias = r.c2.bias;
      ep.residual = res.p;
      ep.ldr = res.C;
      launches += 1;
      Timed tm(this, CAT_CONV, 2.0 * x.rows() * r.cout * 9.0 * r.cout, "conv3_res", x.rows(), r.cout, 9LL * r.c2.cin_pad);
      ckop(op_conv3x3(h2.p, r.c2.w.p, out.p, r.cout, x.NF, x.H, x.W, r.c2.cin_pad, r.c2.cout_pad, 1, &ep, st), "resnet conv2");
    }
    return out;
  }

  // Transformer3DModel + TemporalBasicTransformerBlock (+ reference-bank read hook)
  Tens spatial_fwd(const SpatialW& w, const Tens& x, const __half* ehs, int B, int F, uint32_t flags) {
    Tens out = alloc_act(x.NF, x.H, x.W, x.C);
    Scope s(ar);
    const int C = w.C, L = x.H * x.W;
    const int64_t tokens = x.rows();
    Tens hn = op_gn(x, nullptr, w.gn, false);
    Tens t = op_linear(hn, w.proj_in, nullptr);
    // ---- attn1
    Tens t1 = alloc_act(x.NF, x.H, x.W, C);  // t after self-attention residual
    {
      Scope s2(ar);
      Tens n1 = op_ln(t, w.ln1, nullptr, 1, nullptr, nullptr, 1);
      const int64_t hp = static_cast<int64_t>(w.heads) * w.dpad;
      __half* qk = alloc_h(tokens * 2 * hp);
      gemm(n1.p, C, nullptr, 0, 0, w.wqk, qk, 2 * hp, tokens, nullptr);
      const int64_t Lp = rup(L, 8), ldvt = static_cast<int64_t>(x.NF) * Lp;
      const int64_t vrows = static_cast<int64_t>(w.heads) * w.dv;
      __half* vt = alloc_h(vrows * ldvt);
      // V^T[C][frame n: n*Lp + j] = Wv [C][C] * n1^T : weights are the "A" operand, activations the batched "B" operand
      if (!ar.dry) {
        launches += 1;
        Timed tm(this, CAT_GEMM, 2.0 * tokens * C * C, "gemm_vt", vrows, tokens, C);
        ckop(op_gemm_batched_b(w.wv.p, C, n1.p, C, vt, ldvt, vrows, x.NF, L, Lp, C, w.vones, st), "V^T gemm");
      }
      // write hook (mutual_self_attention.py:137-146): the bank IS LayerNorm-1's output; attention then runs on itself only
      if (bank_out != nullptr && !ar.dry) {
        launches += 1;
        ck(cudaMemcpyAsync(bank_out[w.reader_idx], n1.p, static_cast<size_t>(tokens) * C * 2, cudaMemcpyDeviceToDevice, st), "bank write");
      }
      // HV_FLAG_CFG: batch = [uncond ; cond], the first half ignores the bank.  HV_FLAG_UNCOND_ONLY / HV_FLAG_COND_ONLY: the batch
      // holds one CFG half only (a (window x CFG-half) unit of a multi-GPU split); a cond-only batch of B items reads the LAST B
      // items of the bank (the bank is written for [uncond ; cond]).
      const bool cfg_split = (flags & HV_FLAG_CFG) != 0 && B >= 2 && B % 2 == 0 && !(flags & (HV_FLAG_UNCOND_ONLY | HV_FLAG_COND_ONLY));
      const bool use_bank = w.bank != nullptr && bank_out == nullptr && !(flags & HV_FLAG_UNCOND_ONLY);
      __half *kb = nullptr, *vbt = nullptr;
      int64_t ldvbt = 0, Lbp = 0;
      if (use_bank) {
        int64_t item0 = 0;
        if (flags & HV_FLAG_COND_ONLY) {
          if (w.bank_B < B) fail(HV_ERR_INVALID, "reference bank of '%s' has batch %lld < forward batch %d", w.name.c_str(), (long long)w.bank_B, B);
          item0 = w.bank_B - B;
        } else if (w.bank_B != B && !ar.dry) {   // (a planning pass only sizes the workspace; it does not know the flags of the forward to come)
          fail(HV_ERR_INVALID, "reference bank of '%s' has batch %lld, forward batch is %d", w.name.c_str(), (long long)w.bank_B, B);
        }
        const __half* bank = w.bank + item0 * w.bank_L * C;
        const int64_t bt = static_cast<int64_t>(B) * w.bank_L;
        kb = alloc_h(bt * hp);
        gemm(bank, C, nullptr, 0, 0, w.wk, kb, hp, bt, nullptr);
        Lbp = rup(w.bank_L, 8);
        ldvbt = static_cast<int64_t>(B) * Lbp;
        vbt = alloc_h(vrows * ldvbt);
        if (!ar.dry) {
          launches += 1;
          Timed tm(this, CAT_GEMM, 2.0 * bt * C * C);
          ckop(op_gemm_batched_b(w.wv.p, C, bank, C, vbt, ldvbt, vrows, B, w.bank_L, Lbp, C, w.vones, st), "bank V^T gemm");
        }
      }
      Tens o = alloc_act(x.NF, x.H, x.W, C);
      if (!ar.dry) {
        AttnArgs a{};
        a.q = qk; a.k = qk + hp; a.vt = vt; a.out = o.p;
        a.NF = x.NF; a.L = L; a.heads = w.heads; a.d = w.d; a.dpad = w.dpad;
        a.ldq = 2 * hp; a.ldk = 2 * hp; a.ldvt = ldvt; a.ldo = C;
        a.kb = kb; a.vbt = vbt; a.Lb = use_bank ? static_cast<int>(w.bank_L) : 0; a.ldkb = hp; a.ldvbt = ldvbt;
        a.F = F; a.nf_nobank = cfg_split ? (B / 2) * F : 0;
        a.vt_stride = Lp; a.vbt_stride = Lbp;
        launches += 1;
        const double nb_frames = x.NF - a.nf_nobank;
        Timed tm(this, CAT_ATTN, 4.0 * L * w.C * (static_cast<double>(x.NF) * L + nb_frames * a.Lb), "attn", x.NF, L, w.d);
        cudaError_t e = launch_attention(a, sms, st);
        if (e != cudaSuccess) fail(HV_ERR_CUDA, "attention (%s): %s %s", w.name.c_str(), cudaGetErrorString(e), tma_last_error());
      }
      hv_epilogue ep{};
      ep.bias = w.out1.bias;
      ep.residual = t.p;
      ep.ldr = C;
      gemm(o.p, C, nullptr, 0, 0, w.out1.w, t1.p, C, tokens, &ep);
    }
    // ---- attn2 with a single key: softmax == 1 -> out = to_out(to_v(ehs_b)) for every token of batch item b
    __half* v2 = op_small_linear(ehs, w.v2, B, HV_ACT_NONE);
    __half* ca = op_small_linear(v2, w.out2, B, HV_ACT_NONE);
    // LayerNorm-3 normalises t2 = t1 + ca[b] without writing t2: the same per-batch vector is added again, in fp32, in FF2's epilogue
    // (t3 = ff2(.) + bias + ca[b] + t1), which saves one full-tensor write per block and one rounding of the residual stream
    Tens n3 = op_ln(t1, w.ln3, ca, static_cast<int64_t>(F) * L, nullptr, nullptr, 1);
    // ---- feed-forward
    Tens ffh = op_linear(n3, w.ff1, nullptr, HV_ACT_NONE, true);
    Tens t3 = alloc_act(x.NF, x.H, x.W, C);
    {
      hv_epilogue ep{};
      ep.bias = w.ff2.bias;
      ep.rowvec = ca;
      ep.rowvec_ld = C;
      ep.rows_per_group = static_cast<int32_t>(static_cast<int64_t>(F) * L);
      ep.residual = t1.p;
      ep.ldr = C;
      gemm(ffh.p, ffh.C, nullptr, 0, 0, w.ff2.w, t3.p, C, tokens, &ep);
    }
    // ---- proj_out + residual
    hv_epilogue ep{};
    ep.bias = w.proj_out.bias;
    ep.residual = x.p;
    ep.ldr = x.C;
    gemm(t3.p, C, nullptr, 0, 0, w.proj_out.w, out.p, C, tokens, &ep);
    return out;
  }

  // one temporal attention sub-block: t = to_out(attn(LN(t) + pe)) + t
  Tens tattn_fwd(const TAttnW& a, const Tens& t, int B, int F, int heads, int d) {
    if (F > a.max_len) fail(HV_ERR_INVALID, "video_length %d exceeds temporal_position_encoding_max_len %d", F, a.max_len);
    Tens out = alloc_act(t.NF, t.H, t.W, t.C);
    Scope s(ar);
    const int C = t.C;
    Tens n = op_ln(t, a.ln, nullptr, 1, nullptr, a.pe, F);
    __half* qkv = alloc_h(t.rows() * 3 * C);
    gemm(n.p, C, nullptr, 0, 0, a.wqkv, qkv, 3 * C, t.rows(), nullptr);
    Tens o = alloc_act(t.NF, t.H, t.W, C);
    if (!ar.dry) {
      launches += 1;
      Timed tm(this, CAT_TATTN, 4.0 * t.rows() * F * C, "tattn", t.rows(), F, C);
      ck(launch_temporal_attention(qkv, o.p, B, F, t.H * t.W, heads, d, st), "temporal attention");
    }
    hv_epilogue ep{};
    ep.bias = a.out.bias;
    ep.residual = t.p;
    ep.ldr = C;
    gemm(o.p, C, nullptr, 0, 0, a.out.w, out.p, C, t.rows(), &ep);
    return out;
  }

  // VanillaTemporalModule (motion_module.py:44-259)
  Tens motion_fwd(const MotionW& m, const Tens& x, int B, int F) {
    Tens out = alloc_act(x.NF, x.H, x.W, x.C);
    Scope s(ar);
    Tens hn = op_gn(x, nullptr, m.gn, false);
    Tens t = op_linear(hn, m.proj_in, nullptr);
    for (const auto& a : m.attn) t = tattn_fwd(a, t, B, F, m.heads, m.d);
    Tens n = op_ln(t, m.ffn, nullptr, 1, nullptr, nullptr, 1);
    Tens ffh = op_linear(n, m.ff1, nullptr, HV_ACT_NONE, true);
    Tens t2 = op_linear(ffh, m.ff2, &t);
    hv_epilogue ep{};
    ep.bias = m.proj_out.bias;
    ep.residual = x.p;
    ep.ldr = x.C;
    gemm(t2.p, x.C, nullptr, 0, 0, m.proj_out.w, out.p, x.C, x.rows(), &ep);
    return out;
  }

  // ================================================================== forwards
  void begin(void* ws, size_t ws_bytes, bool dry, cudaStream_t stream) {
    st = stream;
    ar.dry = dry;
    ar.off = 0;
    ar.peak = 0;
    launches = 0;
    tap_list.clear();
    if (!dry) { recs.clear(); ev_used = 0; }
    if (dry) {
      ar.base = nullptr;
      ar.cap = 0;
    } else {
      ar.base = static_cast<uint8_t*>(ws);
      ar.cap = ws_bytes;
    }
  }

  void unet_forward(const __half* sample, int64_t timestep, const __half* ehs, const __half* pose, __half* outp, int B, int F, int H, int W,
                    uint32_t flags) {
    const int* ch = cfg.block_out_channels;
    const int NF = B * F;
    if ((H % 8) || (W % 8)) fail(HV_ERR_INVALID, "latent size %dx%d must be a multiple of 8 (three stride-2 levels)", H, W);
    // time embedding (unet_3d.py:446-467)
    __half* tsin = alloc_h(static_cast<int64_t>(B) * ch[0]);
    if (!ar.dry) { launches += 1; ck(launch_timestep_embedding(timestep, ts_table, ts_index, tsin, B, ch[0], st), "timestep embedding"); }
    __half* e1 = op_small_linear(tsin, te1, B, HV_ACT_NONE);
    __half* emb = op_small_linear(e1, te2, B, HV_ACT_SILU);
    // conv_in (+ pose_cond_fea as the epilogue residual); the 4 latent channels are zero-padded to one 64-wide k-block
    Tens x0 = alloc_act(NF, H, W, conv_in.cin_pad);
    Tens pc = pose ? alloc_act(NF, H, W, ch[0]) : Tens{};
    if (!ar.dry) {
      launches += 2 + (pose ? 1 : 0);
      ck(cudaMemsetAsync(x0.p, 0, static_cast<size_t>(x0.numel()) * 2, st), "memset");
      ck(launch_ncfhw_to_nhwc(sample, x0.p, B, cfg.in_channels, F, H, W, 0, st, x0.C), "sample layout");
      if (pose) ck(launch_ncfhw_to_nhwc(pose, pc.p, B, ch[0], F, H, W, 0, st), "pose layout");
    }
    Tens h = op_conv3(x0, conv_in, 1, nullptr, 1, HV_ACT_NONE, pose ? &pc : nullptr);
    tap("conv_in", h);
    std::vector<Tens> skips{h};
    for (size_t i = 0; i < down.size(); ++i) {
      auto& d = down[i];
      const std::string pb = "down_blocks." + std::to_string(i);
      for (size_t j = 0; j < d.res.size(); ++j) {
        h = resnet_fwd(d.res[j], h, nullptr, emb, B, F);
        tap(pb + ".resnets." + std::to_string(j), h);
        if (!d.attn.empty()) { h = spatial_fwd(d.attn[j], h, ehs, B, F, flags); tap(pb + ".attentions." + std::to_string(j), h); }
        if (!d.mm.empty()) { h = motion_fwd(d.mm[j], h, B, F); tap(pb + ".motion_modules." + std::to_string(j), h); }
        skips.push_back(h);
      }
      if (d.has_down) {
        h = op_conv3(h, d.down, 2, nullptr, 1, HV_ACT_NONE, nullptr);
        tap(pb + ".downsamplers.0", h);
        skips.push_back(h);
      }
    }
    h = resnet_fwd(mid.res[0], h, nullptr, emb, B, F);
    tap("mid_block.resnets.0", h);
    h = spatial_fwd(mid.attn[0], h, ehs, B, F, flags);
    tap("mid_block.attentions.0", h);
    if (!mid.mm.empty()) { h = motion_fwd(mid.mm[0], h, B, F); tap("mid_block.motion_modules.0", h); }
    h = resnet_fwd(mid.res[1], h, nullptr, emb, B, F);
    tap("mid_block.resnets.1", h);
    for (size_t i = 0; i < up.size(); ++i) {
      auto& u = up[i];
      const std::string pb = "up_blocks." + std::to_string(i);
      for (size_t j = 0; j < u.res.size(); ++j) {
        Tens skip = skips.back();
        skips.pop_back();
        h = resnet_fwd(u.res[j], h, &skip, emb, B, F);
        tap(pb + ".resnets." + std::to_string(j), h);
        if (!u.attn.empty()) { h = spatial_fwd(u.attn[j], h, ehs, B, F, flags); tap(pb + ".attentions." + std::to_string(j), h); }
        if (!u.mm.empty()) { h = motion_fwd(u.mm[j], h, B, F); tap(pb + ".motion_modules." + std::to_string(j), h); }
      }
      if (u.has_up) {
        // Upsample3D (resnet.py:68-71): nearest 2x + 3x3 conv as four 2x2 convs of the source -- the 4x tensor is never materialised
        Tens big = alloc_act(h.NF, 2 * h.H, 2 * h.W, u.up.cout);
        if (!ar.dry) {
          hv_epilogue ep{};
          ep.bias = u.up.bias;
          launches += 1;
          Timed tm(this, CAT_CONV, 2.0 * big.rows() * u.up.cout * 4.0 * u.up.cin, "upconv2x2", big.rows(), u.up.cout, 4LL * u.up.cin);
          ckop(op_upconv2x2(h.p, u.up.w.p, big.p, u.up.cout, h.NF, h.H, h.W, u.up.cin, u.up.cout, &ep, st), "upsampler conv");
        }
        h = big;
        tap(pb + ".upsamplers.0", h);
      }
    }
    if (cfg.kind == HV_KIND_UNET2D_REF) {   // no post-process: the forward's value is the last up block's output (B, C0, h, w)
      if (outp != nullptr && !ar.dry) { launches += 1; ck(launch_nhwc_to_ncfhw(h.p, h.C, outp, B, ch[0], F, H, W, st), "hidden layout"); }
      return;
    }
    Tens hn = op_gn(h, nullptr, norm_out, true);
    Tens y = op_conv3(hn, conv_out, 1, nullptr, 1, HV_ACT_NONE, nullptr);
    if (!ar.dry) { launches += 1; ck(launch_nhwc_to_ncfhw(y.p, y.C, outp, B, cfg.out_channels, F, H, W, st), "output layout"); }
  }

  void pose_guider_forward(const __half* cond, __half* outp, int B, int F, int H, int W) {
    const int NF = B * F;
    if ((H % 8) || (W % 8)) fail(HV_ERR_INVALID, "pose image %dx%d must be a multiple of 8", H, W);
    Tens h;
    if (pg_fast_in) {
      // conv_in (3 -> 16) straight from the (B, 3, F, H, W) image: no layout pass, 16-channel output
      h = alloc_act(NF, H, W, pg_in.cout);
      if (!ar.dry) {
        launches += 1;
        Timed tm(this, CAT_CONV, 2.0 * h.rows() * pg_in.cout * 9.0 * pg_in.cin, "pg_conv_in", h.rows(), pg_in.cout, 9LL * pg_in.cin);
        ck(launch_pg_conv_in(cond, pg_in.w, pg_in.bias, h.p, B, F, H, W, HV_ACT_SILU, sms, st), "pose conv_in");
      }
    } else {
      Tens x0 = alloc_act(NF, H, W, cfg.pg_cond_channels);
      h = alloc_act(NF, H, W, pg_convs[0].cin_pad);
      if (!ar.dry) {
        launches += 3;
        ck(launch_ncfhw_to_nhwc(cond, x0.p, B, cfg.pg_cond_channels, F, H, W, 0, st), "pose image layout");
        ck(cudaMemsetAsync(h.p, 0, static_cast<size_t>(h.numel()) * 2, st), "memset");
        ck(launch_conv3x3_direct_padded(x0.p, pg_in.w, pg_in.bias, h.p, NF, H, W, pg_in.cin, pg_in.cout, h.C, HV_ACT_SILU, sms, st), "pose conv_in");
      }
    }
    for (size_t i = 0; i < pg_convs.size(); ++i) {
      const bool last = i + 1 == pg_convs.size();
      const Conv3& c = pg_convs[i];
      if (pg_small[i]) {
        // the next layer either is a small-channel layer too (dense channels) or the first implicit-GEMM layer (its 64-multiple k-blocks)
        const int ld = pg_small[i + 1] ? c.cout : pg_convs[i + 1].cin_pad;
        const int Ho = pg_strides[i] == 1 ? h.H : h.H / 2, Wo = pg_strides[i] == 1 ? h.W : h.W / 2;
        if (h.C != c.cin) fail(HV_ERR_INVALID, "pose guider small conv %zu: input has %d channels, expects %d", i, h.C, c.cin);
        Tens out = alloc_act(NF, Ho, Wo, ld);
        if (!ar.dry) {
          if (ld > c.cout) { launches += 1; ck(cudaMemsetAsync(out.p, 0, static_cast<size_t>(out.numel()) * 2, st), "memset"); }
          launches += 1;
          Timed tm(this, CAT_CONV, 2.0 * out.rows() * c.cout * 9.0 * c.cin, pg_strides[i] == 1 ? "smallconv" : "smallconv_s2", out.rows(), c.cout, 9LL * c.cin);
          ck(launch_smallconv(h.p, c.w.p, c.bias, out.p, NF, h.H, h.W, c.cin, c.cout, pg_strides[i], ld, HV_ACT_SILU, st), "pose small conv");
        }
        h = out;
        continue;
      }
      h = op_conv3(h, c, pg_strides[i], nullptr, 1, last ? HV_ACT_NONE : HV_ACT_SILU, nullptr);
      // the next conv expects cin_pad channels: cout_pad of this conv equals it by construction (both round to 64)
      if (!last && h.C != pg_convs[i + 1].cin_pad) fail(HV_ERR_INVALID, "pose guider channel padding mismatch at conv %zu", i);
    }
    if (!ar.dry) { launches += 1; ck(launch_nhwc_to_ncfhw(h.p, h.C, outp, B, cfg.pg_out_channels, F, h.H, h.W, st), "pose output layout"); }
  }

  void camera_forward(const __half* plucker, __half* outp, int B, int F, int H, int W, const float* rays_K = nullptr, const float* rays_c2w = nullptr) {
    const int r = cfg.cam_downscale, NF = B * F, C = cfg.cam_channels;
    if ((H % r) || (W % r)) fail(HV_ERR_INVALID, "plucker map %dx%d must be a multiple of %d", H, W, r);
    const int cin0 = cfg.cam_cin / (r * r);
    Tens u = alloc_act(NF, H / r, W / r, cfg.cam_cin);
    if (!ar.dry) {
      launches += 1;
      if (rays_K != nullptr) {   // Plucker embedding generated in place of the unshuffle's gather
        if (cin0 != 6) fail(HV_ERR_INVALID, "the Plucker producer makes 6 channels, encoder expects %d", cin0);
        ck(launch_plucker_unshuffle(rays_K, rays_c2w, u.p, NF, H, W, r, sms, st), "plucker unshuffle");
      } else {
        ck(launch_pixel_unshuffle(plucker, u.p, B, cin0, F, H, W, r, sms, st), "pixel unshuffle");
      }
    }
    Tens x = op_conv3(u, cam_in, 1, nullptr, 1, HV_ACT_NONE, nullptr);
    for (int j = 0; j < cfg.cam_nums_rb; ++j) {
      // ResnetBlock (pose_adaptor.py:102-135, sk=True, ksize=1): x = block2(relu(block1(x))) + x
      Tens h1 = op_conv3(x, cam_res[j].block1, 1, nullptr, 1, HV_ACT_RELU, nullptr);
      x = op_linear(h1, cam_res[j].block2, &x);
      // TemporalTransformerBlock over the frame axis
      const MotionW& m = cam_att[j];
      Tens t = tattn_fwd(m.attn[0], x, B, F, m.heads, m.d);
      Tens n = op_ln(t, m.ffn, nullptr, 1, nullptr, nullptr, 1);
      Tens ffh = op_linear(n, m.ff1, nullptr, HV_ACT_NONE, true);
      x = op_linear(ffh, m.ff2, &t);
    }
    Tens z = op_linear(x, cam_zero, nullptr);
    if (!ar.dry) { launches += 1; ck(launch_nhwc_to_ncfhw(z.p, z.C, outp, NF, C, 1, z.H, z.W, st), "camera output layout"); }
  }
};

// ------------------------------------------------------------------------------------------ C ABI
#define HV_GUARD(h, ...)                                  \
  try {                                                   \
    __VA_ARGS__;                                          \
    return HV_OK;                                         \
  } catch (const Err& e) {                                \
    if (h) (h)->err = e.msg;                              \
    set_error("%s", e.msg.c_str());                       \
    return e.code;                                        \
  } catch (const std::exception& e) {                     \
    if (h) (h)->err = e.what();                           \
    set_error("%s", e.what());                            \
    return HV_ERR_INVALID;                                \
  }

// The forwards never allocate: a caller that passes no workspace must have called hv_reserve_workspace for a shape at least as
// large (cudaMalloc synchronises the device and is illegal under stream capture).
static void* get_ws(hv_model* m, void* ws, size_t ws_bytes, size_t need, size_t* have) {
  if (ws) {
    *have = ws_bytes;
    return ws;
  }
  if (m->private_ws_bytes < need)
    fail(HV_ERR_STATE, "workspace == NULL and the handle's reserved workspace (%zu bytes) is smaller than the %zu this shape needs: call hv_reserve_workspace first",
         m->private_ws_bytes, need);
  *have = m->private_ws_bytes;
  return m->private_ws;
}

extern "C" {

int hv_create(const hv_config* cfg, hv_handle* out) {
  if (!cfg || !out) return HV_ERR_INVALID;
  hv_model* m = nullptr;
  try {
    const int sms = device_sms();
    if (!sms) fail(HV_ERR_CUDA, "%s", last_error());
    m = new hv_model();
    m->cfg = *cfg;
    m->sms = sms;
    if (cfg->kind < 0 || cfg->kind > 3) fail(HV_ERR_INVALID, "unknown kind %d", cfg->kind);
    if (cfg->kind == HV_KIND_UNET2D_REF && cfg->use_motion_module) fail(HV_ERR_INVALID, "the reference (writer) UNet has no motion modules");
    *out = m;
    return HV_OK;
  } catch (const Err& e) {
    set_error("%s", e.msg.c_str());
    delete m;
    return e.code;
  }
}

void hv_destroy(hv_handle h) { delete h; }
const char* hv_last_error(hv_handle h) { return h ? h->err.c_str() : last_error(); }
int64_t hv_last_launch_count(hv_handle h) { return h ? h->launches : 0; }

int hv_set_profiling(hv_handle h, int32_t enable) {
  if (!h) return HV_ERR_INVALID;
  h->profiling = enable != 0;
  return HV_OK;
}

int hv_dump_profile(hv_handle h, const char* path) {
  if (!h || !path) return HV_ERR_INVALID;
  HV_GUARD(h, {
    FILE* f = fopen(path, "w");
    if (!f) fail(HV_ERR_INVALID, "cannot open %s", path);
    if (!h->recs.empty()) ck(cudaEventSynchronize(h->recs.back().b), "profile sync");
    fprintf(f, "idx,cat,label,M,N,K,ms,tflops\n");
    int i = 0;
    for (const auto& r : h->recs) {
      float t = 0;
      cudaEventElapsedTime(&t, r.a, r.b);
      fprintf(f, "%d,%d,%s,%lld,%lld,%lld,%.4f,%.1f\n", i++, r.cat, r.label, r.m, r.n, r.k, t, t > 0 ? r.flops / 1e9 / t : 0.0);
    }
    fclose(f);
  });
}

int hv_get_profile(hv_handle h, double* ms, double* flops, int64_t* count, int32_t ncat) {
  if (!h || !ms || !flops || !count) return HV_ERR_INVALID;
  HV_GUARD(h, {
    for (int i = 0; i < ncat; ++i) { ms[i] = 0; flops[i] = 0; count[i] = 0; }
    if (!h->recs.empty()) ck(cudaEventSynchronize(h->recs.back().b), "profile sync");
    for (const auto& r : h->recs) {
      if (r.cat >= ncat) continue;
      float t = 0;
      ck(cudaEventElapsedTime(&t, r.a, r.b), "cudaEventElapsedTime");
      ms[r.cat] += t;
      flops[r.cat] += r.flops;
      count[r.cat] += 1;
    }
  });
}

__global__ void cast_f32_to_f16(const float* __restrict__ x, __half* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = __float2half_rn(x[i]);
}

int hv_set_weight(hv_handle h, const char* key, const void* dev_ptr, const int64_t* shape, int32_t ndim, int32_t dtype, hv_stream_t stream) {
  if (!h || !key || !dev_ptr || ndim < 0 || ndim > 8) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (h->finalized) fail(HV_ERR_STATE, "hv_set_weight after hv_finalize");
    Raw r;
    r.shape.assign(shape, shape + ndim);
    const int64_t n = r.numel();
    if (n <= 0) fail(HV_ERR_INVALID, "weight '%s' is empty", key);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* p = nullptr;
    ck(cudaMalloc(&p, static_cast<size_t>(rup(n, 8)) * 2), "cudaMalloc(weight)");
    h->owned.push_back(p);
    r.p = static_cast<__half*>(p);
    if (dtype == HV_F16) {
      ck(cudaMemcpyAsync(r.p, dev_ptr, static_cast<size_t>(n) * 2, cudaMemcpyDeviceToDevice, st), "weight copy");
    } else if (dtype == HV_F32) {
      long long blocks = std::min<long long>((n + 255) / 256, 4096);
      cast_f32_to_f16<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const float*>(dev_ptr), r.p, n);
      ck(cudaGetLastError(), "weight cast");
    } else {
      fail(HV_ERR_INVALID, "weight '%s': unsupported dtype %d", key, dtype);
    }
    h->raw[key] = r;
  });
}

int hv_finalize(hv_handle h, hv_stream_t stream) {
  if (!h) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (h->finalized) return HV_OK;
    h->st = static_cast<cudaStream_t>(stream);
    if (h->cfg.kind == HV_KIND_UNET3D || h->cfg.kind == HV_KIND_UNET2D_REF) h->build_unet();
    else if (h->cfg.kind == HV_KIND_POSE_GUIDER) h->build_pose_guider();
    else h->build_camera();
    h->finalized = true;
  });
}

int hv_num_ref_blocks(hv_handle h) { return (h && h->finalized) ? static_cast<int>(h->readers.size()) : 0; }
int hv_ref_block_dim(hv_handle h, int32_t i) {
  return (h && h->finalized && i >= 0 && i < static_cast<int>(h->readers.size())) ? h->readers[i]->C : 0;
}

int hv_set_ref_bank(hv_handle h, int32_t idx, const void* dev_ptr, int64_t B_ref, int64_t L, int64_t C, hv_stream_t stream) {
  if (!h) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized) fail(HV_ERR_STATE, "hv_set_ref_bank before hv_finalize");
    if (idx < 0 || idx >= static_cast<int>(h->readers.size())) fail(HV_ERR_INVALID, "bank index %d out of range", idx);
    SpatialW* w = h->readers[idx];
    if (C != w->C) fail(HV_ERR_INVALID, "bank %d has width %lld, block '%s' has %d", idx, (long long)C, w->name.c_str(), w->C);
    const int64_t n = B_ref * L * C;
    if (w->bank_cap < n) {
      w->bank_store = h->dmalloc(n);
      w->bank_cap = n;
    }
    w->bank = w->bank_store;
    w->bank_B = B_ref;
    w->bank_L = L;
    ck(cudaMemcpyAsync(w->bank, dev_ptr, static_cast<size_t>(n) * 2, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)), "bank copy");
  });
}

int hv_clear_ref_banks(hv_handle h) {
  if (!h) return HV_ERR_INVALID;
  for (SpatialW* w : h->readers) {
    w->bank = nullptr;  // bank_store is kept for the next hv_set_ref_bank
    w->bank_B = w->bank_L = 0;
  }
  return HV_OK;
}

static size_t measure(hv_model* h, int B, int F, int H, int W, uint32_t flags = HV_FLAG_CFG) {
  h->begin(nullptr, 0, true, nullptr);
  if (h->cfg.kind == HV_KIND_UNET3D || h->cfg.kind == HV_KIND_UNET2D_REF) {
    static __half dummy;
    h->unet_forward(&dummy, 0, &dummy, h->cfg.kind == HV_KIND_UNET3D ? &dummy : nullptr, &dummy, B, F, H, W, flags);
  } else if (h->cfg.kind == HV_KIND_POSE_GUIDER) {
    h->pose_guider_forward(nullptr, nullptr, B, F, H, W);
  } else {
    h->camera_forward(nullptr, nullptr, B, F, H, W);
  }
  return h->ar.peak + 4096;
}

int hv_reserve_workspace(hv_handle h, int32_t B, int32_t F, int32_t height, int32_t width) {
  if (!h) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized) fail(HV_ERR_STATE, "hv_reserve_workspace before hv_finalize");
    const size_t need = measure(h, B, F, height, width);
    if (h->private_ws_bytes < need) {
      if (h->private_ws) ck(cudaFree(h->private_ws), "cudaFree(workspace)");
      h->private_ws = nullptr;
      h->private_ws_bytes = 0;
      ck(cudaMalloc(&h->private_ws, need), "cudaMalloc(workspace)");
      h->private_ws_bytes = need;
    }
  });
}

int hv_set_timestep_source(hv_handle h, const int64_t* dev_table, const int32_t* dev_index) {
  if (!h || ((dev_table == nullptr) != (dev_index == nullptr))) return HV_ERR_INVALID;
  h->ts_table = reinterpret_cast<const long long*>(dev_table);
  h->ts_index = dev_index;
  return HV_OK;
}

int hv_debug_tap_count(hv_handle h, int32_t B, int32_t F, int32_t height, int32_t width) {
  if (!h || !h->finalized) return HV_ERR_INVALID;
  try {
    measure(h, B, F, height, width);
    return static_cast<int>(h->tap_list.size());
  } catch (const Err& e) {
    h->err = e.msg;
    return e.code;
  }
}

int hv_debug_tap_info(hv_handle h, int32_t i, char* name, int32_t name_cap, int64_t* dims4) {
  if (!h || !name || !dims4 || i < 0 || i >= static_cast<int>(h->tap_list.size()) || name_cap <= 0) return HV_ERR_INVALID;
  const auto& t = h->tap_list[i];
  snprintf(name, static_cast<size_t>(name_cap), "%s", t.name.c_str());
  dims4[0] = t.NF; dims4[1] = t.H; dims4[2] = t.W; dims4[3] = t.C;
  return HV_OK;
}

int hv_debug_set_taps(hv_handle h, void* const* dst, int32_t n) {
  if (!h || n < 0 || (n > 0 && !dst)) return HV_ERR_INVALID;
  h->tap_dst.assign(dst, dst + n);
  return HV_OK;
}

size_t hv_workspace_bytes(hv_handle h, int32_t B, int32_t F, int32_t height, int32_t width) {
  if (!h || !h->finalized) return 0;
  try {
    return measure(h, B, F, height, width);
  } catch (const Err& e) {
    h->err = e.msg;
    return 0;
  }
}

int hv_unet3d_forward(hv_handle h, const void* sample, int64_t timestep, const void* ehs, const void* pose, void* out, int32_t B, int32_t F,
                      int32_t height, int32_t width, uint32_t flags, void* workspace, size_t ws_bytes, hv_stream_t stream) {
  if (!h || !sample || !ehs || !out) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized || h->cfg.kind != HV_KIND_UNET3D) fail(HV_ERR_STATE, "handle is not a finalized UNet3D");
    if ((flags & HV_FLAG_UNCOND_ONLY) && (flags & HV_FLAG_COND_ONLY)) fail(HV_ERR_INVALID, "HV_FLAG_UNCOND_ONLY and HV_FLAG_COND_ONLY are exclusive");
    const size_t need = measure(h, B, F, height, width, flags);
    size_t have = 0;
    void* ws = get_ws(h, workspace, ws_bytes, need, &have);
    h->begin(ws, have, false, static_cast<cudaStream_t>(stream));
    h->unet_forward(static_cast<const __half*>(sample), timestep, static_cast<const __half*>(ehs), static_cast<const __half*>(pose),
                    static_cast<__half*>(out), B, F, height, width, flags);
  });
}

int hv_unet2d_reference_forward(hv_handle h, const void* sample, int64_t timestep, const void* ehs, void* hidden_out, void* const* bank_out,
                                int32_t n_banks, int32_t B, int32_t height, int32_t width, void* workspace, size_t ws_bytes, hv_stream_t stream) {
  if (!h || !sample || !ehs || !bank_out) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized || h->cfg.kind != HV_KIND_UNET2D_REF) fail(HV_ERR_STATE, "handle is not a finalized reference (writer) UNet2D");
    if (n_banks != static_cast<int>(h->readers.size())) fail(HV_ERR_INVALID, "expected %d bank destinations, got %d", (int)h->readers.size(), n_banks);
    for (int i = 0; i < n_banks; ++i)
      if (bank_out[i] == nullptr) fail(HV_ERR_INVALID, "bank destination %d is NULL", i);
    const size_t need = measure(h, B, 1, height, width);
    size_t have = 0;
    void* ws = get_ws(h, workspace, ws_bytes, need, &have);
    h->begin(ws, have, false, static_cast<cudaStream_t>(stream));
    h->bank_out = reinterpret_cast<__half* const*>(bank_out);
    try {
      h->unet_forward(static_cast<const __half*>(sample), timestep, static_cast<const __half*>(ehs), nullptr, static_cast<__half*>(hidden_out), B, 1,
                      height, width, 0);
    } catch (...) {
      h->bank_out = nullptr;
      throw;
    }
    h->bank_out = nullptr;
  });
}

int hv_pose_guider_forward(hv_handle h, const void* cond, void* out, int32_t B, int32_t F, int32_t H, int32_t W, void* workspace,
                           size_t ws_bytes, hv_stream_t stream) {
  if (!h || !cond || !out) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized || h->cfg.kind != HV_KIND_POSE_GUIDER) fail(HV_ERR_STATE, "handle is not a finalized PoseGuider");
    const size_t need = measure(h, B, F, H, W);
    size_t have = 0;
    void* ws = get_ws(h, workspace, ws_bytes, need, &have);
    h->begin(ws, have, false, static_cast<cudaStream_t>(stream));
    h->pose_guider_forward(static_cast<const __half*>(cond), static_cast<__half*>(out), B, F, H, W);
  });
}

int hv_camera_encoder_forward(hv_handle h, const void* plucker, void* out, int32_t B, int32_t F, int32_t H, int32_t W, void* workspace,
                              size_t ws_bytes, hv_stream_t stream) {
  if (!h || !plucker || !out) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized || h->cfg.kind != HV_KIND_CAMERA_ENCODER) fail(HV_ERR_STATE, "handle is not a finalized CameraPoseEncoder");
    const size_t need = measure(h, B, F, H, W);
    size_t have = 0;
    void* ws = get_ws(h, workspace, ws_bytes, need, &have);
    h->begin(ws, have, false, static_cast<cudaStream_t>(stream));
    h->camera_forward(static_cast<const __half*>(plucker), static_cast<__half*>(out), B, F, H, W);
  });
}

int hv_camera_encoder_forward_rays(hv_handle h, const float* intrinsics, const float* c2w, void* out, int32_t B, int32_t F, int32_t H, int32_t W,
                                   void* workspace, size_t ws_bytes, hv_stream_t stream) {
  if (!h || !intrinsics || !c2w || !out) return HV_ERR_INVALID;
  HV_GUARD(h, {
    if (!h->finalized || h->cfg.kind != HV_KIND_CAMERA_ENCODER) fail(HV_ERR_STATE, "handle is not a finalized CameraPoseEncoder");
    const size_t need = measure(h, B, F, H, W);
    size_t have = 0;
    void* ws = get_ws(h, workspace, ws_bytes, need, &have);
    h->begin(ws, have, false, static_cast<cudaStream_t>(stream));
    h->camera_forward(nullptr, static_cast<__half*>(out), B, F, H, W, intrinsics, c2w);
  });
}

}  // extern "C"
