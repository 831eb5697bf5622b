// C-ABI (include/i2v_b200.h) + host-side orchestration of the four networks of the sampling path.
// The host code only sequences kernel launches on the caller's stream inside the caller's workspace;
// it never allocates device memory and never synchronises.
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/i2v_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace i2v {

static thread_local char g_err[1024] = "";

// ------------------------------------------------------------------ launch accounting / timing
namespace {
struct ProfRec { cudaEvent_t a, b; int cat; double flops, bytes; };
struct ProfState {
    bool on = false;
    long long launches[PROF_NCAT] = {0, 0, 0, 0, 0, 0, 0};
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    size_t pool_used = 0;
    cudaEvent_t get() {
        if (pool_used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[pool_used++];
    }
};
// per host thread: handles are "one per host thread / stream" (include/i2v_b200.h), and events belong to the device
// that was current when they were created -- a thread profiles the one device it drives
thread_local ProfState g_prof;
std::string g_prof_dump_path;
TuneOptions g_tune;
}  // namespace

TuneOptions& tune() { return g_tune; }

ProfScope::ProfScope(int cat, double flops, double bytes, cudaStream_t stream) : idx_(-1), stream_(stream) {
    g_prof.launches[cat]++;
    if (!g_prof.on) return;
    ProfRec r{g_prof.get(), g_prof.get(), cat, flops, bytes};
    cudaEventRecord(r.a, stream);
    idx_ = (int)g_prof.recs.size();
    g_prof.recs.push_back(r);
}
ProfScope::~ProfScope() {
    if (idx_ >= 0) cudaEventRecord(g_prof.recs[idx_].b, stream_);
}

bool pdl_enabled() { return g_tune.pdl != 0; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

struct TensorTable {
    std::unordered_map<std::string, std::pair<const void*, size_t>> t;
    int set(const char* name, const void* p, size_t nbytes) {
        I2V_REQUIRE(name != nullptr && p != nullptr, "set_tensor: null name or pointer");
        I2V_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "set_tensor(%s): pointer not 16-byte aligned", name);
        t[name] = {p, nbytes};
        return 0;
    }
    bool has(const std::string& name) const { return t.count(name) != 0; }
    // returns nullptr (and sets the error) when missing or of the wrong size
    template <class T = float>
    const T* get(const std::string& name, size_t elems) const {
        auto it = t.find(name);
        if (it == t.end()) { set_error("tensor '%s' was not registered", name.c_str()); return nullptr; }
        if (it->second.second != elems * sizeof(T)) {
            set_error("tensor '%s': %zu bytes registered, %zu expected", name.c_str(), it->second.second, elems * sizeof(T));
            return nullptr;
        }
        return static_cast<const T*>(it->second.first);
    }
};

// Bump allocator over the caller's workspace; in dry mode it only measures.
struct Arena {
    char* base; size_t cap; size_t off = 0; size_t peak = 0; bool dry;
    Arena(void* b, size_t c, bool d) : base(static_cast<char*>(b)), cap(c), dry(d) {}
    template <class T> T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = dry ? reinterpret_cast<T*>(static_cast<uintptr_t>(256)) : reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        if (off > peak) peak = off;
        return p;
    }
    bool ok() const { return dry || off <= cap; }
};

#define I2V_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
#define I2V_PTR(var, expr) auto var = (expr); if (!dry && var == nullptr) return -3

// split-K scratch of the SIMT engine (encoders' deep-K / small-M layers)
struct SplitK { float* ws = nullptr; size_t bytes = 0; unsigned* counters = nullptr; int max_tiles = 0; };
constexpr size_t kSplitKBytes = 48ull << 20;
constexpr int kSplitKTiles = 4096;

// conv helper (stride-1 'same' or general), channels-last
int conv(int engine, const float* x, const float* w, const float* bias, const float* res, float* y, int B, int Ti, int Hi,
         int Wi, int Cin, int Cout, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, int rut, int ruh,
         int ruw, int act, int out_mode, cudaStream_t s, const SplitK* sk = nullptr) {
    ConvArgs a;
    if (sk) { a.splitk_ws = sk->ws; a.splitk_ws_bytes = sk->bytes; a.splitk_counters = sk->counters; a.splitk_max_tiles = sk->max_tiles; }
    a.x = x; a.w = w; a.bias = bias; a.res = res; a.y = y;
    a.B = B; a.Ti = Ti; a.Hi = Hi; a.Wi = Wi; a.Cin = Cin;
    a.To = (Ti + 2 * pt - kt) / st + 1; a.Ho = (Hi + 2 * ph - kh) / sh + 1; a.Wo = (Wi + 2 * pw - kw) / sw + 1;
    a.Cout = Cout; a.kt = kt; a.kh = kh; a.kw = kw; a.st = st; a.sh = sh; a.sw = sw; a.pt = pt; a.ph = ph; a.pw = pw;
    a.res_ut = rut; a.res_uh = ruh; a.res_uw = ruw; a.act = act; a.out_mode = out_mode;
    (void)engine;
    return launch_conv_simt(a, s);
}

int modulate(const float* x, const float* coef, const float* gb, const float* r, const float* coef2, float* out, int B,
             int T, int H, int W, int C, int ut, int uh, int uw, int act, cudaStream_t s) {
    ModArgs m;
    m.x = x; m.coef = coef; m.gb = gb; m.r = r; m.coef2 = coef2; m.out = out;
    m.B = B; m.T = T; m.H = H; m.W = W; m.C = C; m.ut = ut; m.uh = uh; m.uw = uw; m.act = act;
    return launch_modulate(m, s);
}

}  // namespace
}  // namespace i2v

using namespace i2v;

// =============================================================================== flow
struct i2v_flow {
    int n_flows, d, zc, hidden, depth;
    std::vector<unsigned char> cond_mode;
    TensorTable tt;
};

static int flow_weights(const i2v_flow* h, FlowWeights& fw) {
    const size_t nf = h->n_flows, H = h->hidden, half = h->d / 2;
    fw.n_flows = h->n_flows; fw.d = h->d; fw.half = h->d / 2; fw.zc = h->zc; fw.hidden = h->hidden; fw.depth = h->depth;
    fw.cond_mode = h->cond_mode.data();
    const bool dry = false;
    I2V_PTR(w1x, h->tt.get("w1x", nf * 2 * 2 * H * half)); fw.w1x = w1x;
    I2V_PTR(w1c, h->tt.get("w1c", nf * 2 * 2 * H * h->zc)); fw.w1c = w1c;
    I2V_PTR(b1, h->tt.get("b1", nf * 2 * 2 * H)); fw.b1 = b1;
    I2V_PTR(wh, h->tt.get("wh", nf * 2 * h->depth * 2 * H * H)); fw.wh = wh;
    I2V_PTR(bh, h->tt.get("bh", nf * 2 * h->depth * 2 * H)); fw.bh = bh;
    I2V_PTR(wo, h->tt.get("wo", nf * 2 * 2 * half * H)); fw.wo = wo;
    I2V_PTR(bo, h->tt.get("bo", nf * 2 * 2 * half)); fw.bo = bo;
    I2V_PTR(loc, h->tt.get("loc", nf * h->d)); fw.loc = loc;
    I2V_PTR(scale, h->tt.get("scale", nf * h->d)); fw.scale = scale;
    I2V_PTR(pf, h->tt.get<int>("perm_fwd", nf * h->d)); fw.perm_fwd = pf;
    I2V_PTR(pb, h->tt.get<int>("perm_bwd", nf * h->d)); fw.perm_bwd = pb;
    if (h->tt.has("wpack") && h->d == 64 && H % 8 == 0) {
        // per coupling and CTA rank: 1 + depth * H/32 + 1 chunks of 32 * H/8 floats
        const size_t chunks = 1 + (size_t)h->depth * (H / 32) + 1;
        I2V_PTR(wp, h->tt.get("wpack", nf * 2 * 16 * chunks * 32 * (H / 8))); fw.wpack = wp;
    }
    return 0;
}

extern "C" {

int i2v_abi_version(void) { return I2V_ABI_VERSION; }
const char* i2v_last_error(void) { return i2v::g_err; }

long long i2v_launch_count(void) {
    long long n = 0;
    for (int c = 0; c < PROF_NCAT; ++c) n += g_prof.launches[c];
    return n;
}
int i2v_prof_is_enabled(void) { return g_prof.on ? 1 : 0; }
void i2v_prof_enable(int on) {
    g_prof.on = on != 0;
    g_prof.recs.clear();
    g_prof.pool_used = 0;
}
int i2v_prof_collect(double* ms, double* flops, double* bytes, long long* launches) {
    for (int c = 0; c < PROF_NCAT; ++c) { ms[c] = 0; flops[c] = 0; bytes[c] = 0; launches[c] = 0; }
    FILE* dump = g_prof_dump_path.empty() ? nullptr : fopen(g_prof_dump_path.c_str(), "w");
    if (dump) fprintf(dump, "idx,family,ms,gflop,mbytes\n");
    int idx = 0;
    for (auto& r : g_prof.recs) {
        I2V_CHECK_CUDA(cudaEventSynchronize(r.b));
        float t = 0.f;
        I2V_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        if (dump) fprintf(dump, "%d,%d,%.6f,%.6f,%.6f\n", idx++, r.cat, t, r.flops * 1e-9, r.bytes * 1e-6);
        ms[r.cat] += t; flops[r.cat] += r.flops; bytes[r.cat] += r.bytes; launches[r.cat]++;
    }
    if (dump) fclose(dump);
    g_prof.recs.clear();
    g_prof.pool_used = 0;
    return 0;
}
void i2v_prof_dump_path(const char* path) { g_prof_dump_path = path ? path : ""; }

int i2v_set_option(const char* name, double value) {
    I2V_REQUIRE(name != nullptr, "set_option: null name");
    const std::string k(name);
    const int v = (int)value;
    TuneOptions& t = tune();
    if (k == "pdl") t.pdl = v != 0;
    else if (k == "tc_flags") t.tc_flags = v;
    else if (k == "tc_persist") t.tc_persist = v != 0;
    else if (k == "tc_min_stages") t.tc_min_stages = v < 2 ? 2 : (v > 6 ? 6 : v);
    else if (k == "tc_pair") t.tc_pair = v != 0;
    else if (k == "tc_pair_stack") t.tc_pair_stack = v < 0 ? 0 : (v > 2 ? 2 : v);
    else if (k == "tc_pair_stages") t.tc_pair_stages = v < 0 ? 0 : v;
    else if (k == "linear_bfly") t.linear_bfly = v != 0;
    else if (k == "flow_cluster") t.flow_cluster = v != 0;
    else if (k == "mod_spade") t.mod_spade = v != 0;
    else if (k == "linear_k64") t.linear_k64 = v != 0;
    else if (k == "tc_t2_split") t.tc_t2_split = v != 0;
    else I2V_REQUIRE(false, "set_option: unknown option '%s'", name);
    return 0;
}

i2v_flow* i2v_flow_create(int n_flows, int d, int zc, int hidden, int depth, const unsigned char* cond_mode) {
    if (n_flows <= 0 || n_flows > 64 || d <= 0 || d % 2 || zc <= 0 || zc % 4 || hidden <= 0 || hidden % 4 || hidden > 512 ||
        depth < 0) {
        set_error("flow_create: unsupported geometry n_flows=%d d=%d zc=%d hidden=%d depth=%d", n_flows, d, zc, hidden, depth);
        return nullptr;
    }
    auto* h = new i2v_flow();
    h->n_flows = n_flows; h->d = d; h->zc = zc; h->hidden = hidden; h->depth = depth;
    h->cond_mode.assign(n_flows, 0);
    if (cond_mode) h->cond_mode.assign(cond_mode, cond_mode + n_flows);
    return h;
}
int i2v_flow_set_tensor(i2v_flow* h, const char* name, const void* p, size_t n) { return h ? h->tt.set(name, p, n) : -1; }
size_t i2v_flow_workspace_bytes(const i2v_flow* h, int batch) {
    FlowWeights fw{};
    fw.n_flows = h->n_flows; fw.d = h->d; fw.half = h->d / 2; fw.zc = h->zc; fw.hidden = h->hidden; fw.depth = h->depth;
    return flow_workspace_bytes(fw, batch);
}
int i2v_flow_reverse(i2v_flow* h, const float* residual, const float* cond, float* z, int batch, void* ws, size_t wsb,
                     void* stream) {
    FlowWeights fw{};
    I2V_TRY(flow_weights(h, fw));
    return launch_flow(fw, residual, cond, z, nullptr, batch, true, ws, wsb, static_cast<cudaStream_t>(stream));
}
int i2v_flow_forward(i2v_flow* h, const float* z, const float* cond, float* out, float* logdet, int batch, void* ws,
                     size_t wsb, void* stream) {
    FlowWeights fw{};
    I2V_TRY(flow_weights(h, fw));
    return launch_flow(fw, z, cond, out, logdet, batch, false, ws, wsb, static_cast<cudaStream_t>(stream));
}
void i2v_flow_destroy(i2v_flow* h) { delete h; }

}  // extern "C"

// =============================================================================== embedder
// Activations entering a tensor-core conv are split as fp16(ACT_SPLIT_SCALE * x): they are normalised /
// modulated values of O(1); 16 keeps |x| < 4094 representable and pushes the fp16 subnormal floor of
// the low word to ~4e-9.  loader.py folds 1/(ACT_SPLIT_SCALE * s_w) into each layer's "<conv>.ws".
static constexpr float ACT_SPLIT_SCALE = 16.f;
struct i2v_embedder {
    int zc, norm_mode;
    int tc_mode = 1;      // 0: fp32 SIMT convs only  1: tensor-core convs where the GEMM fills the machine  2: wherever supported
    int tc_min_ctas = 8;  // mode 1: fewest 128-row x 128-column tiles for which the tensor-core engine is used
    TensorTable tt;
};

// ResNet-50 v1.5 trunk on channels-last tensors (AE.py:131-141).  norm_mode 0: every conv is followed
// by InstanceNorm2d(affine=False) -> statistics pass + fused normalise/ReLU(/residual) pass;
// norm_mode 1: BatchNorm (eval) was folded into conv weight+bias at load, ReLU/residual ride in the
// conv epilogue.
//
// InstanceNorm variant, tc_mode >= 1: the stride-1 convs whose GEMM fills the machine (layer1/layer2 at 64x64 input,
// one more stage at 128x128) run on the tensor-core engine (error-compensated fp16 split, fp32-grade, conv_tc.cu).
// The normalise/ReLU pass that feeds such a conv writes the fp16 split directly, and the block's final pass
// writes both the fp32 tensor (the next identity branch) and its split.  Small-M layers (layer3/4: a handful of
// 128-row tiles) stay on the fp32 SIMT engine with its split-K.
static int embedder_run(const i2v_embedder* m, const float* x0, float* embed, int B, int H, int W, Arena& ar,
                        cudaStream_t s, bool dry) {
    const bool inorm = m->norm_mode == 0;
    const int H1 = (H + 6 - 7) / 2 + 1, W1 = (W + 6 - 7) / 2 + 1;          // conv1 7x7 s2 p3
    const int H2 = (H1 + 2 - 3) / 2 + 1, W2 = (W1 + 2 - 3) / 2 + 1;        // maxpool 3x3 s2 p1
    const size_t act_max = (size_t)B * H1 * W1 * 64;                       // == B*H2*W2*256, the largest activation
    float* xin = ar.take<float>((size_t)B * H * W * 3);
    float* buf[5];
    for (auto& b : buf) b = ar.take<float>(act_max);
    // fp16 (hi | lo) copies of conv inputs for the tensor-core engine: same bytes as the fp32 tensor
    float* sbuf[3] = {nullptr, nullptr, nullptr};
    const bool any_tc = m->tc_mode != 0;
    if (any_tc) for (auto& b : sbuf) b = ar.take<float>(act_max);
    double* sums = ar.take<double>((size_t)B * 2048 * 2);
    double* sums2 = ar.take<double>((size_t)B * 2048 * 2);
    float* coef = ar.take<float>((size_t)B * 2048 * 2);
    float* coef2 = ar.take<float>((size_t)B * 2048 * 2);
    float* pooled = ar.take<float>((size_t)B * 2048);
    SplitK sk;
    sk.ws = ar.take<float>(kSplitKBytes / sizeof(float)); sk.bytes = kSplitKBytes;
    sk.counters = ar.take<unsigned>(kSplitKTiles); sk.max_tiles = kSplitKTiles;
    if (!ar.ok()) { set_error("embedder: workspace too small (%zu needed, %zu given)", ar.peak, ar.cap); return -4; }
    if (dry) return 0;
    I2V_CHECK_CUDA(cudaMemsetAsync(sk.counters, 0, sizeof(unsigned) * kSplitKTiles, s));   // counters self-reset afterwards

    auto W_ = [&](const std::string& n, size_t e) { return m->tt.get(n, e); };
    auto Bv = [&](const std::string& n, size_t e) -> const float* { return inorm ? nullptr : m->tt.get(n, e); };
    // does this conv run on the tensor-core engine?  (mode 2 forces it wherever the shape is supported: test aid)
    auto tc_ok = [&](const std::string& name, int Hi, int Wi, int Cin, int Cout, int k, int stride) -> bool {
        if (!any_tc || stride != 1 || !m->tt.has(name + ".wh")) return false;
        if (!conv_tc_supported(B, 1, Hi, Wi, Cin, Cout, 1, k, k)) return false;
        // BatchNorm variant: bias / residual / ReLU ride in the conv epilogue, which the K-split launches of very long
        // reductions cannot do (they need a linear epilogue); ResNet-50's largest K is 3*3*512 = 4608
        if (!inorm && (long long)k * k * Cin > 9600) return false;
        const long long ctas = (((long long)B * Hi * Wi + 127) / 128) * ((Cout + 127) / 128);
        return m->tc_mode == 2 || ctas >= m->tc_min_ctas;
    };
    // raw = conv(in) + per-(sample, channel) sums of raw; `xs` != nullptr: split input, tensor-core engine
    auto conv_stats = [&](const std::string& name, const float* in, const float* xs, float* raw, double* sums_, int Hi, int Wi, int Cin,
                          int Cout, int k, int stride, int pad) -> int {
        const int Ho = (Hi + 2 * pad - k) / stride + 1, Wo = (Wi + 2 * pad - k) / stride + 1;
        if (xs != nullptr) {
            const int cpad = (Cout + 15) / 16 * 16;
            const size_t wn = (size_t)k * k * cpad * Cin, n_in = (size_t)B * Hi * Wi * Cin;
            const __half* wh = m->tt.get<__half>(name + ".wh", wn);
            const __half* wl = m->tt.get<__half>(name + ".wl", wn);
            const float* ws = m->tt.get(name + ".ws", 1);
            if (!wh || !wl || !ws) return -3;
            ConvTcArgs a;
            a.x_hi = reinterpret_cast<const __half*>(xs); a.x_lo = a.x_hi + n_in;
            a.w_hi = wh; a.w_lo = wl; a.scale_ptr = ws; a.bias = nullptr; a.res = nullptr; a.y = raw;
            a.B = B; a.T = 1; a.H = Hi; a.W = Wi; a.Cin = Cin; a.Cout = Cout; a.cout_pad = cpad;
            a.kt = 1; a.kh = k; a.kw = k; a.res_ut = a.res_uh = a.res_uw = 1; a.act = ACT_NONE; a.out_mode = 0;
            a.terms = 3;
            const bool fuse = conv_tc_fuses_stats(1, Ho, Wo);
            if (fuse) {
                I2V_CHECK_CUDA(cudaMemsetAsync(sums_, 0, sizeof(double) * 2 * (size_t)B * Cout, s));
                a.stats = sums_;
            }
            I2V_TRY(launch_conv_tc(a, s));
            if (!fuse) I2V_TRY(launch_channel_stats(raw, sums_, B, (long long)Ho * Wo, Cout, s));
            return 0;
        }
        const float* w = W_(name + ".w", (size_t)k * k * Cout * Cin);
        if (!w) return -3;
        I2V_TRY(conv(0, in, w, nullptr, nullptr, raw, B, 1, Hi, Wi, Cin, Cout, 1, k, k, 1, stride, stride, 0, pad, pad, 1, 1, 1, ACT_NONE, 0,
                     s, &sk));
        return launch_channel_stats(raw, sums_, B, (long long)Ho * Wo, Cout, s);
    };
    // BatchNorm variant (BN folded into weight + bias at load): out = act(conv(in) + bias [+ res]).  Stride-1 convs whose GEMM
    // fills the machine run on the tensor-core engine: their fp32 input is split into the fp16 pair first (one extra pass over
    // a tensor of a few MB; `split_src` remembers which tensor sbuf[0] holds, conv1 and the downsample conv share theirs).
    const float* split_src = nullptr;
    auto conv_bn = [&](const std::string& name, const float* in, float* out, const float* res, int Hi, int Wi, int Cin, int Cout, int k,
                       int stride, int pad, int relu) -> int {
        const float* b = m->tt.get(name + ".b", Cout);
        if (!b) return -3;
        if (tc_ok(name, Hi, Wi, Cin, Cout, k, stride)) {
            const size_t n_in = (size_t)B * Hi * Wi * Cin;
            __half* xs = reinterpret_cast<__half*>(sbuf[0]);
            if (split_src != in) {
                I2V_TRY(launch_split_fp16(in, xs, xs + n_in, ACT_SPLIT_SCALE, (long long)n_in, s));
                split_src = in;
            }
            const int cpad = (Cout + 15) / 16 * 16;
            const size_t wn = (size_t)k * k * cpad * Cin;
            const __half* wh = m->tt.get<__half>(name + ".wh", wn);
            const __half* wl = m->tt.get<__half>(name + ".wl", wn);
            const float* ws = m->tt.get(name + ".ws", 1);
            if (!wh || !wl || !ws) return -3;
            ConvTcArgs a;
            a.x_hi = xs; a.x_lo = xs + n_in; a.w_hi = wh; a.w_lo = wl; a.scale_ptr = ws; a.bias = b; a.res = res; a.y = out;
            a.B = B; a.T = 1; a.H = Hi; a.W = Wi; a.Cin = Cin; a.Cout = Cout; a.cout_pad = cpad;
            a.kt = 1; a.kh = k; a.kw = k; a.res_ut = a.res_uh = a.res_uw = 1; a.act = relu ? ACT_RELU : ACT_NONE; a.out_mode = 0;
            a.terms = 3;
            if (out == split_src) split_src = nullptr;      // (never the case today: outputs go to other buffers)
            return launch_conv_tc(a, s);
        }
        const float* w = W_(name + ".w", (size_t)k * k * Cout * Cin);
        if (!w) return -3;
        if (out == split_src) split_src = nullptr;
        return conv(0, in, w, b, res, out, B, 1, Hi, Wi, Cin, Cout, 1, k, k, 1, stride, stride, 0, pad, pad, 1, 1, 1,
                    relu ? ACT_RELU : ACT_NONE, 0, s, &sk);
    };
    // normalise (+ ReLU) pass: fp32 result, or the fp16 split a following tensor-core conv consumes
    auto norm_act = [&](const float* raw, const float* coef_, float* out, bool split, int Ho, int Wo, int C, int relu) -> int {
        ModArgs ma;
        ma.x = raw; ma.coef = coef_; ma.gb = nullptr; ma.r = nullptr; ma.coef2 = nullptr; ma.out = out;
        ma.B = B; ma.T = 1; ma.H = Ho; ma.W = Wo; ma.C = C; ma.ut = ma.uh = ma.uw = 1; ma.act = relu ? ACT_RELU : ACT_NONE;
        if (split) {
            ma.out_hi = reinterpret_cast<__half*>(out); ma.out_lo = ma.out_hi + (size_t)B * Ho * Wo * C; ma.split_scale = ACT_SPLIT_SCALE;
        }
        return launch_modulate(ma, s);
    };
    // conv (+ IN + ReLU) : in -> out ; raw conv output goes through `tmp` when instance-normalised
    auto conv_norm_relu = [&](const std::string& name, const float* in, const float* in_split, float* tmp, float* out, bool out_split,
                              int Hi, int Wi, int Cin, int Cout, int k, int stride, int pad, int relu) -> int {
        const int Ho = (Hi + 2 * pad - k) / stride + 1, Wo = (Wi + 2 * pad - k) / stride + 1;
        if (inorm) {
            I2V_TRY(conv_stats(name, in, in_split, tmp, sums, Hi, Wi, Cin, Cout, k, stride, pad));
            I2V_TRY(launch_norm_coeffs(sums, coef, B, Cout, (long long)Ho * Wo, 0, 1e-5f, nullptr, nullptr, nullptr, s));
            I2V_TRY(norm_act(tmp, coef, out, out_split, Ho, Wo, Cout, relu));
        } else {
            I2V_TRY(conv_bn(name, in, out, nullptr, Hi, Wi, Cin, Cout, k, stride, pad, relu));
        }
        return 0;
    };

    I2V_TRY(launch_resize_bilinear_nchw_to_nhwc(x0, xin, B, 3, H, W, H, W, s));
    I2V_TRY(conv_norm_relu("conv1", xin, nullptr, buf[1], buf[0], false, H, W, 3, 64, 7, 2, 3, 1));
    I2V_TRY(launch_maxpool3x3s2(buf[0], buf[1], B, H1, W1, 64, s));
    float* cur = buf[1];
    float* t1 = buf[0]; float* t2 = buf[2]; float* t3 = buf[3]; float* t4 = buf[4];
    float* cur_s = sbuf[0]; float* t1s = sbuf[1]; float* t2s = sbuf[2];
    bool cur_split = false;                 // cur_s holds the split of cur
    int Hc = H2, Wc = W2, Cc = 64;
    const int nblocks[4] = {3, 4, 6, 3}, planes[4] = {64, 128, 256, 512};
    for (int li = 0; li < 4; ++li) {
        for (int bi = 0; bi < nblocks[li]; ++bi) {
            const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
            const int pl = planes[li], stride = (li > 0 && bi == 0) ? 2 : 1;
            const int Ho = (Hc + 2 - 3) / stride + 1, Wo = (Wc + 2 - 3) / stride + 1;
            // (InstanceNorm variant: the split operands are produced by the normalise passes; the BatchNorm variant decides and
            // splits inside conv_bn)
            const bool tc1 = inorm && tc_ok(p + "conv1", Hc, Wc, Cc, pl, 1, 1), tc2 = inorm && tc_ok(p + "conv2", Hc, Wc, pl, pl, 3, stride);
            const bool tc3 = inorm && tc_ok(p + "conv3", Ho, Wo, pl, 4 * pl, 1, 1);
            const bool tcd = inorm && bi == 0 && tc_ok(p + "ds", Hc, Wc, Cc, 4 * pl, 1, stride);
            // does the next block's first conv want the split of this block's output?
            bool next_tc = false;
            if (inorm) {
                int nli = li, nbi = bi + 1;
                if (nbi == nblocks[li]) { nli = li + 1; nbi = 0; }
                if (nli < 4) {
                    const std::string np = "layer" + std::to_string(nli + 1) + "." + std::to_string(nbi) + ".";
                    next_tc = tc_ok(np + "conv1", Ho, Wo, 4 * pl, planes[nli], 1, 1) ||
                              (nbi == 0 && tc_ok(np + "ds", Ho, Wo, 4 * pl, 4 * planes[nli], 1, 2));
                }
            }
            if ((tc1 || tcd) && !cur_split) {
                I2V_TRY(launch_split_fp16(cur, reinterpret_cast<__half*>(cur_s), reinterpret_cast<__half*>(cur_s) + (size_t)B * Hc * Wc * Cc,
                                          ACT_SPLIT_SCALE, (long long)B * Hc * Wc * Cc, s));
                cur_split = true;
            }
            // conv1 1x1 -> t1 ; conv2 3x3 (stride) -> t2
            I2V_TRY(conv_norm_relu(p + "conv1", cur, tc1 ? cur_s : nullptr, t3, tc2 ? t1s : t1, tc2, Hc, Wc, Cc, pl, 1, 1, 0, 1));
            I2V_TRY(conv_norm_relu(p + "conv2", t1, tc2 ? t1s : nullptr, t3, tc3 ? t2s : t2, tc3, Hc, Wc, pl, pl, 3, stride, 1, 1));
            if (inorm) {
                // o3 raw -> t1 ; identity branch raw (downsample conv) -> t3 ; out = relu(IN(o3) + IN(ds) | h) -> t4
                I2V_TRY(conv_stats(p + "conv3", t2, tc3 ? t2s : nullptr, t1, sums, Ho, Wo, pl, 4 * pl, 1, 1, 0));
                I2V_TRY(launch_norm_coeffs(sums, coef, B, 4 * pl, (long long)Ho * Wo, 0, 1e-5f, nullptr, nullptr, nullptr, s));
                const float* idt = cur; const float* c2 = nullptr;
                if (bi == 0) {
                    I2V_TRY(conv_stats(p + "ds", cur, tcd ? cur_s : nullptr, t3, sums2, Hc, Wc, Cc, 4 * pl, 1, stride, 0));
                    I2V_TRY(launch_norm_coeffs(sums2, coef2, B, 4 * pl, (long long)Ho * Wo, 0, 1e-5f, nullptr, nullptr, nullptr, s));
                    idt = t3; c2 = coef2;
                }
                ModArgs ma;
                ma.x = t1; ma.coef = coef; ma.gb = nullptr; ma.r = idt; ma.coef2 = c2; ma.out = t4;
                ma.B = B; ma.T = 1; ma.H = Ho; ma.W = Wo; ma.C = 4 * pl; ma.ut = ma.uh = ma.uw = 1; ma.act = ACT_RELU;
                if (next_tc) {      // fp32 result (the next identity branch) AND its split (the next conv1's operand)
                    ma.out_hi = reinterpret_cast<__half*>(cur_s); ma.out_lo = ma.out_hi + (size_t)B * Ho * Wo * 4 * pl;
                    ma.split_scale = ACT_SPLIT_SCALE; ma.out_f32 = t4;
                }
                I2V_TRY(launch_modulate(ma, s));
                cur_split = next_tc;
            } else {
                const float* idt = cur;
                if (bi == 0) {
                    I2V_TRY(conv_bn(p + "ds", cur, t3, nullptr, Hc, Wc, Cc, 4 * pl, 1, stride, 0, 0));
                    idt = t3;
                }
                I2V_TRY(conv_bn(p + "conv3", t2, t4, idt, Ho, Wo, pl, 4 * pl, 1, 1, 0, 1));
            }
            float* old = cur; cur = t4; t4 = old;   // rotate: previous input buffer becomes scratch
            Hc = Ho; Wc = Wo; Cc = 4 * pl;
        }
    }
    // adaptive average pool -> fc (1x1 conv): only the first zc outputs (the mean) are consumed
    I2V_TRY(launch_channel_stats(cur, sums, B, (long long)Hc * Wc, Cc, s));
    I2V_TRY(launch_mean_from_sums(sums, pooled, B, Cc, (long long)Hc * Wc, s));
    const float* fw = m->tt.get("fc.w", (size_t)m->zc * 2048);
    const float* fb = m->tt.get("fc.b", (size_t)m->zc);
    if (!fw || !fb) return -3;
    I2V_TRY(launch_linear(pooled, fw, fb, embed, B, 2048, m->zc, ACT_NONE, s));
    return 0;
}

extern "C" {
i2v_embedder* i2v_embedder_create(int zc, int norm_mode) {
    if (zc <= 0 || (norm_mode != 0 && norm_mode != 1)) { set_error("embedder_create: bad arguments"); return nullptr; }
    auto* h = new i2v_embedder();
    h->zc = zc; h->norm_mode = norm_mode;
    return h;
}
int i2v_embedder_set_tensor(i2v_embedder* h, const char* n, const void* p, size_t b) { return h ? h->tt.set(n, p, b) : -1; }
int i2v_embedder_set_scalar(i2v_embedder* h, const char* n, double v) {
    I2V_REQUIRE(h && n, "embedder_set_scalar: null argument");
    const std::string k(n);
    if (k == "tc_min_ctas" && v >= 1 && v <= 65536) { h->tc_min_ctas = (int)v; return 0; }
    I2V_REQUIRE(k == "tc_mode" && (v == 0 || v == 1 || v == 2), "embedder_set_scalar: unknown option '%s' = %g", n, v);
    h->tc_mode = (int)v;
    return 0;
}
size_t i2v_embedder_workspace_bytes(const i2v_embedder* h, int batch, int height, int width) {
    Arena ar(nullptr, 0, true);
    embedder_run(h, nullptr, nullptr, batch, height, width, ar, nullptr, true);
    return ar.peak + 256;
}
int i2v_embedder_forward(i2v_embedder* h, const float* x0, float* embed, int batch, int height, int width, void* ws,
                         size_t wsb, void* stream) {
    I2V_REQUIRE(h && x0 && embed && ws, "embedder_forward: null argument");
    I2V_REQUIRE(batch > 0 && height >= 32 && width >= 32, "embedder_forward: bad shape B=%d H=%d W=%d", batch, height, width);
    Arena ar(ws, wsb, false);
    return embedder_run(h, x0, embed, batch, height, width, ar, static_cast<cudaStream_t>(stream), false);
}
void i2v_embedder_destroy(i2v_embedder* h) { delete h; }
}  // extern "C"

// =============================================================================== decoder
struct i2v_decoder {
    int nf, z_dim, us[2], ut[2], engine;
    TensorTable tt;
    std::unordered_map<std::string, double> scalars;   // host-side per-layer constants (split scales)
};


struct DecBlock { const char* name; int cin, cout, ut, uh, uw; };

static void decoder_blocks(const i2v_decoder* m, DecBlock out[6]) {
    const int nf = m->nf;
    const DecBlock b[6] = {{"head_0", 16 * nf, 16 * nf, 1, 1, 1},
                           {"g_0", 16 * nf, 16 * nf, 2, 2, 2},
                           {"g_1", 16 * nf, 8 * nf, 2, 2, 2},
                           {"g_2", 8 * nf, 4 * nf, 2, 2, 2},
                           {"g_3", 4 * nf, 2 * nf, m->ut[0], m->us[0], m->us[0]},
                           {"g_4", 2 * nf, 1 * nf, m->ut[1], m->us[1], m->us[1]}};
    for (int i = 0; i < 6; ++i) out[i] = b[i];
}

// Generator.forward (decoder.py:97-120) on channels-last tensors.  Per GeneratorBlock (decoder.py:33-49):
//   stats(x)                       one pass over the PRE-upsample tensor (GroupNorm statistics are
//                                  invariant under nearest upsampling)
//   gb   = SPADE maps              bilinear(img) -> conv3x3+lrelu -> conv3x3 to (gamma|beta), 2-D only
//   a0   = lrelu(GN(x)*(1+g)+b)    fused modulate pass, reads x through the upsample index map
//   dx   = conv_0(a0)+bias
//   a1   = lrelu(AdaIN(dx, z))     stats(dx) + Linear(z) + fused modulate pass
//   xs   = conv_s(GN_affine(x))    at the PRE-upsample resolution (1x1x1 conv commutes with nearest
//                                  upsampling) or x itself
//   out  = conv_1(a1)+bias+up(xs)  residual read through the upsample map in the conv epilogue
static int decoder_run(const i2v_decoder* m, const float* img, const float* z, float* frames, int B, int H, int W,
                       Arena& ar, cudaStream_t s, bool dry) {
    DecBlock blk[6];
    decoder_blocks(m, blk);
    const int zd = m->z_dim;
    // geometry walk
    int T = 1, Hc = 4, Wc = 4;
    size_t max_out = (size_t)B * 16 * blk[0].cin, max_p = 0, max_d = 0, max_gb = 0, max_sh = 0, max_low_in = 0, max_low_out = 0;
    int cmax = 0;
    for (int i = 0; i < 6; ++i) {
        const size_t low = (size_t)B * T * Hc * Wc;
        T *= blk[i].ut; Hc *= blk[i].uh; Wc *= blk[i].uw;
        const size_t vox = (size_t)B * T * Hc * Wc;
        const int cmid = blk[i].cin < blk[i].cout ? blk[i].cin : blk[i].cout;
        max_out = std::max(max_out, vox * blk[i].cout);
        max_p = std::max(max_p, vox * blk[i].cin);
        max_d = std::max(max_d, vox * cmid);
        max_gb = std::max(max_gb, (size_t)B * Hc * Wc * 2 * blk[i].cin);
        max_sh = std::max(max_sh, (size_t)B * Hc * Wc * 128);
        if (blk[i].cin != blk[i].cout) {
            max_low_in = std::max(max_low_in, low * blk[i].cin);
            max_low_out = std::max(max_low_out, low * blk[i].cout);
        }
        cmax = std::max(cmax, blk[i].cin);
    }
    I2V_REQUIRE(Hc == H && Wc == W, "decoder: geometry yields %dx%d frames, caller asked for %dx%d", Hc, Wc, H, W);
    const int Tout = T;

    float* xa = ar.take<float>(max_out);
    float* xb = ar.take<float>(max_out);
    float* bufp = ar.take<float>(std::max(max_p, (size_t)B * Tout * H * W * m->nf));
    float* bufd = ar.take<float>(max_d);
    float* gb = ar.take<float>(max_gb);
    float* sh = ar.take<float>(max_sh);
    float* imgr = ar.take<float>((size_t)B * H * W * 3);
    float* lowin = ar.take<float>(max_low_in);
    float* lowout = ar.take<float>(max_low_out);
    double* sums = ar.take<double>((size_t)B * cmax * 2);
    double* sums_mid = ar.take<double>((size_t)B * cmax * 2);
    double* sums_next = ar.take<double>((size_t)B * cmax * 2);
    float* coef = ar.take<float>((size_t)B * cmax * 2);
    float* coefs = ar.take<float>((size_t)B * cmax * 2);
    float* mod = ar.take<float>((size_t)B * 2 * cmax);
    if (!ar.ok()) { set_error("decoder: workspace too small (%zu needed, %zu given)", ar.peak, ar.cap); return -4; }
    if (dry) return 0;

    const int eng = m->engine;
    auto G = [&](const std::string& n, size_t e) { return m->tt.get(n, e); };
    const bool tc = eng >= 1;
    I2V_REQUIRE(!tc || m->nf % 16 == 0, "decoder: the tensor-core engine needs channel_factor %% 16 == 0 (got %d)", m->nf);
    // Tensor-core conv on split fp16 operands: x lives in `xbuf` as (hi | lo) halves of n_in elements each.
    auto conv_tc = [&](const std::string& wname, const float* xbuf, size_t n_in, const float* bias, const float* res, float* y,
                       int Bc, int Tc, int Hc_, int Wc_, int cin_, int cout_, int kt, int kh, int kw, int rut, int ruh, int ruw,
                       int act, int out_mode, double* stats_out = nullptr, int t_phase = 0, const float* x2buf = nullptr,
                       size_t n_in2 = 0, int cin2 = 0) -> int {
        const int cpad = (cout_ + 15) / 16 * 16;
        const size_t wn = (size_t)(t_phase ? 4 : kt) * kh * kw * cpad * cin_;
        const __half* wh = m->tt.get<__half>(wname + (t_phase ? ".wph" : ".wh"), wn);
        const __half* wl = m->tt.get<__half>(wname + (t_phase ? ".wpl" : ".wl"), wn);
        const float* ws = m->tt.get(wname + (t_phase ? ".wps" : ".ws"), 1);
        if (!wh || !wl || !ws) return -3;
        ConvTcArgs a;
        a.x_hi = reinterpret_cast<const __half*>(xbuf); a.x_lo = a.x_hi + n_in;
        a.w_hi = wh; a.w_lo = wl; a.scale_ptr = ws; a.bias = bias; a.res = res; a.y = y;
        a.B = Bc; a.T = Tc; a.H = Hc_; a.W = Wc_; a.Cin = cin_; a.Cout = cout_; a.cout_pad = cpad;
        a.kt = kt; a.kh = kh; a.kw = kw; a.res_ut = rut; a.res_uh = ruh; a.res_uw = ruw; a.act = act; a.out_mode = out_mode;
        a.terms = eng == 1 ? 3 : 1;
        a.stats = stats_out;
        a.t_phase = t_phase;
        if (x2buf != nullptr) {      // fused learned shortcut: <conv>x holds conv_s as a kw-stacked centre-tap slab
            const size_t w2n = (size_t)3 * cpad * cin2;
            const __half* w2h = m->tt.get<__half>(wname + "x.wh", w2n);
            const __half* w2l = m->tt.get<__half>(wname + "x.wl", w2n);
            if (!w2h || !w2l) return -3;
            a.x2_hi = reinterpret_cast<const __half*>(x2buf); a.x2_lo = a.x2_hi + n_in2;
            a.w2_hi = w2h; a.w2_lo = w2l; a.Cin2 = cin2;
        }
        if (stats_out) {
            // can this launch fuse the statistics?  Ask the launcher itself (tile shapes, transpose-tile room, one sample per
            // tile ...); geometries it cannot serve get the separate statistics pass instead of an error
            ConvTcArgs probe = a;
            probe.dry_run = 1;
            if (launch_conv_tc(probe, s) != 0) {
                a.stats = nullptr;
                I2V_TRY(launch_conv_tc(a, s));
                const long long V = (long long)Tc * Hc_ * Wc_;
                I2V_REQUIRE(out_mode == 0, "decoder: statistics of a frame-layout output are not defined");
                return launch_channel_stats(y, stats_out, Bc, V, cout_, s);
            }
            I2V_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, sizeof(double) * 2 * (size_t)Bc * cout_, s));
        }
        return launch_conv_tc(a, s);
    };
    // fused modulate pass writing either fp32 (SIMT engine) or the fp16 split (tensor-core engine)
    auto modulate_to = [&](const float* x_, const float* coef_, const float* gb_, float* out_, size_t n_out, int Bc, int Tc,
                           int Hc_, int Wc_, int C_, int ut_, int uh_, int uw_, int act, const float* coef_b = nullptr,
                           float* outb = nullptr) -> int {
        ModArgs ma;
        ma.x = x_; ma.coef = coef_; ma.gb = gb_; ma.r = nullptr; ma.coef2 = nullptr; ma.out = out_;
        ma.B = Bc; ma.T = Tc; ma.H = Hc_; ma.W = Wc_; ma.C = C_; ma.ut = ut_; ma.uh = uh_; ma.uw = uw_; ma.act = act;
        if (tc) {
            ma.out_hi = reinterpret_cast<__half*>(out_); ma.out_lo = ma.out_hi + n_out; ma.split_scale = ACT_SPLIT_SCALE;
            if (outb != nullptr) { ma.coef_b = coef_b; ma.outb_hi = reinterpret_cast<__half*>(outb); ma.outb_lo = ma.outb_hi + n_out; }
        }
        return launch_modulate(ma, s);
    };

    // fc: rows pre-permuted at load so the output is already [B,1,4,4,C] channels-last
    const int C0 = blk[0].cin;
    I2V_PTR(fcw, G("fc.w", (size_t)16 * C0 * zd));
    I2V_PTR(fcb, G("fc.b", (size_t)16 * C0));
    I2V_TRY(launch_linear(z, fcw, fcb, xa, B, zd, 16 * C0, ACT_NONE, s));

    float* x = xa; float* xn = xb;
    T = 1; Hc = 4; Wc = 4;
    bool have_in_stats = false;     // statistics of x already produced by the previous conv_1's epilogue
    for (int i = 0; i < 6; ++i) {
        const DecBlock& k = blk[i];
        const std::string nm = k.name;
        const int cin = k.cin, cout = k.cout, cmid = cin < cout ? cin : cout;
        const int Tl = T, Hl = Hc, Wl = Wc;
        T *= k.ut; Hc *= k.uh; Wc *= k.uw;
        const long long vlow = (long long)Tl * Hl * Wl, vhi = (long long)T * Hc * Wc;

        // statistics of the block input (pre-upsample): fused into the producing conv when possible
        if (!have_in_stats) I2V_TRY(launch_channel_stats(x, sums, B, vlow, cin, s));
        const bool fuse = tc && conv_tc_fuses_stats(T, Hc, Wc);
        // SPADE maps (normalization_layer.py:20-23)
        I2V_PTR(scw, G(nm + ".spade.conv.w", (size_t)9 * 128 * 3));
        I2V_PTR(scb, G(nm + ".spade.conv.b", 128));
        I2V_PTR(sgb, G(nm + ".spade.gb.b", (size_t)2 * cin));
        I2V_TRY(launch_resize_bilinear_nchw_to_nhwc(img, imgr, B, 3, H, W, Hc, Wc, s));
        if (!tc) {
            I2V_PTR(sgw, G(nm + ".spade.gb.w", (size_t)9 * 2 * cin * 128));
            I2V_TRY(conv(0, imgr, scw, scb, nullptr, sh, B, 1, Hc, Wc, 3, 128, 1, 3, 3, 1, 1, 1, 0, 1, 1, 1, 1, 1, ACT_LRELU02, 0, s));
            I2V_TRY(conv(0, sh, sgw, sgb, nullptr, gb, B, 1, Hc, Wc, 128, 2 * cin, 1, 3, 3, 1, 1, 1, 0, 1, 1, 1, 1, 1, ACT_NONE, 0, s));
        } else {
            // 3->128 conv stays on the SIMT engine (Cin = 3) but emits the fp16 split the gamma|beta conv consumes
            auto it = m->scalars.find(nm + ".spade.sa");
            I2V_REQUIRE(it != m->scalars.end(), "decoder: scalar '%s.spade.sa' was not registered", nm.c_str());
            const size_t nsh = (size_t)B * Hc * Wc * 128;
            ConvArgs ca;
            ca.x = imgr; ca.w = scw; ca.bias = scb; ca.res = nullptr; ca.y = nullptr;
            ca.B = B; ca.Ti = 1; ca.Hi = Hc; ca.Wi = Wc; ca.Cin = 3; ca.To = 1; ca.Ho = Hc; ca.Wo = Wc; ca.Cout = 128;
            ca.kt = 1; ca.kh = 3; ca.kw = 3; ca.st = ca.sh = ca.sw = 1; ca.pt = 0; ca.ph = ca.pw = 1;
            ca.res_ut = ca.res_uh = ca.res_uw = 1; ca.act = ACT_LRELU02; ca.out_mode = 0;
            ca.y_hi = reinterpret_cast<__half*>(sh); ca.y_lo = ca.y_hi + nsh; ca.split_scale = (float)it->second;
            // dedicated K = 27 kernel (same bits as the SIMT engine); scalar "opt.spade_simt" keeps the implicit-GEMM path (A/B switch)
            const bool spade_simt = m->scalars.count("opt.spade_simt") && m->scalars.at("opt.spade_simt") != 0;
            if (!spade_simt && spade_conv3_tiles(Hc, Wc)) I2V_TRY(launch_spade_conv3(imgr, scw, scb, ca.y_hi, ca.y_lo, ca.split_scale, B, Hc, Wc, ACT_LRELU02, s));
            else I2V_TRY(launch_conv_simt(ca, s));
            I2V_TRY(conv_tc(nm + ".spade.gb", sh, nsh, sgb, nullptr, gb, B, 1, Hc, Wc, 128, 2 * cin, 1, 3, 3, 1, 1, 1, ACT_NONE, 0));
        }
        // a0 = lrelu(GN16(x) * (1+gamma) + beta), upsampled on the fly
        int groups = 16;
        while (cin % groups) --groups;
        I2V_TRY(launch_norm_coeffs(sums, coef, B, cin, vlow, groups, 1e-5f, nullptr, nullptr, nullptr, s));
        // conv_0 behind a x2 temporal upsample: keep a0 at T/2 and let the conv use its 2-tap phase form
        // (planes >= 16x16: halo / CTA-pair kernel; smaller planes, i.e. g_0 at 8x8: per-tap kernel, one launch per phase)
        const bool phase = tc && k.ut == 2 && T % 2 == 0 && m->tt.has(nm + ".conv_0.wph");
        const int Ta = phase ? T / 2 : T;                       // stored planes of a0
        const size_t n_a0 = (size_t)B * Ta * Hc * Wc * cin;
        // a block that keeps the resolution runs its learned shortcut inside conv_1 (side input through the centre tap)
        const bool no_fuse_s = m->scalars.count("opt.no_fuse_s") && m->scalars.at("opt.no_fuse_s") != 0;   // A/B switch
        const bool fuse_s = !no_fuse_s && tc && cin != cout && k.ut == 1 && k.uh == 1 && k.uw == 1 && m->tt.has(nm + ".conv_1x.wh") &&
                            conv_tc_side_eligible(Hc, Wc, cmid, cin, (cout + 15) / 16 * 16, eng == 1 ? 3 : 1);
        // ... and its GroupNorm-affine input comes out of the same read of x as a0 (8-channel split path only)
        auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
        const bool dual = fuse_s && cin % 8 == 0 && pow2(cin / 8) && pow2(Wc);
        if (cin != cout) {
            I2V_PTR(nsw, G(nm + ".norm_s.w", cin));
            I2V_PTR(nsb, G(nm + ".norm_s.b", cin));
            I2V_TRY(launch_norm_coeffs(sums, coefs, B, cin, vlow, 16, 1e-5f, nsw, nsb, nullptr, s));
        }
        I2V_TRY(modulate_to(x, coef, gb, bufp, n_a0, B, Ta, Hc, Wc, cin, phase ? 1 : k.ut, k.uh, k.uw, ACT_LRELU02,
                            dual ? coefs : nullptr, dual ? lowin : nullptr));
        // shortcut at low resolution
        const float* xs = x;
        if (cin != cout) {
            if (!dual) I2V_TRY(modulate_to(x, coefs, nullptr, lowin, (size_t)B * vlow * cin, B, Tl, Hl, Wl, cin, 1, 1, 1, ACT_NONE));
            if (fuse_s) {
                xs = nullptr;                                     // conv_1 below takes lowin as its side input
            } else if (!tc) {
                I2V_PTR(csw, G(nm + ".conv_s.w", (size_t)cout * cin));
                I2V_TRY(conv(0, lowin, csw, nullptr, nullptr, lowout, B, Tl, Hl, Wl, cin, cout, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1,
                             ACT_NONE, 0, s));
            } else {
                I2V_TRY(conv_tc(nm + ".conv_s", lowin, (size_t)B * vlow * cin, nullptr, nullptr, lowout, B, Tl, Hl, Wl, cin, cout, 1, 1, 1,
                                1, 1, 1, ACT_NONE, 0));
            }
            if (!fuse_s) xs = lowout;
        }
        // dx = conv_0(a0)
        I2V_PTR(b0, G(nm + ".conv_0.b", cmid));
        if (!tc) {
            I2V_PTR(w0, G(nm + ".conv_0.w", (size_t)27 * cmid * cin));
            I2V_TRY(conv(0, bufp, w0, b0, nullptr, bufd, B, T, Hc, Wc, cin, cmid, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, ACT_NONE, 0, s));
        } else {
            I2V_TRY(conv_tc(nm + ".conv_0", bufp, n_a0, b0, nullptr, bufd, B, T, Hc, Wc, cin, cmid, 3, 3, 3, 1, 1, 1,
                            ACT_NONE, 0, fuse ? sums_mid : nullptr, phase ? 1 : 0));
        }
        // a1 = lrelu(AdaIN(dx, z))
        I2V_PTR(aw, G(nm + ".adain.w", (size_t)2 * cmid * zd));
        I2V_PTR(ab, G(nm + ".adain.b", (size_t)2 * cmid));
        I2V_TRY(launch_linear(z, aw, ab, mod, B, zd, 2 * cmid, ACT_NONE, s));
        if (!fuse) I2V_TRY(launch_channel_stats(bufd, sums_mid, B, vhi, cmid, s));
        I2V_TRY(launch_norm_coeffs(sums_mid, coef, B, cmid, vhi, 0, 1e-5f, nullptr, nullptr, mod, s));
        I2V_TRY(modulate_to(bufd, coef, nullptr, bufp, (size_t)B * vhi * cmid, B, T, Hc, Wc, cmid, 1, 1, 1, ACT_LRELU02));
        // out = conv_1(a1) + up(xs)
        I2V_PTR(b1, G(nm + ".conv_1.b", cout));
        if (!tc) {
            I2V_PTR(w1, G(nm + ".conv_1.w", (size_t)27 * cout * cmid));
            I2V_TRY(conv(0, bufp, w1, b1, xs, xn, B, T, Hc, Wc, cmid, cout, 3, 3, 3, 1, 1, 1, 1, 1, 1, k.ut, k.uh, k.uw, ACT_NONE, 0, s));
        } else {
            // the next block normalises this output: its statistics ride in the epilogue (not needed after g_4)
            const bool fuse_out = fuse && i < 5;
            I2V_TRY(conv_tc(nm + ".conv_1", bufp, (size_t)B * vhi * cmid, b1, xs, xn, B, T, Hc, Wc, cmid, cout, 3, 3, 3, k.ut, k.uh, k.uw,
                            ACT_NONE, 0, fuse_out ? sums_next : nullptr, 0, fuse_s ? lowin : nullptr, (size_t)B * vlow * cin,
                            fuse_s ? cin : 0));
            have_in_stats = fuse_out;
        }
        if (have_in_stats) { double* tsw = sums; sums = sums_next; sums_next = tsw; }
        float* t = x; x = xn; xn = t;
    }
    // frames = tanh(conv_img(lrelu(x)))  written as [B,T,3,H,W]
    I2V_PTR(bi, G("conv_img.b", 3));
    I2V_TRY(modulate_to(x, nullptr, nullptr, bufp, (size_t)B * T * Hc * Wc * m->nf, B, T, Hc, Wc, m->nf, 1, 1, 1, ACT_LRELU02));
    if (!tc) {
        I2V_PTR(wi, G("conv_img.w", (size_t)27 * 3 * m->nf));
        I2V_TRY(conv(0, bufp, wi, bi, nullptr, frames, B, T, Hc, Wc, m->nf, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, ACT_TANH, 1, s));
    } else {
        I2V_TRY(conv_tc("conv_img", bufp, (size_t)B * T * Hc * Wc * m->nf, bi, nullptr, frames, B, T, Hc, Wc, m->nf, 3, 3, 3, 3, 1, 1, 1,
                        ACT_TANH, 1));
    }
    return 0;
}

extern "C" {
i2v_decoder* i2v_decoder_create(int nf, int z_dim, const int us[2], const int ut[2], int engine) {
    if (nf <= 0 || nf % 4 || z_dim <= 0 || z_dim % 4 || !us || !ut) { set_error("decoder_create: bad arguments"); return nullptr; }
    auto* h = new i2v_decoder();
    h->nf = nf; h->z_dim = z_dim; h->us[0] = us[0]; h->us[1] = us[1]; h->ut[0] = ut[0]; h->ut[1] = ut[1]; h->engine = engine;
    return h;
}
int i2v_decoder_set_tensor(i2v_decoder* h, const char* n, const void* p, size_t b) { return h ? h->tt.set(n, p, b) : -1; }
int i2v_decoder_set_scalar(i2v_decoder* h, const char* n, double v) {
    I2V_REQUIRE(h && n, "decoder_set_scalar: null argument");
    h->scalars[n] = v;
    return 0;
}
size_t i2v_decoder_workspace_bytes(const i2v_decoder* h, int batch, int height, int width) {
    Arena ar(nullptr, 0, true);
    if (decoder_run(h, nullptr, nullptr, nullptr, batch, height, width, ar, nullptr, true)) return 0;
    return ar.peak + 256;
}
int i2v_decoder_forward(i2v_decoder* h, const float* img, const float* z, float* frames, int batch, int height, int width,
                        void* ws, size_t wsb, void* stream) {
    I2V_REQUIRE(h && img && z && frames && ws, "decoder_forward: null argument");
    I2V_REQUIRE(batch > 0, "decoder_forward: empty batch");
    Arena ar(ws, wsb, false);
    return decoder_run(h, img, z, frames, batch, height, width, ar, static_cast<cudaStream_t>(stream), false);
}
void i2v_decoder_destroy(i2v_decoder* h) { delete h; }
}  // extern "C"

// =============================================================================== 3-D encoder
struct i2v_encoder3d {
    int ch[5], ss[4], st[4], z_dim;
    int tc_mode = 1;      // 0: fp32 SIMT convs only  1: tensor-core engine for stride-1 convs that fill >= tc_min_ctas tiles  2: wherever supported
    int tc_min_ctas = 8;
    TensorTable tt;
};

// Encoder.forward (resnet3D.py:208-219): conv1 (3,7,7)/2 -> GN16 -> ReLU -> 4 stages x 2 BasicBlocks
// (resnet3D.py:101-135; the downsample branch is a 3x3x3 conv + GN, :185-193) -> conv_mu on the
// 4x4 map, emitted as (mu | logvar) [B, 2*z_dim]; the reparameterised sample (resnet3D.py:202-206)
// is formed by the caller from CPU noise like the reference does.
//
// Engines (round 2): the stride-1 3x3x3 convs -- conv2 of every block, conv1 of every second block, and conv1 / downsample of a
// stride-1 stage: 3/4 of the encoder's 52 GFLOP per clip -- run on the tensor-core engine (error-compensated fp16 split,
// conv_tc.cu) whenever their GEMM has >= tc_min_ctas 128 x 128 tiles: the GroupNorm+ReLU pass feeding such a conv writes the
// fp16 pair, the block's final pass writes the fp32 tensor (next residual) AND its split, and GroupNorm's sums come out of the
// conv epilogue.  Strided convs, the Cin = 3 stem and the 1x4x4 tail stay on the fp32 SIMT engine (split-K).
static int encoder3d_run(const i2v_encoder3d* m, const float* seq, float* mu, int B, int T, int H, int W, Arena& ar,
                         cudaStream_t s, bool dry) {
    auto od = [](int n, int k, int st, int p) { return (n + 2 * p - k) / st + 1; };
    int T1 = od(T, 3, 2, 1), H1 = od(H, 7, 2, 3), W1 = od(W, 7, 2, 3);
    size_t amax = (size_t)B * T1 * H1 * W1 * m->ch[0];
    int cmax = m->ch[0];
    {
        int t = T1, h = H1, w = W1;
        for (int l = 0; l < 4; ++l) {
            t = od(t, 3, m->st[l], 1); h = od(h, 3, m->ss[l], 1); w = od(w, 3, m->ss[l], 1);
            amax = std::max(amax, (size_t)B * t * h * w * m->ch[l + 1]);
            cmax = std::max(cmax, m->ch[l + 1]);
        }
        I2V_REQUIRE(t == 1 && h == 4 && w == 4,
                    "encoder3d: clip %dx%dx%d collapses to %dx%dx%d, the reference needs 1x4x4 (resnet3D.py:179,219)", T, H, W, t, h, w);
    }
    float* xin = ar.take<float>((size_t)B * T * H * W * 3);
    float* buf[4];
    for (auto& b : buf) b = ar.take<float>(amax);
    // fp16 (hi | lo) copies of tensor-core conv inputs: same bytes as the fp32 tensor
    const bool any_tc = m->tc_mode != 0;
    float* cur_s = any_tc ? ar.take<float>(amax) : nullptr;     // split of the block input / output
    float* mid_s = any_tc ? ar.take<float>(amax) : nullptr;     // split of relu(GN(conv1))
    double* sums = ar.take<double>((size_t)B * cmax * 2);
    double* sums2 = ar.take<double>((size_t)B * cmax * 2);
    float* coef = ar.take<float>((size_t)B * cmax * 2);
    float* coef2 = ar.take<float>((size_t)B * cmax * 2);
    SplitK sk;
    sk.ws = ar.take<float>(kSplitKBytes / sizeof(float)); sk.bytes = kSplitKBytes;
    sk.counters = ar.take<unsigned>(kSplitKTiles); sk.max_tiles = kSplitKTiles;
    if (!ar.ok()) { set_error("encoder3d: workspace too small (%zu needed, %zu given)", ar.peak, ar.cap); return -4; }
    if (dry) return 0;
    I2V_CHECK_CUDA(cudaMemsetAsync(sk.counters, 0, sizeof(unsigned) * kSplitKTiles, s));
    auto G = [&](const std::string& n, size_t e) { return m->tt.get(n, e); };
    // does this 3x3x3 conv run on the tensor-core engine?  (mode 2 forces it wherever the shape is supported: test aid)
    auto tc_ok = [&](const std::string& name, int Ti, int Hi, int Wi, int Cin, int Cout, int st_, int ss_) -> bool {
        if (!any_tc || st_ != 1 || ss_ != 1 || !m->tt.has(name + ".wh")) return false;
        if (!conv_tc_supported(B, Ti, Hi, Wi, Cin, Cout, 3, 3, 3)) return false;
        const long long ctas = (((long long)B * Ti * Hi * Wi + 127) / 128) * ((Cout + 127) / 128);
        return m->tc_mode == 2 || ctas >= m->tc_min_ctas;
    };
    // raw = conv3x3x3(split input) on the tensor-core engine + per-(sample, channel) sums of raw (fused where the launcher can)
    auto conv_tc_stats = [&](const std::string& name, const float* xs, float* raw, double* sums_, int Ti, int Hi, int Wi, int Cin,
                             int Cout) -> int {
        const int cpad = (Cout + 15) / 16 * 16;
        const size_t wn = (size_t)27 * cpad * Cin, n_in = (size_t)B * Ti * Hi * Wi * Cin;
        const __half* wh = m->tt.get<__half>(name + ".wh", wn);
        const __half* wl = m->tt.get<__half>(name + ".wl", wn);
        const float* ws = m->tt.get(name + ".ws", 1);
        if (!wh || !wl || !ws) return -3;
        ConvTcArgs a;
        a.x_hi = reinterpret_cast<const __half*>(xs); a.x_lo = a.x_hi + n_in;
        a.w_hi = wh; a.w_lo = wl; a.scale_ptr = ws; a.bias = nullptr; a.res = nullptr; a.y = raw;
        a.B = B; a.T = Ti; a.H = Hi; a.W = Wi; a.Cin = Cin; a.Cout = Cout; a.cout_pad = cpad;
        a.kt = a.kh = a.kw = 3; a.res_ut = a.res_uh = a.res_uw = 1; a.act = ACT_NONE; a.out_mode = 0;
        a.terms = 3;
        a.stats = sums_;
        ConvTcArgs probe = a;
        probe.dry_run = 1;
        if (launch_conv_tc(probe, s) != 0) {        // this geometry cannot fuse the statistics: separate pass
            a.stats = nullptr;
            I2V_TRY(launch_conv_tc(a, s));
            return launch_channel_stats(raw, sums_, B, (long long)Ti * Hi * Wi, Cout, s);
        }
        I2V_CHECK_CUDA(cudaMemsetAsync(sums_, 0, sizeof(double) * 2 * (size_t)B * Cout, s));
        return launch_conv_tc(a, s);
    };
    // GroupNorm (+ residual branch) + ReLU pass: fp32 result, the fp16 split a tensor-core conv consumes, or both
    auto norm_act = [&](const float* raw, const float* coef_, const float* res_, const float* coef2_, float* out_f32, float* out_split,
                        int To_, int Ho_, int Wo_, int C_) -> int {
        ModArgs ma;
        ma.x = raw; ma.coef = coef_; ma.gb = nullptr; ma.r = res_; ma.coef2 = coef2_; ma.out = out_f32;
        ma.B = B; ma.T = To_; ma.H = Ho_; ma.W = Wo_; ma.C = C_; ma.ut = ma.uh = ma.uw = 1; ma.act = ACT_RELU;
        if (out_split != nullptr) {
            ma.out_hi = reinterpret_cast<__half*>(out_split); ma.out_lo = ma.out_hi + (size_t)B * To_ * Ho_ * Wo_ * C_;
            ma.split_scale = ACT_SPLIT_SCALE; ma.out_f32 = out_f32;
        }
        return launch_modulate(ma, s);
    };

    // [B,T,3,H,W] -> [B,T,H,W,3]
    I2V_TRY(launch_resize_bilinear_nchw_to_nhwc(seq, xin, B * T, 3, H, W, H, W, s));
    I2V_PTR(w1, G("conv1.w", (size_t)147 * m->ch[0] * 3));
    I2V_PTR(n1w, G("norm1.w", m->ch[0]));
    I2V_PTR(n1b, G("norm1.b", m->ch[0]));
    I2V_TRY(conv(0, xin, w1, nullptr, nullptr, buf[1], B, T, H, W, 3, m->ch[0], 3, 7, 7, 2, 2, 2, 1, 3, 3, 1, 1, 1, ACT_NONE, 0, s, &sk));
    long long V = (long long)T1 * H1 * W1;
    I2V_TRY(launch_channel_stats(buf[1], sums, B, V, m->ch[0], s));
    I2V_TRY(launch_norm_coeffs(sums, coef, B, m->ch[0], V, 16, 1e-5f, n1w, n1b, nullptr, s));
    I2V_TRY(modulate(buf[1], coef, nullptr, nullptr, nullptr, buf[0], B, T1, H1, W1, m->ch[0], 1, 1, 1, ACT_RELU, s));
    float* cur = buf[0]; float* ta = buf[1]; float* tb = buf[2]; float* tc = buf[3];
    int Tc = T1, Hc = H1, Wc = W1, Cc = 64;   // resnet3D.py:140 hard-codes inplanes = 64
    I2V_REQUIRE(m->ch[0] == 64, "encoder3d: channels[0] must be 64 (resnet3D.py:140)");
    bool cur_split = false;                 // cur_s holds the split of cur
    for (int l = 0; l < 4; ++l) {
        const int pl = m->ch[l + 1];
        for (int bi = 0; bi < 2; ++bi) {
            const std::string p = "layer." + std::to_string(l) + "." + std::to_string(bi) + ".";
            const int st = bi == 0 ? m->st[l] : 1, ss = bi == 0 ? m->ss[l] : 1;
            const int To = od(Tc, 3, st, 1), Ho = od(Hc, 3, ss, 1), Wo = od(Wc, 3, ss, 1);
            const long long Vo = (long long)To * Ho * Wo;
            const bool has_ds = m->tt.has(p + "ds.w");
            I2V_PTR(g1w, G(p + "bn1.w", pl)); I2V_PTR(g1b, G(p + "bn1.b", pl));
            I2V_PTR(g2w, G(p + "bn2.w", pl)); I2V_PTR(g2b, G(p + "bn2.b", pl));
            const bool tc1 = tc_ok(p + "conv1", Tc, Hc, Wc, Cc, pl, st, ss), tc2 = tc_ok(p + "conv2", To, Ho, Wo, pl, pl, 1, 1);
            const bool tcd = has_ds && tc_ok(p + "ds", Tc, Hc, Wc, Cc, pl, st, ss);
            // does the next block's conv1 / downsample conv want the split of this block's output?
            bool next_tc = false;
            {
                const int nl = bi == 0 ? l : l + 1, nbi = bi == 0 ? 1 : 0;
                if (nl < 4) {
                    const std::string np = "layer." + std::to_string(nl) + "." + std::to_string(nbi) + ".";
                    const int nst = nbi == 0 ? m->st[nl] : 1, nss = nbi == 0 ? m->ss[nl] : 1;
                    next_tc = tc_ok(np + "conv1", To, Ho, Wo, pl, m->ch[nl + 1], nst, nss) ||
                              (m->tt.has(np + "ds.w") && tc_ok(np + "ds", To, Ho, Wo, pl, m->ch[nl + 1], nst, nss));
                }
            }
            if ((tc1 || tcd) && !cur_split) {
                const long long n = (long long)B * Tc * Hc * Wc * Cc;
                I2V_TRY(launch_split_fp16(cur, reinterpret_cast<__half*>(cur_s), reinterpret_cast<__half*>(cur_s) + n, ACT_SPLIT_SCALE, n, s));
                cur_split = true;
            }
            // o = relu(GN(conv1(h)))
            if (tc1) {
                I2V_TRY(conv_tc_stats(p + "conv1", cur_s, ta, sums, Tc, Hc, Wc, Cc, pl));
            } else {
                I2V_PTR(wc1, G(p + "conv1.w", (size_t)27 * pl * Cc));
                I2V_TRY(conv(0, cur, wc1, nullptr, nullptr, ta, B, Tc, Hc, Wc, Cc, pl, 3, 3, 3, st, ss, ss, 1, 1, 1, 1, 1, 1, ACT_NONE, 0, s, &sk));
                I2V_TRY(launch_channel_stats(ta, sums, B, Vo, pl, s));
            }
            I2V_TRY(launch_norm_coeffs(sums, coef, B, pl, Vo, 16, 1e-5f, g1w, g1b, nullptr, s));
            I2V_TRY(norm_act(ta, coef, nullptr, nullptr, tc2 ? nullptr : tb, tc2 ? mid_s : nullptr, To, Ho, Wo, pl));
            // o = GN(conv2(o))
            if (tc2) {
                I2V_TRY(conv_tc_stats(p + "conv2", mid_s, ta, sums, To, Ho, Wo, pl, pl));
            } else {
                I2V_PTR(wc2, G(p + "conv2.w", (size_t)27 * pl * pl));
                I2V_TRY(conv(0, tb, wc2, nullptr, nullptr, ta, B, To, Ho, Wo, pl, pl, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, ACT_NONE, 0, s, &sk));
                I2V_TRY(launch_channel_stats(ta, sums, B, Vo, pl, s));
            }
            I2V_TRY(launch_norm_coeffs(sums, coef, B, pl, Vo, 16, 1e-5f, g2w, g2b, nullptr, s));
            const float* res = cur; const float* c2 = nullptr;
            if (has_ds) {
                I2V_PTR(gdw, G(p + "ds.gn.w", pl)); I2V_PTR(gdb, G(p + "ds.gn.b", pl));
                if (tcd) {
                    I2V_TRY(conv_tc_stats(p + "ds", cur_s, tb, sums2, Tc, Hc, Wc, Cc, pl));
                } else {
                    I2V_PTR(wd, G(p + "ds.w", (size_t)27 * pl * Cc));
                    I2V_TRY(conv(0, cur, wd, nullptr, nullptr, tb, B, Tc, Hc, Wc, Cc, pl, 3, 3, 3, st, ss, ss, 1, 1, 1, 1, 1, 1, ACT_NONE, 0, s, &sk));
                    I2V_TRY(launch_channel_stats(tb, sums2, B, Vo, pl, s));
                }
                I2V_TRY(launch_norm_coeffs(sums2, coef2, B, pl, Vo, 16, 1e-5f, gdw, gdb, nullptr, s));
                res = tb; c2 = coef2;
            } else {
                I2V_REQUIRE(st == 1 && ss == 1 && Cc == pl, "encoder3d: block %s needs a downsample branch but none was registered", p.c_str());
            }
            // h = relu(GN(o) + identity): fp32 (the next residual / SIMT input) and, when a tensor-core conv follows, its split
            I2V_TRY(norm_act(ta, coef, res, c2, tc, next_tc ? cur_s : nullptr, To, Ho, Wo, pl));
            cur_split = next_tc;
            float* old = cur; cur = tc; tc = old;
            Tc = To; Hc = Ho; Wc = Wo; Cc = pl;
        }
    }
    // conv_mu | conv_var (4x4 valid convs on the 4x4 map = one Linear over the flattened (h,w,c) map)
    I2V_PTR(mw, G("muvar.w", (size_t)2 * m->z_dim * 16 * Cc));
    I2V_PTR(mb, G("muvar.b", (size_t)2 * m->z_dim));
    I2V_TRY(launch_linear(cur, mw, mb, mu, B, 16 * Cc, 2 * m->z_dim, ACT_NONE, s));
    return 0;
}

extern "C" {
i2v_encoder3d* i2v_encoder3d_create(const int channels[5], const int stride_s[4], const int stride_t[4], int z_dim) {
    if (!channels || !stride_s || !stride_t || z_dim <= 0) { set_error("encoder3d_create: bad arguments"); return nullptr; }
    auto* h = new i2v_encoder3d();
    for (int i = 0; i < 5; ++i) h->ch[i] = channels[i];
    for (int i = 0; i < 4; ++i) { h->ss[i] = stride_s[i]; h->st[i] = stride_t[i]; }
    h->z_dim = z_dim;
    return h;
}
int i2v_encoder3d_set_tensor(i2v_encoder3d* h, const char* n, const void* p, size_t b) { return h ? h->tt.set(n, p, b) : -1; }
int i2v_encoder3d_set_scalar(i2v_encoder3d* h, const char* n, double v) {
    I2V_REQUIRE(h && n, "encoder3d_set_scalar: null argument");
    const std::string k(n);
    if (k == "tc_min_ctas" && v >= 1 && v <= 65536) { h->tc_min_ctas = (int)v; return 0; }
    I2V_REQUIRE(k == "tc_mode" && (v == 0 || v == 1 || v == 2), "encoder3d_set_scalar: unknown option '%s' = %g", n, v);
    h->tc_mode = (int)v;
    return 0;
}
size_t i2v_encoder3d_workspace_bytes(const i2v_encoder3d* h, int batch, int frames, int height, int width) {
    Arena ar(nullptr, 0, true);
    if (encoder3d_run(h, nullptr, nullptr, batch, frames, height, width, ar, nullptr, true)) return 0;
    return ar.peak + 256;
}
int i2v_encoder3d_forward(i2v_encoder3d* h, const float* seq, float* mu, int batch, int frames, int height, int width,
                          void* ws, size_t wsb, void* stream) {
    I2V_REQUIRE(h && seq && mu && ws, "encoder3d_forward: null argument");
    Arena ar(ws, wsb, false);
    return encoder3d_run(h, seq, mu, batch, frames, height, width, ar, static_cast<cudaStream_t>(stream), false);
}
void i2v_encoder3d_destroy(i2v_encoder3d* h) { delete h; }

// =============================================================================== single ops
int i2v_op_conv(const float* x, const float* w, const float* bias, const float* res, float* y, int B, int Ti, int Hi, int Wi,
                int Cin, int Cout, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, int rut, int ruh,
                int ruw, int act, int out_mode, int engine, void* stream) {
    I2V_REQUIRE(x && w && y, "op_conv: null argument");
    return conv(engine, x, w, bias, res, y, B, Ti, Hi, Wi, Cin, Cout, kt, kh, kw, st, sh, sw, pt, ph, pw, rut, ruh, ruw, act, out_mode,
                static_cast<cudaStream_t>(stream));
}
int i2v_op_spade_conv3(const float* img, const float* w, const float* bias, void* y_hi, void* y_lo, float split_scale, int B, int H,
                       int W, int act, void* stream) {
    I2V_REQUIRE(img && w && bias && y_hi && y_lo, "op_spade_conv3: null argument");
    return launch_spade_conv3(img, w, bias, static_cast<__half*>(y_hi), static_cast<__half*>(y_lo), split_scale, B, H, W, act,
                              static_cast<cudaStream_t>(stream));
}
int i2v_op_conv_tc(const float* x, const float* w, const float* bias, const float* res, float* y, int B, int T, int H, int W, int Cin,
                   int Cout, int cout_pad, int kt, int kh, int kw, int rut, int ruh, int ruw, int act, int out_mode, int terms,
                   int variant, float scale_a, float scale_w, void* ws, size_t ws_bytes, void* stream) {
    I2V_REQUIRE(x && w && y && ws, "op_conv_tc: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nx = (size_t)B * T * H * W * Cin, nw = (size_t)kt * kh * kw * cout_pad * Cin;
    Arena ar(ws, ws_bytes, false);
    __half* xh = ar.take<__half>(nx); __half* xl = ar.take<__half>(nx);
    __half* wh = ar.take<__half>(nw); __half* wl = ar.take<__half>(nw);
    float* sc = ar.take<float>(1);
    I2V_REQUIRE(ar.ok(), "op_conv_tc: workspace too small (%zu needed)", ar.peak);
    I2V_TRY(launch_split_fp16(x, xh, xl, scale_a, (long long)nx, s));
    I2V_TRY(launch_split_fp16(w, wh, wl, scale_w, (long long)nw, s));
    const float inv = 1.f / (scale_a * scale_w);
    I2V_CHECK_CUDA(cudaMemcpyAsync(sc, &inv, sizeof(float), cudaMemcpyHostToDevice, s));
    ConvTcArgs a;
    a.x_hi = xh; a.x_lo = xl; a.w_hi = wh; a.w_lo = wl; a.scale_ptr = sc; a.bias = bias; a.res = res; a.y = y;
    a.B = B; a.T = T; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.cout_pad = cout_pad; a.kt = kt; a.kh = kh; a.kw = kw;
    a.res_ut = rut; a.res_uh = ruh; a.res_uw = ruw; a.act = act; a.out_mode = out_mode; a.terms = terms; a.variant = variant;
    return launch_conv_tc(a, s);
}
int i2v_op_conv_tc_phase(const float* x, const float* w, const float* bias, float* y, int B, int T, int H, int W, int Cin, int Cout,
                         int cout_pad, int terms, int variant, float scale_a, float scale_w, void* ws, size_t ws_bytes, void* stream) {
    I2V_REQUIRE(x && w && y && ws, "op_conv_tc_phase: null argument");
    I2V_REQUIRE(T % 2 == 0, "op_conv_tc_phase: T must be even (x holds T/2 planes)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nx = (size_t)B * (T / 2) * H * W * Cin, nw = (size_t)4 * 9 * cout_pad * Cin;
    Arena ar(ws, ws_bytes, false);
    __half* xh = ar.take<__half>(nx); __half* xl = ar.take<__half>(nx);
    __half* wh = ar.take<__half>(nw); __half* wl = ar.take<__half>(nw);
    float* sc = ar.take<float>(1);
    I2V_REQUIRE(ar.ok(), "op_conv_tc_phase: workspace too small (%zu needed)", ar.peak);
    I2V_TRY(launch_split_fp16(x, xh, xl, scale_a, (long long)nx, s));
    I2V_TRY(launch_split_fp16(w, wh, wl, scale_w, (long long)nw, s));
    const float inv = 1.f / (scale_a * scale_w);
    I2V_CHECK_CUDA(cudaMemcpyAsync(sc, &inv, sizeof(float), cudaMemcpyHostToDevice, s));
    ConvTcArgs a;
    a.x_hi = xh; a.x_lo = xl; a.w_hi = wh; a.w_lo = wl; a.scale_ptr = sc; a.bias = bias; a.res = nullptr; a.y = y;
    a.B = B; a.T = T; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.cout_pad = cout_pad; a.kt = 3; a.kh = 3; a.kw = 3;
    a.res_ut = a.res_uh = a.res_uw = 1; a.act = ACT_NONE; a.out_mode = 0; a.terms = terms; a.variant = variant; a.t_phase = 1;
    return launch_conv_tc(a, s);
}
int i2v_op_conv_tc_side(const float* x, const float* w, const float* x2, const float* w2, const float* bias, float* y, int B, int T,
                        int H, int W, int Cin, int Cin2, int Cout, int cout_pad, int act, int out_mode, int terms, int variant,
                        float scale_a, float scale_w, void* ws, size_t ws_bytes, void* stream) {
    I2V_REQUIRE(x && w && x2 && w2 && y && ws, "op_conv_tc_side: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nx = (size_t)B * T * H * W * Cin, nw = (size_t)27 * cout_pad * Cin;
    const size_t nx2 = (size_t)B * T * H * W * Cin2, nw2 = (size_t)3 * cout_pad * Cin2;
    Arena ar(ws, ws_bytes, false);
    __half* xh = ar.take<__half>(nx); __half* xl = ar.take<__half>(nx);
    __half* wh = ar.take<__half>(nw); __half* wl = ar.take<__half>(nw);
    __half* x2h = ar.take<__half>(2 * nx2);                       // (hi | lo) halves back to back, like the decoder's buffers
    __half* w2h = ar.take<__half>(nw2); __half* w2l = ar.take<__half>(nw2);
    float* sc = ar.take<float>(1);
    I2V_REQUIRE(ar.ok(), "op_conv_tc_side: workspace too small (%zu needed)", ar.peak);
    I2V_TRY(launch_split_fp16(x, xh, xl, scale_a, (long long)nx, s));
    I2V_TRY(launch_split_fp16(w, wh, wl, scale_w, (long long)nw, s));
    I2V_TRY(launch_split_fp16(x2, x2h, x2h + nx2, scale_a, (long long)nx2, s));
    I2V_TRY(launch_split_fp16(w2, w2h, w2l, scale_w, (long long)nw2, s));
    const float inv = 1.f / (scale_a * scale_w);
    I2V_CHECK_CUDA(cudaMemcpyAsync(sc, &inv, sizeof(float), cudaMemcpyHostToDevice, s));
    ConvTcArgs a;
    a.x_hi = xh; a.x_lo = xl; a.w_hi = wh; a.w_lo = wl; a.scale_ptr = sc; a.bias = bias; a.res = nullptr; a.y = y;
    a.B = B; a.T = T; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.cout_pad = cout_pad; a.kt = 3; a.kh = 3; a.kw = 3;
    a.res_ut = a.res_uh = a.res_uw = 1; a.act = act; a.out_mode = out_mode; a.terms = terms; a.variant = variant;
    a.x2_hi = x2h; a.x2_lo = x2h + nx2; a.w2_hi = w2h; a.w2_lo = w2l; a.Cin2 = Cin2;
    return launch_conv_tc(a, s);
}
int i2v_debug_flow_timestamps(void* buf) { return flow_set_debug(static_cast<unsigned long long*>(buf)); }
int i2v_debug_conv_tc_timestamps(void* buf, int ctas) { return conv_tc_set_debug(static_cast<unsigned long long*>(buf), ctas); }
int i2v_op_channel_stats(const float* x, double* sums, int B, int64_t V, int C, void* stream) {
    return launch_channel_stats(x, sums, B, V, C, static_cast<cudaStream_t>(stream));
}
int i2v_op_norm_coeffs(const double* sums, float* coef, int B, int C, int64_t V, int groups, float eps, const float* gamma,
                       const float* beta, const float* mod, void* stream) {
    return launch_norm_coeffs(sums, coef, B, C, V, groups, eps, gamma, beta, mod, static_cast<cudaStream_t>(stream));
}
int i2v_op_modulate(const float* x, const float* coef, const float* gb, const float* r, const float* coef2, float* out, int B,
                    int T, int H, int W, int C, int ut, int uh, int uw, int act, void* stream) {
    return modulate(x, coef, gb, r, coef2, out, B, T, H, W, C, ut, uh, uw, act, static_cast<cudaStream_t>(stream));
}
int i2v_op_modulate_split(const float* x, const float* coef, const float* gb, void* out_hi, void* out_lo, int B, int T, int H, int W,
                          int C, int ut, int uh, int uw, int act, float split_scale, const float* coef_b, void* outb_hi, void* outb_lo,
                          void* stream) {
    I2V_REQUIRE(x && out_hi && out_lo, "op_modulate_split: null argument");
    ModArgs m;
    m.x = x; m.coef = coef; m.gb = gb; m.r = nullptr; m.coef2 = nullptr; m.out = nullptr;
    m.B = B; m.T = T; m.H = H; m.W = W; m.C = C; m.ut = ut; m.uh = uh; m.uw = uw; m.act = act;
    m.out_hi = static_cast<__half*>(out_hi); m.out_lo = static_cast<__half*>(out_lo); m.split_scale = split_scale;
    m.coef_b = coef_b; m.outb_hi = static_cast<__half*>(outb_hi); m.outb_lo = static_cast<__half*>(outb_lo);
    return launch_modulate(m, static_cast<cudaStream_t>(stream));
}
int i2v_op_linear(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act, void* stream) {
    return launch_linear(x, w, bias, y, B, K, N, act, static_cast<cudaStream_t>(stream));
}
int i2v_op_resize_bilinear(const float* img, float* out, int B, int C, int H0, int W0, int H, int W, void* stream) {
    return launch_resize_bilinear_nchw_to_nhwc(img, out, B, C, H0, W0, H, W, static_cast<cudaStream_t>(stream));
}
int i2v_op_preprocess_u8(const unsigned char* img, float* out, int H0, int W0, int H, int W, int bgr, void* stream) {
    I2V_REQUIRE(img && out, "op_preprocess_u8: null argument");
    return launch_preprocess_u8(img, out, H0, W0, H, W, bgr, static_cast<cudaStream_t>(stream));
}
int i2v_op_frames_max(const float* frames, float* mx, int64_t n, void* stream) {
    I2V_REQUIRE(frames && mx && n > 0, "op_frames_max: null argument or empty clip");
    return launch_frames_max(frames, mx, n, static_cast<cudaStream_t>(stream));
}
int i2v_op_frames_to_u8(const float* frames, const float* mx, unsigned char* out, int N, int T, int H, int W, int64_t sn, int64_t st,
                        int64_t sh, void* stream) {
    I2V_REQUIRE(frames && mx && out && N > 0 && T > 0 && H > 0 && W > 0, "op_frames_to_u8: null argument or empty clip");
    return launch_frames_to_u8(frames, mx, out, N, T, H, W, sn, st, sh, static_cast<cudaStream_t>(stream));
}
int i2v_op_maxpool3x3s2(const float* x, float* y, int B, int H, int W, int C, void* stream) {
    return launch_maxpool3x3s2(x, y, B, H, W, C, static_cast<cudaStream_t>(stream));
}
}  // extern "C"
