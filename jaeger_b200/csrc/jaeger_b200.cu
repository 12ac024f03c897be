// C-ABI implementation of the B200-native `jaeger predict` hot path (see include/jaeger_b200.h).
#include "../../include/jaeger_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include <zlib.h>

#include "conv_launch.cuh"
#include "conv_resident.cuh"
#include "dust_kernels.cuh"
#include "model_kernels.cuh"
#include "post_kernels.cuh"
#include "seq_kernels.cuh"
#include "termini_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return 2;
}
#define JG_CUDA(x)                                          \
  do {                                                      \
    cudaError_t e_ = (x);                                   \
    if (e_ != cudaSuccess) return cuda_fail(e_, #x);        \
  } while (0)

inline int grid_for(long long n, int block, int num_sms, int per_sm = 8) {
  long long g = (n + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

inline uint16_t f32_to_f16(float f) {
  const __half h = __float2half_rn(f);        // host-callable; saturates to inf only beyond 65504 (weights never are)
  uint16_t u;
  std::memcpy(&u, &h, 2);
  return u;
}

}  // namespace

struct jg_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  int64_t launches = 0;
};

// layer descriptor integer fields (mirrored in jaeger_b200/plan.py)
enum LayerField {
  LF_KIND = 0, LF_CIN, LF_COUT, LF_K, LF_DIL, LF_PAD_LEFT, LF_SHRINK, LF_IN_BUF, LF_OUT_BUF,
  LF_SC_BUF, LF_ACT1, LF_HAS_AFF2, LF_ACT2, LF_TAP_MODE, LF_TAP_SLOT, LF_POOL_MODE, LF_MASK_IN,
  LF_MASK_OUT, LF_SC_MASK, LF_MASKING, LF_CUM_SHRINK_IN, LF_HALVINGS, LF_DYT1, LF_DYT2, LF_EPI_F32, LF_LEN_CEIL,
  LF_REAL_CIN, LF_REAL_COUT,     // channel counts before the plan's padding to 64 (0 = not given)
  LF_LN1, LF_LN_EPS,             // the first norm is a MaskedLayerNormalization (scale1 = gamma, shift1 = beta, bias kept apart); its
                                 // epsilon as float bits
  LF_MASK_THR                    // valid taps an output row needs: 0 / 1 "any", (k + 1) / 2 "majority", k "strict" (layers.py:1245-1252)
};
// layer kinds: 1 = conv (fused epilogue), 2 = MaxPooling1D(2) per frame, 3 = frame sum + global max pool
enum LayerPtr { LP_KERNEL = 0, LP_BIAS, LP_SCALE1, LP_SHIFT1, LP_SCALE2, LP_SHIFT2, LP_SC_CONST, LP_TAP_MEAN,
                LP_DYT_G1, LP_DYT_B1, LP_DYT_G2, LP_DYT_B2, LP_KERNEL_ODD };

// The device weight images of one convolution (or of one slice of its output channels)
struct WeightImages {
  jg::act_t* w = nullptr;      // single-CTA kernel (w_index)
  jg::act_t* w2 = nullptr;     // CTA-pair kernel (w2_index)
  jg::act_t* w3 = nullptr;     // weights-stationary kernel (ws::w3_index), 128-output-channel layers only
};

struct Layer {
  int32_t f[JG_LAYER_INT_FIELDS];
  jg::act_t* w = nullptr;      // weights image of the single-CTA kernel (w_index)
  jg::act_t* w2 = nullptr;     // weights image of the CTA-pair kernel (w2_index)
  jg::act_t* w3 = nullptr;     // weights image of the weights-stationary kernel (ws::w3_index), 128-output-channel layers only
  WeightImages odd;            // strided convs: the images of the kernel for an odd input frame length (plan.py:split_phases)
  // A layer whose weights fit no kernel's shared memory (e.g. 256 -> 256 channels, k5) runs as slices of 64 output
  // channels on the CTA-pair kernel: one launch per slice, each reading the whole input.  [parity][slice]
  std::vector<WeightImages> slices[2];
  int slice_width = 0;
  float* par = nullptr;   // bias, scale1, shift1, scale2, shift2, sc_const, dyt gamma1, beta1, gamma2, beta2  (10 x cout)
  int* shifts = nullptr;  // device copy for the mask kernel
  bool folded = false;         // scale1 folded into the weights (par scale1 == 1)
  float* w_tap = nullptr;      // stem with an NMD tap on one-hot input: fp16-rounded weights [k][64][cout] + their tap sum [64][cout]
  int shifts_h[jg::kMaxTaps];
  int halo_l = 0, halo_r = 0;
  const char* last_kernel = "";  // which conv kernel the last forward pass launched for this layer
};

struct jg_model {
  jg_ctx* ctx = nullptr;
  std::vector<Layer> layers;
  int frames = 6, tok_offset = 1;
  int n_bufs = 0, n_masks = 0, n_taps = 0, tap_width = 0;
  std::vector<int> buf_channels;
  jg_head_desc head{};
  float *cls_w = nullptr, *cls_b = nullptr, *rel_w1 = nullptr, *rel_b1 = nullptr, *rel_w2 = nullptr,
        *rel_b2 = nullptr, *tap_mean = nullptr, *mlp_w1 = nullptr, *mlp_b1 = nullptr, *mlp_w2 = nullptr,
        *mlp_b2 = nullptr;
  int final_mask = 0;
  int conv_impl = 0;            // JG_CONV_IMPL: 0 auto (weights-stationary kernel where eligible), 1 single-CTA kernel only,
                                // 2 CTA-pair kernel wherever eligible, 3 the round-1 choice (no weights-stationary kernel)
  std::vector<int> tap_mask_slot;
  // workspace -------------------------------------------------------------------------------
  long long cap_rows = 0, cap_windows = 0;
  std::vector<jg::act_t*> bufs;
  std::vector<uint8_t*> masks;
  int* counts = nullptr;        // [n_masks][cap_windows]
  float* tap_sum = nullptr;     // [n_taps][cap_windows][tap_width]
  float* pool = nullptr;        // [cap_windows][feat]
  const int** tap_count_ptrs = nullptr;
  int* err = nullptr;
  int64_t ws_bytes = 0;
  // window-resident kernel for narrow stacks (conv_resident.cuh): candidate decided at model creation, geometry checked per call
  bool rs_ok = true;                  // every layer so far fits the resident kernel's envelope
  std::vector<uint8_t> rs_w, rs_p;    // host images: weights [layer][tap][kc][32][8] fp16, parameters 512 B per layer
  uint8_t* rs_block = nullptr;        // device copy: weights, then parameters
  jg::rs::ResidentParams rs_par{};
  int rs_stem_span = 0;
  bool last_resident = false;         // the last forward pass ran the resident kernel
  // per-conv-launch CUDA-event timing (jg_model_set_profiling)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;   // pending (start, stop) pairs
  std::vector<int> prof_layer;
  std::vector<double> prof_ms;        // accumulated per layer
  std::vector<int64_t> prof_launches;
  std::vector<double> prof_rows;      // accumulated rows (tiles*128) per layer
};

namespace {

// frame length a layer sees: floor halving after MaxPool(2) stages, ceil halving after SAME stride-2 convs
inline int len_round_of(const Layer& L) { return L.f[LF_LEN_CEIL] ? (1 << L.f[LF_HALVINGS]) - 1 : 0; }
inline int layer_len(const Layer& L, int lc) { return (lc - L.f[LF_CUM_SHRINK_IN] + len_round_of(L)) >> L.f[LF_HALVINGS]; }

void model_geometry(const jg_model* m, int lc, int* period, int* rpw) {
  int p = lc;
  for (const Layer& L : m->layers) {
    if (L.f[LF_KIND] != 1) continue;
    const int l_in = layer_len(L, lc);
    const int halo = L.halo_l > L.halo_r ? L.halo_l : L.halo_r;
    const int need = l_in + (L.f[LF_SHRINK] == 0 ? halo : 0);
    if (need > p) p = need;
  }
  *period = p;
  // a multiple of 256 rows so that the CTA-pair kernel always sees an even tile count
  *rpw = (m->frames * p + 2 * jg::kTileM - 1) / (2 * jg::kTileM) * (2 * jg::kTileM);
}

long long bytes_per_window(const jg_model* m, int rpw) {
  long long b = 0;
  for (int c : m->buf_channels) b += static_cast<long long>(rpw) * c * 2;
  b += static_cast<long long>(m->n_masks) * rpw;
  b += static_cast<long long>(m->n_masks) * 4 + static_cast<long long>(m->n_taps) * m->tap_width * 4 +
       static_cast<long long>(m->head.feat_dim) * 4;
  return b;
}

void free_workspace(jg_model* m) {
  for (auto p : m->bufs) cudaFree(p);
  for (auto p : m->masks) cudaFree(p);
  m->bufs.clear();
  m->masks.clear();
  cudaFree(m->counts); m->counts = nullptr;
  cudaFree(m->tap_sum); m->tap_sum = nullptr;
  cudaFree(m->pool); m->pool = nullptr;
  cudaFree(m->tap_count_ptrs); m->tap_count_ptrs = nullptr;
  m->cap_rows = m->cap_windows = 0;
  m->ws_bytes = 0;
}

int ensure_workspace(jg_model* m, long long n_windows, long long rows) {
  if (rows <= m->cap_rows && n_windows <= m->cap_windows) return 0;
  free_workspace(m);
  const long long plane = rows + 2 * jg::kGuardRows;
  int64_t total = 0;
  for (int c : m->buf_channels) {
    jg::act_t* p = nullptr;
    const size_t bytes = static_cast<size_t>(c / 64) * plane * 128;
    JG_CUDA(cudaMalloc(&p, bytes));
    JG_CUDA(cudaMemsetAsync(p, 0, bytes, m->ctx->stream));
    m->bufs.push_back(p);
    total += bytes;
  }
  for (int i = 0; i < m->n_masks; ++i) {
    uint8_t* p = nullptr;
    JG_CUDA(cudaMalloc(&p, plane));
    JG_CUDA(cudaMemsetAsync(p, 0, plane, m->ctx->stream));
    m->masks.push_back(p);
    total += plane;
  }
  JG_CUDA(cudaMalloc(&m->counts, static_cast<size_t>(m->n_masks) * n_windows * 4));
  const size_t tap_bytes = static_cast<size_t>(m->n_taps > 0 ? m->n_taps : 1) * n_windows * (m->tap_width > 0 ? m->tap_width : 1) * 4;
  JG_CUDA(cudaMalloc(&m->tap_sum, tap_bytes));
  JG_CUDA(cudaMalloc(&m->pool, static_cast<size_t>(n_windows) * m->head.feat_dim * 4));
  JG_CUDA(cudaMalloc(&m->tap_count_ptrs, sizeof(int*) * (m->n_taps > 0 ? m->n_taps : 1)));
  std::vector<const int*> ptrs(m->n_taps > 0 ? m->n_taps : 1, nullptr);
  for (int t = 0; t < m->n_taps; ++t) ptrs[t] = m->counts + static_cast<long long>(m->tap_mask_slot[t]) * n_windows;
  JG_CUDA(cudaMemcpyAsync(m->tap_count_ptrs, ptrs.data(), sizeof(int*) * ptrs.size(), cudaMemcpyHostToDevice, m->ctx->stream));
  JG_CUDA(cudaStreamSynchronize(m->ctx->stream));
  total += static_cast<int64_t>(m->n_masks) * n_windows * 4 + tap_bytes + static_cast<int64_t>(n_windows) * m->head.feat_dim * 4;
  m->cap_rows = rows;
  m->cap_windows = n_windows;
  m->ws_bytes = total;
  return 0;
}

// fp32 TF-layout kernel [k][cin][cout_total] (output channels [co0, co0 + cout)) -> the device images of the conv kernels
int build_weight_images(const float* wk, int k, int cin, int cout_total, int co0, int cout, bool want_w3, WeightImages* out) {
  std::vector<uint16_t> img(static_cast<size_t>(k) * cin * cout);
  auto at = [&](int t, int ci, int co) { return f32_to_f16(wk[(static_cast<size_t>(t) * cin + ci) * cout_total + co0 + co]); };
  for (int t = 0; t < k; ++t)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co) img[jg::w_index(t, ci, co, cin, cout)] = at(t, ci, co);
  JG_CUDA(cudaMalloc(&out->w, img.size() * 2));
  JG_CUDA(cudaMemcpy(out->w, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  for (int t = 0; t < k; ++t)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co) img[jg::tc2::w2_index(t, ci, co, cin, cout, k)] = at(t, ci, co);
  JG_CUDA(cudaMalloc(&out->w2, img.size() * 2));
  JG_CUDA(cudaMemcpy(out->w2, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  if (want_w3) {       // transposed image for the tensor-memory resident A operand
    for (int t = 0; t < k; ++t)
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co) img[jg::ws::w3_index(t, ci, co, cin, k)] = at(t, ci, co);
    JG_CUDA(cudaMalloc(&out->w3, img.size() * 2));
    JG_CUDA(cudaMemcpy(out->w3, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  }
  return 0;
}
void free_weight_images(WeightImages& w) { cudaFree(w.w); cudaFree(w.w2); cudaFree(w.w3); w = WeightImages{}; }

int upload_f32(const float* h, size_t n, float** d) {
  JG_CUDA(cudaMalloc(d, n * 4));
  JG_CUDA(cudaMemcpy(*d, h, n * 4, cudaMemcpyHostToDevice));
  return 0;
}

// Window-resident kernel: wire the layers' buffers / masks (the plan's buffer ids -> the kernel's X / H arrays) and upload the images.
// Leaves m->rs_ok false when the plan's dataflow is not the in-place two-buffer chain the kernel implements.
int finish_resident(jg_model* m, const jg_head_desc* head) {
  const int n = static_cast<int>(m->layers.size());
  bool ok = m->rs_ok && n >= 2 && (m->n_taps == 0 || m->tap_width == 64) && head->feat_dim == 64 && !std::getenv("JG_NO_RESIDENT");
  if (const char* e = std::getenv("JG_RESIDENT")) ok = ok && std::atoi(e) != 0;
  int arr_of[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mask_of[3] = {-2, -2, -2}, next_arr = 1;
  for (int l = 0; ok && l < n; ++l) {
    const Layer& L = m->layers[l];
    jg::rs::LayerRs& R = m->rs_par.layer[l];
    const int ib = L.f[LF_IN_BUF], ob = L.f[LF_OUT_BUF], sb = L.f[LF_SC_BUF];
    if (ib < 0 || ib >= 8 || ob >= 8 || sb >= 8) { ok = false; break; }
    if (l == 0) {
      mask_of[0] = L.f[LF_MASK_IN];
      R.in_arr = 0;
    } else {
      if (ib == 0 || arr_of[ib] == 0) { ok = false; break; }
      R.in_arr = arr_of[ib];
    }
    if (mask_of[R.in_arr] != L.f[LF_MASK_IN]) { ok = false; break; }
    R.has_sc = sb >= 0 ? 1 : 0;
    R.sc_arr = 0;
    if (sb >= 0) {
      if (sb == 0 || arr_of[sb] == 0 || arr_of[sb] == R.in_arr) { ok = false; break; }
      R.sc_arr = arr_of[sb];
      R.sc_all_valid = L.f[LF_SC_MASK] < 0 ? 1 : 0;
      if (L.f[LF_SC_MASK] >= 0 && mask_of[R.sc_arr] != L.f[LF_SC_MASK]) { ok = false; break; }
    }
    R.out_arr = 0;
    if (ob >= 0) {
      if (ob == 0) { ok = false; break; }
      if (arr_of[ob] == 0) {
        if (next_arr > 2) { ok = false; break; }
        arr_of[ob] = next_arr++;
      }
      R.out_arr = arr_of[ob];
      if (R.out_arr == R.in_arr || (R.sc_arr != 0 && R.sc_arr != R.out_arr)) { ok = false; break; }
      mask_of[R.out_arr] = L.f[LF_MASK_OUT];
    }
    if ((L.f[LF_POOL_MODE] != 0) != (l == n - 1) || (ob < 0) != (l == n - 1)) { ok = false; break; }
    // the compile-time epilogue shape, if the layer has one (conv_resident.cuh:EpiModeRs)
    const bool gelu1 = R.folded && R.act1 == jg::ACT_GELU_TANH;
    R.mode = jg::rs::EPI_RS_GENERIC;
    if (R.tap_mode != 0) R.mode = jg::rs::EPI_RS_GENERIC;
    else if (gelu1 && !R.has_aff2 && R.pool_mode == 0 && R.out_arr != 0) R.mode = R.has_sc ? jg::rs::EPI_RS_LIGHT_SC : jg::rs::EPI_RS_LIGHT;
    else if (gelu1 && R.has_sc && R.has_aff2 && R.act2 == jg::ACT_GELU_TANH && R.pool_mode == 2 && R.out_arr == 0) R.mode = jg::rs::EPI_RS_FINAL_SUM;
    if (std::getenv("JG_RS_GENERIC")) R.mode = jg::rs::EPI_RS_GENERIC;
  }
  m->rs_ok = ok;
  if (!ok) { m->rs_w.clear(); m->rs_p.clear(); return 0; }
  m->rs_par.n_layers = n;
  m->rs_par.w_bytes = static_cast<uint32_t>(m->rs_w.size());
  JG_CUDA(cudaMalloc(&m->rs_block, m->rs_w.size() + m->rs_p.size()));
  JG_CUDA(cudaMemcpy(m->rs_block, m->rs_w.data(), m->rs_w.size(), cudaMemcpyHostToDevice));
  JG_CUDA(cudaMemcpy(m->rs_block + m->rs_w.size(), m->rs_p.data(), m->rs_p.size(), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

extern "C" {

const char* jg_last_error(void) { return g_err.c_str(); }
int jg_version(void) { return 1; }

int jg_ctx_create(int device, jg_ctx** out) { return jg_ctx_create_on_stream(device, nullptr, out); }

int jg_ctx_create_on_stream(int device, void* stream, jg_ctx** out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail("no CUDA device: the jaeger_b200 hot path has no CPU fallback");
  if (device < 0 || device >= n) return fail("invalid device index");
  JG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  JG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(std::string("jaeger_b200 is built for sm_100a only; found ") + prop.name);
  jg_ctx* c = new jg_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  // keep stream-ordered scratch allocations cached across synchronisations (the default pool gives
  // memory back to the driver at every sync, and re-acquiring it stalls synchronous callers)
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (stream) {
    c->stream = static_cast<cudaStream_t>(stream);
    c->owns_stream = false;
  } else {
    JG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  }
  *out = c;
  return 0;
}
int jg_ctx_destroy(jg_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}
int jg_ctx_sync(jg_ctx* ctx) {
  JG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
void* jg_ctx_stream(jg_ctx* ctx) { return ctx->stream; }
int64_t jg_ctx_launch_count(jg_ctx* ctx) { return ctx->launches; }

// ---- stage 0: FASTA ingest (host) ----------------------------------------------------------------
// One pass over the file (plain or gzip, through zlib): record name = header up to the first whitespace, sequence = the record's
// lines with line ends and blanks removed (what pyfastx hands to seqops/io.py:98-104).
namespace {
struct FastaScan { int64_t n_records = 0, n_bases = 0, name_bytes = 0; };

int fasta_walk(const char* path, FastaScan* scan, uint8_t* bases, int64_t* offsets, char* names) {
  gzFile fh = gzopen(path, "rb");            // reads plain and gzip-compressed files alike
  if (!fh) return fail(std::string("cannot open ") + path);
  gzbuffer(fh, 1 << 20);
  std::vector<char> buf(1 << 22);
  int64_t rec = -1, nb = 0, nn = 0;
  bool in_header = false, header_name_done = false, at_line_start = true;
  int got;
  // line-wise over each chunk: memchr finds the line end, sequence lines are block-copied (the
  // per-character path only runs for lines that contain blanks)
  while ((got = gzread(fh, buf.data(), static_cast<unsigned>(buf.size()))) > 0) {
    const char* p = buf.data();
    const char* end = p + got;
    while (p < end) {
      if (at_line_start && *p == '>') {
        ++rec;
        if (offsets) offsets[rec] = nb;
        in_header = true; header_name_done = false; at_line_start = false;
        ++p;
        continue;
      }
      const char* nl = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
      const char* seg_end = nl ? nl : end;
      if (in_header) {
        for (const char* q = p; q < seg_end && !header_name_done; ++q) {
          if (*q == ' ' || *q == '\t' || *q == '\r') header_name_done = true;
          else { if (names) names[nn] = *q; ++nn; }
        }
      } else if (rec >= 0 && seg_end > p) {
        const char* last = seg_end;
        if (nl && last[-1] == '\r') --last;                       // CRLF line end
        const size_t len = static_cast<size_t>(last - p);
        const bool blanks = std::memchr(p, ' ', len) || std::memchr(p, '\t', len) || std::memchr(p, '\r', len);
        if (!blanks) {
          if (bases) std::memcpy(bases + nb, p, len);
          nb += static_cast<int64_t>(len);
        } else {
          for (const char* q = p; q < last; ++q)
            if (*q != ' ' && *q != '\t' && *q != '\r') { if (bases) bases[nb] = static_cast<uint8_t>(*q); ++nb; }
        }
      }
      if (seg_end > p) at_line_start = false;
      if (nl) {
        if (in_header) { if (names) names[nn] = 0; ++nn; in_header = false; }
        at_line_start = true;
        p = nl + 1;
      } else {
        p = end;
      }
    }
  }
  if (in_header) { if (names) names[nn] = 0; ++nn; }
  const bool read_error = got < 0;
  gzclose(fh);
  if (read_error) return fail(std::string("read error (corrupt gzip stream?) in ") + path);
  if (offsets) offsets[rec + 1] = nb;
  scan->n_records = rec + 1; scan->n_bases = nb; scan->name_bytes = nn;
  return 0;
}
}  // namespace

int jg_fasta_scan(const char* path, int64_t* n_records, int64_t* n_bases, int64_t* name_bytes) {
  FastaScan sc;
  if (fasta_walk(path, &sc, nullptr, nullptr, nullptr)) return 1;
  *n_records = sc.n_records; *n_bases = sc.n_bases; *name_bytes = sc.name_bytes;
  return 0;
}

int jg_fasta_load(const char* path, uint8_t* h_bases, int64_t* h_offsets, char* h_names) {
  FastaScan sc;
  return fasta_walk(path, &sc, h_bases, h_offsets, h_names);
}

// ---- stage 0b: chunked FASTA ingest (host) --------------------------------------------------------------
// A resumable reader: every call hands back the next run of WHOLE records holding at most max_bases bases (one record when
// a single record is longer), so a file of any size streams through two fixed pinned buffers while the previous chunk is
// on the device.  A byte range [begin, end) makes the reader own exactly the records whose '>' lies inside it: N ranks
// parse N disjoint slices of one plain file, nothing is read twice.
struct jg_fasta_reader {
  gzFile fh = nullptr;
  std::vector<char> buf;
  int64_t buf_pos = 0, buf_len = 0;      // unread part of buf
  int64_t file_pos = 0;                  // uncompressed offset of buf[buf_pos]
  int64_t byte_end = -1;                 // records starting at or after this offset belong to the next reader (-1: none)
  bool at_line_start = true, eof = false, done = false;
  // the record under construction when a call ran out of room (its header is parsed, some of its bases may be buffered)
  bool have_pending = false;
  std::string pend_name;
  std::vector<uint8_t> pend_bases;
  bool in_header = false, header_name_done = false;

  bool fill() {
    if (eof) return false;
    const int got = gzread(fh, buf.data(), static_cast<unsigned>(buf.size()));
    if (got <= 0) { eof = true; return false; }
    buf_pos = 0; buf_len = got;
    return true;
  }
};

int jg_fasta_open(const char* path, int64_t byte_begin, int64_t byte_end, jg_fasta_reader** out) {
  jg_fasta_reader* r = new jg_fasta_reader();
  r->fh = gzopen(path, "rb");
  if (!r->fh) { delete r; return fail(std::string("cannot open ") + path); }
  gzbuffer(r->fh, 1 << 20);
  r->buf.resize(1 << 22);
  r->byte_end = byte_end;
  if (byte_begin > 0) {
    if (!gzdirect(r->fh)) { gzclose(r->fh); delete r; return fail("byte ranges need an uncompressed FASTA file"); }
    // start one byte early: a '>' at byte_begin only opens a record when the byte before it ends a line
    if (gzseek(r->fh, byte_begin - 1, SEEK_SET) < 0) { gzclose(r->fh); delete r; return fail("seek failed"); }
    r->file_pos = byte_begin - 1;
    r->at_line_start = false;
    // skip to the first record start inside the range
    bool found = false;
    while (!found) {
      if (r->buf_pos >= r->buf_len && !r->fill()) break;
      while (r->buf_pos < r->buf_len) {
        const char c = r->buf[r->buf_pos];
        if (r->at_line_start && c == '>' && r->file_pos >= byte_begin) { found = true; break; }
        r->at_line_start = c == '\n';
        ++r->buf_pos; ++r->file_pos;
      }
    }
    if (!found) r->done = true;
  }
  *out = r;
  return 0;
}

int jg_fasta_close(jg_fasta_reader* r) {
  if (!r) return 0;
  if (r->fh) gzclose(r->fh);
  delete r;
  return 0;
}

// Next chunk: up to max_records whole records with at most max_bases bases in total (always at least one record if any is
// left and it fits cap_bases).  h_offsets gets n + 1 entries, h_names the NUL-terminated names back to back.
// Returns 0 and *n_records = 0 at the end of the reader's range; 3 when the next record alone exceeds cap_bases (the caller
// grows its buffer to *need_bases and calls again: the record is kept).
int jg_fasta_next(jg_fasta_reader* r, int64_t max_bases, int64_t cap_bases, int64_t max_records, int64_t cap_name_bytes,
                  uint8_t* h_bases, int64_t* h_offsets, char* h_names, int64_t* n_records, int64_t* n_bases,
                  int64_t* name_bytes, int64_t* need_bases) {
  int64_t nrec = 0, nb = 0, nn = 0;
  *n_records = *n_bases = *name_bytes = 0;
  if (need_bases) *need_bases = 0;
  h_offsets[0] = 0;
  auto emit_pending = [&]() -> int {        // move the finished pending record into the caller's buffers
    const int64_t len = static_cast<int64_t>(r->pend_bases.size());
    if (len > cap_bases) { if (need_bases) *need_bases = len; return 3; }
    if (nrec > 0 && (nb + len > max_bases || nb + len > cap_bases || nrec >= max_records ||
                     nn + static_cast<int64_t>(r->pend_name.size()) + 1 > cap_name_bytes)) return 1;      // chunk is full: keep it for the next call
    if (nn + static_cast<int64_t>(r->pend_name.size()) + 1 > cap_name_bytes) return fail("record name longer than the name buffer");
    std::memcpy(h_bases + nb, r->pend_bases.data(), static_cast<size_t>(len));
    nb += len;
    std::memcpy(h_names + nn, r->pend_name.c_str(), r->pend_name.size() + 1);
    nn += static_cast<int64_t>(r->pend_name.size()) + 1;
    ++nrec;
    h_offsets[nrec] = nb;
    r->have_pending = false;
    r->pend_name.clear();
    r->pend_bases.clear();
    return 0;
  };
  bool record_complete = false;      // the pending record has seen its last line
  for (;;) {
    if (r->done && !r->have_pending) break;
    if (r->done || record_complete) {
      const int rc = emit_pending();
      record_complete = false;
      if (rc == 1) break;
      if (rc != 0) { *n_records = nrec; *n_bases = nb; *name_bytes = nn; return rc; }
      if (r->done) break;
      continue;
    }
    if (r->buf_pos >= r->buf_len && !r->fill()) { r->done = true; if (r->in_header) r->in_header = false; continue; }
    const char* base = r->buf.data();
    const char* p = base + r->buf_pos;
    const char* end = base + r->buf_len;
    while (p < end) {
      if (r->at_line_start && *p == '>') {
        const int64_t here = r->file_pos + (p - (base + r->buf_pos));
        if (r->have_pending) { record_complete = true; break; }          // the previous record is finished: emit it first
        if (r->byte_end >= 0 && here >= r->byte_end) { r->done = true; break; }
        r->have_pending = true;
        r->in_header = true; r->header_name_done = false; r->at_line_start = false;
        ++p;
        continue;
      }
      const char* nl = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
      const char* seg_end = nl ? nl : end;
      if (r->in_header) {
        for (const char* q = p; q < seg_end && !r->header_name_done; ++q) {
          if (*q == ' ' || *q == '\t' || *q == '\r') r->header_name_done = true;
          else r->pend_name.push_back(*q);
        }
      } else if (r->have_pending && seg_end > p) {
        const char* last = seg_end;
        if (nl && last[-1] == '\r') --last;
        const size_t len = static_cast<size_t>(last - p);
        if (!std::memchr(p, ' ', len) && !std::memchr(p, '\t', len) && !std::memchr(p, '\r', len)) {
          r->pend_bases.insert(r->pend_bases.end(), reinterpret_cast<const uint8_t*>(p), reinterpret_cast<const uint8_t*>(last));
        } else {
          for (const char* q = p; q < last; ++q)
            if (*q != ' ' && *q != '\t' && *q != '\r') r->pend_bases.push_back(static_cast<uint8_t>(*q));
        }
      }
      if (seg_end > p) r->at_line_start = false;
      if (nl) {
        if (r->in_header) r->in_header = false;
        r->at_line_start = true;
        p = nl + 1;
      } else {
        p = end;
      }
    }
    const int64_t consumed = p - (base + r->buf_pos);
    r->buf_pos += consumed;
    r->file_pos += consumed;
  }
  *n_records = nrec; *n_bases = nb; *name_bytes = nn;
  return 0;
}

// ---- stage 1 -----------------------------------------------------------------------------------
int jg_pack_bases(jg_ctx* ctx, const uint8_t* d_ascii, int64_t n, uint32_t* d_codes, uint32_t* d_valid) {
  if (n <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  const long long groups = (n + 31) / 32;
  jg::pack_bases_kernel<<<grid_for(groups, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(d_ascii, n, d_codes, d_valid);
  ctx->launches++;
  JG_CUDA(cudaGetLastError());
  return 0;
}

// ---- stage 1b: low-complexity soft-mask ----------------------------------------------------------
int jg_dust_mask(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid, const int64_t* d_core_begin,
                 const int64_t* d_core_end, const int64_t* d_contig_begin, const int64_t* d_contig_end, int64_t n_chunks,
                 int32_t threshold, uint32_t* d_soft) {
  if (n_chunks <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  const int block = 64;
  jg::dust_kernel<<<static_cast<unsigned>((n_chunks + block - 1) / block), block, 0, ctx->stream>>>(
      d_codes, d_valid, reinterpret_cast<const long long*>(d_core_begin), reinterpret_cast<const long long*>(d_core_end),
      reinterpret_cast<const long long*>(d_contig_begin), reinterpret_cast<const long long*>(d_contig_end), n_chunks,
      threshold, d_soft);
  ctx->launches++;
  JG_CUDA(cudaGetLastError());
  return 0;
}

// ---- stage 2a ----------------------------------------------------------------------------------
int jg_plan_windows(const int64_t* h_len, int64_t n_contigs, int32_t fsize, int32_t stride,
                    int32_t dynamic_stride, double dynamic_stride_threshold, int32_t min_len,
                    int64_t max_len, int32_t short_pass, int64_t* n_windows, int32_t* h_contig,
                    int64_t* h_start, int32_t* h_nbases, int32_t* h_ordinal, uint8_t* h_is_last) {
  if (fsize <= 0 || stride <= 0) return fail("fsize and stride must be positive");
  int64_t n = 0;
  const bool write = h_contig != nullptr;
  std::vector<int64_t> idx;
  for (int64_t c = 0; c < n_contigs; ++c) {
    const int64_t len = h_len[c];
    if (max_len > 0 && len > max_len) continue;
    idx.clear();
    int32_t nb = fsize;
    if (len >= fsize) {
      if (!dynamic_stride || static_cast<double>(len) >= dynamic_stride_threshold * fsize) {
        for (int64_t s = 0; s < len - (fsize - 1); s += stride) idx.push_back(s);
      } else {
        // seqops/io.py:56-71
        int64_t nw = static_cast<int64_t>(std::ceil(static_cast<double>(len) / static_cast<double>(fsize)));
        if (nw < 1) nw = 1;
        if (nw == 1) {
          idx.push_back(0);
        } else {
          const double raw = static_cast<double>(len - fsize) / static_cast<double>(nw - 1);
          std::vector<int64_t> tmp(nw);
          for (int64_t i = 0; i < nw; ++i) tmp[i] = static_cast<int64_t>(std::nearbyint(static_cast<double>(i) * raw));
          tmp[nw - 1] = len - fsize;
          std::unordered_set<int64_t> seen;
          for (int64_t v : tmp)
            if (seen.insert(v).second) idx.push_back(v);
        }
      }
    } else if (len >= min_len && short_pass) {
      idx.push_back(0);
      nb = static_cast<int32_t>(len);
    }
    for (size_t i = 0; i < idx.size(); ++i) {
      if (write) {
        h_contig[n] = static_cast<int32_t>(c);
        h_start[n] = idx[i];
        h_nbases[n] = nb;
        h_ordinal[n] = static_cast<int32_t>(i);
        h_is_last[n] = (i + 1 == idx.size());
      }
      ++n;
    }
  }
  *n_windows = n;
  return 0;
}

// ---- stage 2b ----------------------------------------------------------------------------------
int jg_encode_windows(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid, const uint32_t* d_soft,
                      const int64_t* d_win_base, const int32_t* d_win_nbases, int64_t n_windows,
                      int32_t crop, int32_t lc, int32_t pitch, const uint8_t* h_lut64,
                      int32_t case_sensitive, uint8_t* d_tokens, int32_t* d_counts, int16_t* d_skew100) {
  if (n_windows <= 0) return 0;
  if (pitch % 4 != 0 || pitch < lc) return fail("token pitch must be a multiple of 4 and >= lc");
  JG_CUDA(cudaSetDevice(ctx->device));
  jg::CodonLut lut;                       // 64 bytes, passed by value: no allocation or copy on the hot path
  std::memcpy(lut.v, h_lut64, 64);
  const int words_c = (crop + 15) / 16 + 1, words_b = (crop + 31) / 32 + 1;
  const size_t smem = static_cast<size_t>(words_c + 2 * words_b) * 4 * jg::kEncWarps;
  long long grid = (n_windows + jg::kEncWarps - 1) / jg::kEncWarps;
  const long long cap = static_cast<long long>(ctx->num_sms) * 16;
  if (grid > cap) grid = cap;
  jg::encode_windows_kernel<<<static_cast<int>(grid), jg::kEncThreads, smem, ctx->stream>>>(
      d_codes, d_valid, d_soft, reinterpret_cast<const long long*>(d_win_base), d_win_nbases, n_windows, crop,
      lc, pitch, lut, case_sensitive, d_tokens, d_counts, d_skew100);
  ctx->launches++;
  JG_CUDA(cudaGetLastError());
  return 0;
}

// ---- stage 3 -----------------------------------------------------------------------------------
int jg_model_create(jg_ctx* ctx, const jg_layer_desc* layers, int32_t n_layers, const jg_head_desc* head,
                    int32_t frames, int32_t tok_offset, jg_model** out) {
  JG_CUDA(cudaSetDevice(ctx->device));
  jg_model* m = new jg_model();
  m->ctx = ctx;
  m->frames = frames;
  m->tok_offset = tok_offset;
  m->head = *head;
  int max_buf = -1, max_mask = -1, max_tap = -1;
  for (int l = 0; l < n_layers; ++l) {
    Layer L;
    std::memcpy(L.f, layers[l].i, sizeof(L.f));
    const int cin = L.f[LF_CIN], cout = L.f[LF_COUT], k = L.f[LF_K];
    if (L.f[LF_KIND] == 2 || L.f[LF_KIND] == 3 || L.f[LF_KIND] == 4) {
      if (cin % 64 != 0) { delete m; return fail("pooling / row-phase layers need a channel count that is a multiple of 64"); }
      for (int b : {L.f[LF_IN_BUF], L.f[LF_OUT_BUF]}) max_buf = b > max_buf ? b : max_buf;
      for (int sl : {L.f[LF_MASK_IN], L.f[LF_MASK_OUT]}) max_mask = sl > max_mask ? sl : max_mask;
      m->rs_ok = false;                 // pooling / row-phase layers are outside the resident kernel's envelope
      m->layers.push_back(L);
      continue;
    }
    if (L.f[LF_KIND] != 1) { delete m; return fail("unknown layer kind in plan"); }
    if (cin % 64 != 0 || cout % 64 != 0 || cout > 256 || k > jg::kMaxTaps) {
      delete m;
      return fail("conv layer outside the tensor-core kernel's envelope (Cin, Cout multiples of 64, Cout <= 256, k <= 16)");
    }
    int mn = 0, mx = 0;
    for (int t = 0; t < k; ++t) {
      L.shifts_h[t] = t * L.f[LF_DIL] - L.f[LF_PAD_LEFT];
      mn = L.shifts_h[t] < mn ? L.shifts_h[t] : mn;
      mx = L.shifts_h[t] > mx ? L.shifts_h[t] : mx;
    }
    L.halo_l = -mn;
    L.halo_r = mx;
    if (L.halo_l > jg::kGuardRows - 8 || L.halo_r > jg::kGuardRows - 8) { delete m; return fail("conv halo exceeds guard rows"); }
    // weights: TF layout [k][cin][cout] fp32 -> swizzled fp16 shared-memory images
    const float* wk = layers[l].p[LP_KERNEL];
    // The first affine's per-channel scale (BatchNorm gamma / sigma) is folded into the fp16 weights, so
    // the epilogue adds the shift only.  Not for a layer whose NMD tap reads the raw conv output in the
    // epilogue (tap mode 1 without the linear stem tap): that needs the unscaled accumulator.
    const bool linear_tap_layer = l == 0 && L.f[LF_TAP_MODE] == 1 && cin == 64;
    L.folded = (L.f[LF_TAP_MODE] != 1 || linear_tap_layer) && layers[l].p[LP_SCALE1] != nullptr && !L.f[LF_LN1] && !std::getenv("JG_NO_BN_FOLD");
    const float* wk_raw = wk;
    const float* wk_odd = layers[l].p[LP_KERNEL_ODD];
    std::vector<float> wfold, wfold_odd;
    if (L.folded) {
      const float* sc1 = layers[l].p[LP_SCALE1];
      wfold.resize(static_cast<size_t>(k) * cin * cout);
      for (size_t i = 0; i < wfold.size(); ++i) wfold[i] = wk[i] * sc1[i % cout];
      wk = wfold.data();
      if (wk_odd) {
        wfold_odd.resize(wfold.size());
        for (size_t i = 0; i < wfold_odd.size(); ++i) wfold_odd[i] = wk_odd[i] * sc1[i % cout];
        wk_odd = wfold_odd.data();
      }
    }
    // does the layer fit one of the kernels as a whole?  If not and it is wide, it runs as slices of 64 output channels
    jg::ConvParams geo{};
    geo.cin = cin; geo.cout = cout; geo.ntaps = k; geo.halo_l = L.halo_l; geo.halo_r = L.halo_r; geo.n_tiles = 2;
    const bool fits_whole = jg::conv_tc_stages(geo) >= 0 || jg::conv_tc2_eligible(geo);
    jg::ConvParams geo_s = geo;
    geo_s.cout = 64;
    if (!fits_whole && cout > 64 && jg::conv_tc2_eligible(geo_s) && !std::getenv("JG_NO_SLICES")) L.slice_width = 64;
    for (int par = 0; par < 2; ++par) {
      const float* src = par == 0 ? wk : wk_odd;
      if (!src) continue;
      if (L.slice_width) {
        for (int co0 = 0; co0 < cout; co0 += L.slice_width) {
          WeightImages wi;
          if (build_weight_images(src, k, cin, cout, co0, L.slice_width, false, &wi)) { delete m; return 2; }
          L.slices[par].push_back(wi);
        }
      } else {
        WeightImages wi;
        if (build_weight_images(src, k, cin, cout, 0, cout, cout == 128 && k * cin / 2 <= jg::ws::kWColsMax, &wi)) { delete m; return 2; }
        if (par == 0) { L.w = wi.w; L.w2 = wi.w2; L.w3 = wi.w3; } else { L.odd = wi; }
      }
    }
    if (linear_tap_layer) {   // linear stem tap (stem_tap_kernel): the raw conv output from the unscaled fp32 weights
      std::vector<float> wt(static_cast<size_t>(k + 1) * 64 * cout, 0.0f);
      for (int t = 0; t < k; ++t)
        for (int ci = 0; ci < 64; ++ci)
          for (int co = 0; co < cout; ++co) {
            const float v = wk_raw[(static_cast<size_t>(t) * cin + ci) * cout + co];
            wt[(static_cast<size_t>(t) * 64 + ci) * cout + co] = v;
            wt[(static_cast<size_t>(k) * 64 + ci) * cout + co] += v;
          }
      JG_CUDA(cudaMalloc(&L.w_tap, wt.size() * 4));
      JG_CUDA(cudaMemcpy(L.w_tap, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
    }
    std::vector<float> par(10 * static_cast<size_t>(cout), 0.0f);
    const int order[10] = {LP_BIAS, LP_SCALE1, LP_SHIFT1, LP_SCALE2, LP_SHIFT2, LP_SC_CONST, LP_DYT_G1, LP_DYT_B1, LP_DYT_G2, LP_DYT_B2};
    for (int a = 0; a < 10; ++a) {
      const float* src = layers[l].p[order[a]];
      for (int c = 0; c < cout; ++c) par[a * cout + c] = src ? src[c] : ((order[a] == LP_SCALE1 || order[a] == LP_SCALE2) ? 1.0f : 0.0f);
    }
    if (L.folded)
      for (int c = 0; c < cout; ++c) par[1 * cout + c] = 1.0f;        // scale1 now lives in the weights
    if (upload_f32(par.data(), par.size(), &L.par)) { delete m; return 2; }
    {   // window-resident kernel: does this layer fit its envelope?  (<= 32 real channels, <= 8 taps within the guard rows, plain
        // affine + tanh-GELU / ReLU epilogue, no NMD tap, no stride)  If so, append its weight image and parameter block.
      const int n_prev = static_cast<int>(m->layers.size());
      const int rcin = L.f[LF_REAL_CIN] > 0 ? L.f[LF_REAL_CIN] : cin, rcout = L.f[LF_REAL_COUT] > 0 ? L.f[LF_REAL_COUT] : cout;
      const bool stem = n_prev == 0;
      bool ok = m->rs_ok && n_prev < jg::rs::kMaxLayersRs && cout == 64 && rcout <= 32 && k <= jg::rs::kMaxTapsRs &&
                (stem ? (cin == 64 && L.f[LF_IN_BUF] == 0) : (cin == 64 && rcin <= 32)) && L.f[LF_DYT1] == 0 &&
                L.f[LF_DYT2] == 0 && L.f[LF_LN1] == 0 && L.f[LF_MASK_THR] <= 1 && L.f[LF_EPI_F32] == 0 && L.f[LF_HALVINGS] == 0 && L.f[LF_LEN_CEIL] == 0 && wk_odd == nullptr &&
                L.f[LF_ACT1] <= 2 && L.f[LF_ACT2] <= 2 && (stem || (L.halo_l <= jg::rs::kGuardRs && L.halo_r <= jg::rs::kGuardRs));
      if (ok) {
        const int kc = stem ? 8 : 4;
        const size_t off = m->rs_w.size();
        m->rs_w.resize(off + static_cast<size_t>(k) * kc * 512);
        uint16_t* img = reinterpret_cast<uint16_t*>(m->rs_w.data() + off);
        for (int t = 0; t < k; ++t)
          for (int ci = 0; ci < kc * 8; ++ci)
            for (int co = 0; co < 32; ++co)
              img[((static_cast<size_t>(t) * kc + ci / 8) * 32 + co) * 8 + ci % 8] = f32_to_f16(wk[(static_cast<size_t>(t) * cin + ci) * cout + co]);
        const size_t poff = m->rs_p.size();
        m->rs_p.resize(poff + jg::rs::kParBytesRs);
        float* pf = reinterpret_cast<float*>(m->rs_p.data() + poff);
        uint16_t* ph = reinterpret_cast<uint16_t*>(m->rs_p.data() + poff + 256);
        for (int c = 0; c < 32; ++c) {
          pf[c] = par[2 * cout + c];                                                  // shift1
          pf[32 + c] = par[1 * cout + c];                                             // scale1 (1 when folded)
          ph[c] = f32_to_f16(L.f[LF_HAS_AFF2] ? par[3 * cout + c] : 1.0f);            // scale2
          ph[32 + c] = f32_to_f16(L.f[LF_HAS_AFF2] ? par[4 * cout + c] : 0.0f);       // shift2
          ph[64 + c] = f32_to_f16(par[5 * cout + c]);                                 // shortcut value at masked rows
          pf[128 + c] = par[0 * cout + c];                                            // conv bias (raw-output NMD tap)
        }
        jg::rs::LayerRs& R = m->rs_par.layer[n_prev];
        R = jg::rs::LayerRs{};
        R.ntaps = k;
        for (int t = 0; t < k; ++t) { R.shifts[t] = L.shifts_h[t]; if (L.shifts_h[t] == 0) R.zero_tap = 1; }
        R.act1 = L.f[LF_ACT1]; R.act2 = L.f[LF_ACT2]; R.has_aff2 = L.f[LF_HAS_AFF2]; R.pool_mode = L.f[LF_POOL_MODE];
        R.masking = L.f[LF_MASKING]; R.folded = L.folded ? 1 : 0; R.shrink_in = L.f[LF_CUM_SHRINK_IN]; R.shrink = L.f[LF_SHRINK];
        R.kc = kc; R.w_off = static_cast<uint32_t>(off);
        // NMD taps: summed in the kernel, except a stem tap on the raw conv output, which stem_tap_kernel takes from token counts
        R.tap_mode = linear_tap_layer ? 0 : L.f[LF_TAP_MODE];
        R.tap_slot = L.f[LF_TAP_SLOT];
        R.count_id = (L.f[LF_TAP_MODE] != 0 || L.f[LF_POOL_MODE] != 0) ? L.f[LF_MASK_OUT] : -1;
        if (stem) m->rs_stem_span = L.halo_l + L.halo_r;
      }
      m->rs_ok = ok;
    }
    JG_CUDA(cudaMalloc(&L.shifts, sizeof(int) * jg::kMaxTaps));
    JG_CUDA(cudaMemcpy(L.shifts, L.shifts_h, sizeof(int) * k, cudaMemcpyHostToDevice));
    for (int b : {L.f[LF_IN_BUF], L.f[LF_OUT_BUF], L.f[LF_SC_BUF]}) max_buf = b > max_buf ? b : max_buf;
    for (int s : {L.f[LF_MASK_IN], L.f[LF_MASK_OUT], L.f[LF_SC_MASK]}) max_mask = s > max_mask ? s : max_mask;
    if (L.f[LF_TAP_MODE] != 0) max_tap = L.f[LF_TAP_SLOT] > max_tap ? L.f[LF_TAP_SLOT] : max_tap;
    m->layers.push_back(L);
  }
  m->n_bufs = max_buf + 1;
  m->n_masks = max_mask + 1;
  m->n_taps = max_tap + 1;
  if (m->n_taps != head->n_taps) { delete m; return fail("head n_taps does not match the taps in the layer plan"); }
  m->buf_channels.assign(m->n_bufs, 64);
  m->tap_mask_slot.assign(m->n_taps > 0 ? m->n_taps : 1, 0);
  for (const Layer& L : m->layers) {
    int& ci = m->buf_channels[L.f[LF_IN_BUF]];
    ci = L.f[LF_CIN] > ci ? L.f[LF_CIN] : ci;
    if (L.f[LF_OUT_BUF] >= 0) { int& co = m->buf_channels[L.f[LF_OUT_BUF]]; co = L.f[LF_COUT] > co ? L.f[LF_COUT] : co; }
    if (L.f[LF_KIND] != 1) continue;
    if (L.f[LF_TAP_MODE] != 0) {
      m->tap_mask_slot[L.f[LF_TAP_SLOT]] = L.f[LF_MASK_OUT];
      if (m->tap_width != 0 && m->tap_width != L.f[LF_COUT]) { delete m; return fail("NMD taps of different widths are not supported"); }
      m->tap_width = L.f[LF_COUT];
    }
    if (L.f[LF_POOL_MODE] != 0) m->final_mask = L.f[LF_MASK_OUT];
  }
  // heads
  const int feat = head->feat_dim, ncls = head->n_classes;
  if (upload_f32(head->cls_w, static_cast<size_t>(feat) * ncls, &m->cls_w)) return 2;
  if (upload_f32(head->cls_b, ncls, &m->cls_b)) return 2;
  if (head->mlp_hidden > 0) {
    if (head->mlp_hidden != feat || feat > 256) { delete m; return fail("MLP head needs hidden width == feature width <= 256"); }
    if (upload_f32(head->mlp_w1, static_cast<size_t>(feat) * feat, &m->mlp_w1)) return 2;
    if (upload_f32(head->mlp_b1, feat, &m->mlp_b1)) return 2;
    if (head->mlp_w2) {
      if (upload_f32(head->mlp_w2, static_cast<size_t>(feat) * feat, &m->mlp_w2)) return 2;
      if (upload_f32(head->mlp_b2, feat, &m->mlp_b2)) return 2;
    }
  }
  if (m->n_taps > 0) {
    const int nmd_dim = m->n_taps * m->tap_width;
    std::vector<float> means(static_cast<size_t>(nmd_dim), 0.0f);
    for (int l = 0; l < n_layers; ++l)
      if (layers[l].i[LF_TAP_MODE] != 0 && layers[l].p[LP_TAP_MEAN])
        std::memcpy(means.data() + static_cast<size_t>(layers[l].i[LF_TAP_SLOT]) * m->tap_width, layers[l].p[LP_TAP_MEAN], m->tap_width * 4);
    if (upload_f32(means.data(), means.size(), &m->tap_mean)) return 2;
    if (head->rel_w1) {
      const int n_sig = head->reserved[0] & 7;       // nmd_plus_signals: the signals follow the NMD vector
      if (n_sig > 5 || (n_sig && head->n_classes > jg::kMaxSignalClasses)) return fail("jg_model_create: at most 5 OOD signals over at most 16 classes");
      if (upload_f32(head->rel_w1, static_cast<size_t>(nmd_dim + n_sig) * head->rel_hidden, &m->rel_w1)) return 2;
      if (upload_f32(head->rel_b1, head->rel_hidden, &m->rel_b1)) return 2;
      if (upload_f32(head->rel_w2, head->rel_hidden, &m->rel_w2)) return 2;
      if (upload_f32(head->rel_b2, 1, &m->rel_b2)) return 2;
    }
  }
  JG_CUDA(cudaMalloc(&m->err, 4));
  JG_CUDA(cudaMemset(m->err, 0, 4));
  if (finish_resident(m, head)) { delete m; return 2; }
  if (const char* impl = std::getenv("JG_CONV_IMPL")) m->conv_impl = std::atoi(impl);
  m->prof_ms.assign(m->layers.size(), 0.0);
  m->prof_launches.assign(m->layers.size(), 0);
  m->prof_rows.assign(m->layers.size(), 0.0);
  *out = m;
  return 0;
}

int jg_model_destroy(jg_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->ctx->device);
  free_workspace(m);
  for (Layer& L : m->layers) {
    cudaFree(L.w); cudaFree(L.w2); cudaFree(L.w3); cudaFree(L.w_tap); cudaFree(L.par); cudaFree(L.shifts);
    free_weight_images(L.odd);
    for (auto& v : L.slices) for (auto& wi : v) free_weight_images(wi);
  }
  for (float* p : {m->cls_w, m->cls_b, m->rel_w1, m->rel_b1, m->rel_w2, m->rel_b2, m->tap_mean, m->mlp_w1, m->mlp_b1,
                   m->mlp_w2, m->mlp_b2}) cudaFree(p);
  cudaFree(m->err);
  cudaFree(m->rs_block);
  delete m;
  return 0;
}

int64_t jg_model_max_windows(jg_model* m, int32_t lc, int64_t workspace_bytes) {
  int period, rpw;
  model_geometry(m, lc, &period, &rpw);
  const long long per = bytes_per_window(m, rpw);
  return per > 0 ? workspace_bytes / per : 0;
}
int64_t jg_model_workspace_bytes(jg_model* m) { return m->ws_bytes; }

double jg_model_flops_per_window(jg_model* m, int32_t lc) {
  double f = 0.0;
  for (const Layer& L : m->layers) {
    if (L.f[LF_KIND] != 1) continue;
    const int l_out = layer_len(L, lc) - L.f[LF_SHRINK];
    f += 2.0 * m->frames * l_out * L.f[LF_K] * static_cast<double>(L.f[LF_CIN]) * L.f[LF_COUT];
  }
  return f;
}

int jg_model_forward(jg_ctx* ctx, jg_model* m, const uint8_t* d_tokens, const int32_t* d_lpad, int64_t n_windows,
                     int32_t lc, int32_t pitch, float* d_logits, float* d_rel, float* d_emb, float* d_nmd,
                     int32_t use_ref) {
  if (n_windows <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  int period, rpw;
  model_geometry(m, lc, &period, &rpw);
  const long long rows = n_windows * rpw;
  if (rows / jg::kTileM > 0x7FFFFFFFLL) return fail("too many windows in one forward call");
  if (ensure_workspace(m, n_windows, rows)) return 2;
  cudaStream_t st = ctx->stream;
  const long long plane = m->cap_rows + 2 * jg::kGuardRows;
  const jg::RowGeom geom{rpw, period, m->frames};
  // rows [rows, rows + guard) may hold stale data of an earlier, larger call
  if (rows < m->cap_rows) {
    for (int b = 0; b < m->n_bufs; ++b)
      for (int g = 0; g < m->buf_channels[b] / 64; ++g)
        JG_CUDA(cudaMemsetAsync(m->bufs[b] + (static_cast<long long>(g) * plane + jg::kGuardRows + rows) * 64, 0, jg::kGuardRows * 128, st));
    for (int s = 0; s < m->n_masks; ++s) JG_CUDA(cudaMemsetAsync(m->masks[s] + jg::kGuardRows + rows, 0, jg::kGuardRows, st));
  }
  JG_CUDA(cudaMemsetAsync(m->counts, 0, static_cast<size_t>(m->n_masks) * m->cap_windows * 4, st));
  if (m->n_taps > 0) JG_CUDA(cudaMemsetAsync(m->tap_sum, 0, static_cast<size_t>(m->n_taps) * n_windows * m->tap_width * 4, st));
  const long long pool_n = n_windows * m->head.feat_dim;
  jg::fill_f32_kernel<<<grid_for(pool_n, 256, ctx->num_sms), 256, 0, st>>>(m->pool, pool_n, m->head.pool_mode == 1 ? -1.0e9f : 0.0f);
  ctx->launches++;

  auto buf_row0 = [&](int b) { return m->bufs[b] + static_cast<long long>(jg::kGuardRows) * 64; };
  auto mask_row0 = [&](int s) { return m->masks[s] + jg::kGuardRows; };

  // Narrow stacks: ONE launch keeps every window in shared memory through all conv layers (conv_resident.cuh); it reads the
  // tokens and writes the pooled features + the valid-row counts of the final mask, nothing else.
  m->last_resident = false;
  if (m->rs_ok && !use_ref && rpw / jg::kTileM <= jg::rs::kMaxTilesRs && m->frames * pitch <= 128 * jg::rs::kTokWordsRs && pitch % 4 == 0) {
    const jg::rs::SmemRs S = jg::rs::smem_rs(rpw, m->rs_par.n_layers, m->rs_par.w_bytes, m->frames, pitch, m->rs_stem_span);
    if (S.total <= jg::kMaxSmem && 2u * S.slot_bytes <= S.buf_bytes) {
      jg::rs::ResidentParams rp = m->rs_par;
      rp.tokens = d_tokens; rp.lpad = d_lpad; rp.n_windows = n_windows; rp.lc = lc; rp.pitch = pitch; rp.tok_offset = m->tok_offset;
      rp.period = period; rp.frames = m->frames; rp.rpw = rpw;
      rp.wblock = m->rs_block;
      rp.pool = m->pool; rp.pool_pitch = m->head.feat_dim;
      rp.count = m->counts; rp.cap_windows = m->cap_windows;
      rp.tap_sum = m->tap_sum; rp.tap_width = m->tap_width;
      rp.err = m->err;
      jg::rs::fill_layer_offsets(rp, S);
      long long* d_trace = nullptr;
      if (std::getenv("JG_RS_TRACE")) {          // probe only: per-tile clock64 timeline of CTA 0's fourth window, printed to stderr
        JG_CUDA(cudaMalloc(&d_trace, 2048 * 8));
        JG_CUDA(cudaMemsetAsync(d_trace, 0, 2048 * 8, st));
      }
      rp.dbg = d_trace;
      JG_CUDA(cudaFuncSetAttribute(jg::rs::stack_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S.total)));
      cudaEvent_t ev0 = nullptr, ev1 = nullptr;
      if (m->profiling) {
        JG_CUDA(cudaEventCreate(&ev0));
        JG_CUDA(cudaEventCreate(&ev1));
        JG_CUDA(cudaEventRecord(ev0, st));
      }
      const int grid = n_windows < ctx->num_sms ? static_cast<int>(n_windows) : ctx->num_sms;
      jg::rs::stack_resident_kernel<<<grid, jg::rs::kThreadsRs, S.total, st>>>(rp);
      ctx->launches++;
      JG_CUDA(cudaGetLastError());
      if (d_trace) {
        std::vector<long long> h(2048);
        JG_CUDA(cudaStreamSynchronize(st));
        JG_CUDA(cudaMemcpy(h.data(), d_trace, 2048 * 8, cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        long long t0 = h[0];
        for (int i = 0; i < m->rs_par.n_layers * 32; ++i) if (h[i] > 0 && h[i] < t0) t0 = h[i];
        std::fprintf(stderr, "resident trace (cycles from the window's first event; per tile: mma wait-done, mma issued, epi acc-ready, epi done)\n");
        for (int l = 0; l < m->rs_par.n_layers; ++l) {
          std::fprintf(stderr, "L%d:", l);
          for (int i = 0; i < rpw / jg::kTileM; ++i)
            std::fprintf(stderr, "  t%d[%lld %lld | %lld %lld]", i, h[l * 32 + i * 4] - t0, h[l * 32 + i * 4 + 1] - t0, h[l * 32 + i * 4 + 2] - t0, h[l * 32 + i * 4 + 3] - t0);
          std::fprintf(stderr, "\n");
        }
      }
      {   // a stem tap on the raw conv output is linear in the one-hot input: taken from token counts, like on the per-layer path
        Layer& L0 = m->layers[0];
        if (L0.w_tap != nullptr && L0.f[LF_TAP_MODE] == 1) {
          const int cout0 = L0.f[LF_COUT];
          jg::StemTapParams tp{};
          tp.tokens = d_tokens; tp.lpad = d_lpad; tp.count = m->counts + static_cast<long long>(L0.f[LF_MASK_OUT]) * m->cap_windows;
          tp.w = L0.w_tap; tp.wsum = L0.w_tap + static_cast<size_t>(L0.f[LF_K]) * 64 * cout0; tp.bias = L0.par;
          tp.tap = m->tap_sum + static_cast<long long>(L0.f[LF_TAP_SLOT]) * n_windows * m->tap_width;
          tp.n_windows = n_windows; tp.lc = lc; tp.pitch = pitch; tp.frames = m->frames; tp.tok_offset = m->tok_offset;
          tp.ntaps = L0.f[LF_K]; tp.shrink = L0.f[LF_SHRINK]; tp.cout = cout0;
          for (int t = 0; t < L0.f[LF_K]; ++t) tp.shifts[t] = L0.shifts_h[t];
          jg::stem_tap_kernel<<<grid_for(n_windows, 1, ctx->num_sms, 16), 128, 0, st>>>(tp);
          ctx->launches++;
          JG_CUDA(cudaGetLastError());
        }
      }
      if (m->profiling) {     // the one launch is booked on the last conv layer; every layer gets the windows it covered
        JG_CUDA(cudaEventRecord(ev1, st));
        m->prof_events.emplace_back(ev0, ev1);
        m->prof_layer.push_back(static_cast<int>(m->layers.size()) - 1);
        for (size_t l = 0; l < m->layers.size(); ++l) m->prof_rows[l] += static_cast<double>(n_windows);
      }
      m->last_resident = true;
    }
  }
  bool pool_final = false;
  if (!m->last_resident) {
  // stem operand: one-hot rows + token mask
  jg::expand_tokens_kernel<<<grid_for(n_windows * (m->frames + 1), 1, ctx->num_sms, 8), 256, 0, st>>>(
      d_tokens, d_lpad, n_windows, lc, pitch, geom, m->tok_offset, buf_row0(m->layers[0].f[LF_IN_BUF]), mask_row0(m->layers[0].f[LF_MASK_IN]),
      m->counts + static_cast<long long>(m->layers[0].f[LF_MASK_IN]) * m->cap_windows);
  ctx->launches++;
  JG_CUDA(cudaGetLastError());

  for (Layer& L : m->layers) {
    const int cout = L.f[LF_COUT];
    if (L.f[LF_KIND] == 2) {
      jg::maxpool2_kernel<<<grid_for(rows * 8 * (L.f[LF_CIN] / 64), 256, ctx->num_sms, 16), 256, 0, st>>>(
          buf_row0(L.f[LF_IN_BUF]), d_lpad, rows, geom, L.f[LF_CUM_SHRINK_IN], L.f[LF_HALVINGS], L.f[LF_CIN] / 64, plane,
          buf_row0(L.f[LF_OUT_BUF]), mask_row0(L.f[LF_MASK_OUT]), m->counts + static_cast<long long>(L.f[LF_MASK_OUT]) * m->cap_windows);
      ctx->launches++;
      continue;
    }
    if (L.f[LF_KIND] == 3) {
      const int threads = (L.f[LF_CIN] / 8) * 8;
      jg::framesum_globalmax_kernel<<<grid_for(n_windows, 1, ctx->num_sms, 8), threads, static_cast<size_t>(8) * L.f[LF_CIN] * 4, st>>>(
          buf_row0(L.f[LF_IN_BUF]), d_lpad, static_cast<int>(n_windows), geom, L.f[LF_CUM_SHRINK_IN], L.f[LF_HALVINGS],
          L.f[LF_CIN], plane, m->pool);
      ctx->launches++;
      pool_final = true;
      continue;
    }
    if (L.f[LF_KIND] == 4) {     // rows -> (even, odd) row planes in front of the strided convs of a residual block
      const int groups = L.f[LF_CIN] / 64;
      jg::rows_to_phases_kernel<<<grid_for(rows * 16 * groups, 256, ctx->num_sms, 16), 256, 0, st>>>(
          buf_row0(L.f[LF_IN_BUF]), d_lpad, rows, geom, L.f[LF_CUM_SHRINK_IN], L.f[LF_HALVINGS], len_round_of(L), groups, plane,
          buf_row0(L.f[LF_OUT_BUF]));
      ctx->launches++;
      JG_CUDA(cudaGetLastError());
      continue;
    }
    // Weight images: a strided conv (plan.py:split_phases) has one kernel per parity of its pre-stride frame length, which is
    // uniform over a forward call except in the padded short pass -- there the plan compiler's caller refuses strided models.
    int par = 0;
    if (L.odd.w != nullptr || !L.slices[1].empty()) {
      const int l_prev = (lc - L.f[LF_CUM_SHRINK_IN] + (1 << (L.f[LF_HALVINGS] - 1)) - 1) >> (L.f[LF_HALVINGS] - 1);
      par = l_prev & 1;
    }
    const jg::act_t* img_w = par ? L.odd.w : L.w;
    const jg::act_t* img_w2 = par ? L.odd.w2 : L.w2;
    const jg::act_t* img_w3 = par ? L.odd.w3 : L.w3;
    const int n_slices = L.slice_width ? cout / L.slice_width : 1;
    const int lcout = L.slice_width ? L.slice_width : cout;       // output channels of one launch
    // A layer whose weights fit no tensor-core kernel's shared memory even in slices runs on the CUDA-core kernel, like the
    // whole model does under use_ref.
    jg::ConvParams geo{};
    geo.cin = L.f[LF_CIN]; geo.cout = lcout; geo.ntaps = L.f[LF_K]; geo.halo_l = L.halo_l; geo.halo_r = L.halo_r;
    geo.n_tiles = static_cast<int>(rows / jg::kTileM);
    const bool fits_tc = jg::conv_tc_stages(geo) >= 0, fits_tc2 = jg::conv_tc2_eligible(geo);
    const bool layer_ref = use_ref || (!fits_tc && !fits_tc2);
    if (layer_ref) {   // the CUDA-core path keeps the stand-alone mask kernel
      jg::propagate_mask_kernel<<<grid_for(rows, 256, ctx->num_sms, 16), 256, 0, st>>>(
          mask_row0(L.f[LF_MASK_IN]), d_lpad, rows, geom, L.f[LF_CUM_SHRINK_IN], L.f[LF_HALVINGS], len_round_of(L), L.f[LF_SHRINK], L.f[LF_K],
          L.shifts, L.f[LF_MASKING], L.f[LF_MASK_THR], mask_row0(L.f[LF_MASK_OUT]), m->counts + static_cast<long long>(L.f[LF_MASK_OUT]) * m->cap_windows);
      ctx->launches++;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (m->profiling) {
      JG_CUDA(cudaEventCreate(&ev0));
      JG_CUDA(cudaEventCreate(&ev1));
      JG_CUDA(cudaEventRecord(ev0, st));
    }
    bool linear_tap = false;
    jg::ConvParams p{};
    for (int sl = 0; sl < n_slices; ++sl) {
      const int c0 = sl * lcout;                                  // first output channel of this launch
      const long long g0 = static_cast<long long>(c0 / 64) * plane * 64;      // its channel-group offset in a g64sw tensor
      p = jg::ConvParams{};
      p.x = buf_row0(L.f[LF_IN_BUF]);
      p.y = L.f[LF_OUT_BUF] >= 0 ? buf_row0(L.f[LF_OUT_BUF]) + g0 : nullptr;
      p.sc = L.f[LF_SC_BUF] >= 0 ? buf_row0(L.f[LF_SC_BUF]) + g0 : nullptr;
      p.sc_mask = (L.f[LF_SC_BUF] >= 0 && L.f[LF_SC_MASK] >= 0) ? mask_row0(L.f[LF_SC_MASK]) : nullptr;
      p.out_mask = mask_row0(L.f[LF_MASK_OUT]);
      p.w = L.slice_width ? L.slices[par][sl].w : img_w;
      p.bias = L.par + c0;
      p.scale1 = L.par + cout + c0;
      p.shift1 = L.par + 2 * cout + c0;
      p.scale2 = L.par + 3 * cout + c0;
      p.shift2 = L.par + 4 * cout + c0;
      p.sc_const = L.par + 5 * cout + c0;
      p.dyt_g1 = L.par + 6 * cout + c0; p.dyt_b1 = L.par + 7 * cout + c0; p.dyt_g2 = L.par + 8 * cout + c0; p.dyt_b2 = L.par + 9 * cout + c0;
      p.dyt1 = L.f[LF_DYT1]; p.dyt2 = L.f[LF_DYT2];
      p.tap_sum = L.f[LF_TAP_MODE] != 0 ? m->tap_sum + static_cast<long long>(L.f[LF_TAP_SLOT]) * n_windows * m->tap_width + c0 : nullptr;
      p.pool = L.f[LF_POOL_MODE] != 0 ? m->pool + c0 : nullptr;
      p.red_pitch = L.slice_width ? cout : 0;
      p.x_plane = plane;
      p.y_plane = plane;
      p.n_tiles = static_cast<int>(rows / jg::kTileM);
      p.rows_per_window = rpw;
      p.cin = L.f[LF_CIN];
      p.cout = lcout;
      p.ntaps = L.f[LF_K];
      for (int t = 0; t < L.f[LF_K]; ++t) p.shifts[t] = L.shifts_h[t];
      p.halo_l = L.halo_l;
      p.halo_r = L.halo_r;
      p.act1 = L.f[LF_ACT1];
      p.act2 = L.f[LF_ACT2];
      p.has_affine2 = L.f[LF_HAS_AFF2];
      p.tap_mode = L.f[LF_TAP_MODE];
      // The stem's tap is linear in its one-hot input: it is taken from token counts after the conv
      // (stem_tap_kernel), which leaves the stem a light layer for the tensor-core kernels.
      linear_tap = L.w_tap != nullptr && L.f[LF_SHRINK] >= 0 && L.f[LF_CUM_SHRINK_IN] == 0 && L.f[LF_HALVINGS] == 0;
      if (linear_tap) { p.tap_mode = 0; p.tap_sum = nullptr; }
      p.pool_mode = L.f[LF_POOL_MODE];
      // the first slice (or the only launch) derives and publishes the layer's row mask / window counts; further slices read it
      p.fuse_mask = (layer_ref || sl > 0) ? 0 : 1;
      p.folded = L.folded ? 1 : 0;
      p.epi_f32 = L.f[LF_EPI_F32];
      p.ln1 = L.f[LF_LN1];
      p.mask_thr = L.f[LF_MASK_THR];
      if (p.ln1) {
        std::memcpy(&p.ln_eps, &L.f[LF_LN_EPS], 4);
        p.ln_inv_c = 1.0f / static_cast<float>(L.f[LF_REAL_COUT] > 0 ? L.f[LF_REAL_COUT] : cout);
        if (L.slice_width || (!layer_ref && !fits_tc)) return fail("MaskedLayerNormalization needs a layer that fits the single-CTA tensor-core kernel");
      }
      p.in_mask = mask_row0(L.f[LF_MASK_IN]);
      p.out_mask_w = mask_row0(L.f[LF_MASK_OUT]);
      p.lpad = d_lpad;
      p.count = m->counts + static_cast<long long>(L.f[LF_MASK_OUT]) * m->cap_windows;
      p.masking = L.f[LF_MASKING];
      p.period = period;
      p.frames = m->frames;
      p.shrink_in = L.f[LF_CUM_SHRINK_IN];
      p.halvings = L.f[LF_HALVINGS];
      p.len_round = len_round_of(L);
      p.shrink = L.f[LF_SHRINK];
      p.err = m->err;
      p.dbg = nullptr;
      // Kernel choice (profiles/conv_kernel_r1.md): light epilogues -> CTA-pair kernel with two epilogue
      // groups and staged bulk stores; layers with an NMD tap / second affine / pool are bound by epilogue
      // instruction issue -> three epilogue groups: the pair kernel's 3-group variant when Cin >= 128
      // (+13 % over the single-CTA kernel in the forward pass), else the single-CTA kernel (stem).
      const bool heavy = p.tap_mode != 0 || p.has_affine2 != 0 || p.pool_mode != 0;
      const bool pair3 = !layer_ref && heavy && p.cin >= 128 && (m->conv_impl == 0 || m->conv_impl == 3) &&
                         jg::conv_tc2_eligible(p, 3);                                       // pair kernel, 3 epilogue groups
      const bool pair = !p.ln1 && (pair3 || (!layer_ref && fits_tc2 && (!fits_tc || (m->conv_impl != 1 && (m->conv_impl == 2 || !heavy)))));
      if (pair) p.w = L.slice_width ? L.slices[par][sl].w2 : img_w2;
      // 128-output-channel layers with one of the specialised epilogue shapes: the weights-stationary kernel (weights resident in
      // tensor memory, transposed accumulators; profiles/conv_kernel_r2.md).  JG_CONV_IMPL=1|2|3 keeps the round-1 kernels.
      const bool ws = !layer_ref && !L.slice_width && img_w3 != nullptr && (m->conv_impl == 0 || m->conv_impl == 4) && jg::conv_ws_mode(p) >= 0;
      if (ws) p.w = img_w3;
      L.last_kernel = layer_ref ? "jg::conv_ref_kernel"
                    : ws ? "jg::ws::conv_ws_kernel"
                    : (pair3 ? "jg::tc2::conv_tc2_kernel<3>" : (pair ? "jg::tc2::conv_tc2_kernel<2>" : "jg::tc::conv_tc_kernel"));
      cudaError_t e = layer_ref ? jg::launch_conv_ref(p, st)
                    : ws ? jg::launch_conv_ws(p, ctx->num_sms, st)
                    : (pair ? jg::launch_conv_tc2(p, ctx->num_sms, st, pair3 ? 3 : 2) : jg::launch_conv_tc(p, ctx->num_sms, st));
      ctx->launches++;
      if (e != cudaSuccess) return cuda_fail(e, "conv launch");
    }
    if (linear_tap) {
      jg::StemTapParams tp{};
      tp.tokens = d_tokens; tp.lpad = d_lpad; tp.count = p.count;
      tp.w = L.w_tap; tp.wsum = L.w_tap + static_cast<size_t>(L.f[LF_K]) * 64 * cout; tp.bias = L.par;
      tp.tap = m->tap_sum + static_cast<long long>(L.f[LF_TAP_SLOT]) * n_windows * m->tap_width;
      tp.n_windows = n_windows; tp.lc = lc; tp.pitch = pitch; tp.frames = m->frames; tp.tok_offset = m->tok_offset;
      tp.ntaps = L.f[LF_K]; tp.shrink = L.f[LF_SHRINK]; tp.cout = cout;
      for (int t = 0; t < L.f[LF_K]; ++t) tp.shifts[t] = L.shifts_h[t];
      jg::stem_tap_kernel<<<grid_for(n_windows, 1, ctx->num_sms, 16), 128, 0, st>>>(tp);
      ctx->launches++;
      JG_CUDA(cudaGetLastError());
    }
    if (m->profiling) {
      JG_CUDA(cudaEventRecord(ev1, st));
      m->prof_events.emplace_back(ev0, ev1);
      m->prof_layer.push_back(static_cast<int>(&L - m->layers.data()));
      m->prof_rows[&L - m->layers.data()] += static_cast<double>(n_windows);
    }
  }
  }   // per-layer path

  jg::HeadParams hp{};
  hp.pool = m->pool;
  hp.pool_count = m->counts + static_cast<long long>(m->final_mask) * m->cap_windows;
  hp.tap_sum = m->tap_sum;
  hp.tap_count = m->tap_count_ptrs;
  hp.tap_mean = m->tap_mean;
  hp.cls_w = m->cls_w; hp.cls_b = m->cls_b;
  hp.rel_w1 = m->rel_w1; hp.rel_b1 = m->rel_b1; hp.rel_w2 = m->rel_w2; hp.rel_b2 = m->rel_b2;
  hp.logits = d_logits;
  hp.rel = m->rel_w1 ? d_rel : nullptr;
  hp.emb = d_emb;
  hp.nmd = d_nmd;
  hp.n_windows = static_cast<int>(n_windows);
  hp.feat = m->head.feat_dim;
  hp.n_classes = m->head.n_classes;
  hp.pool_mode = m->head.pool_mode;
  hp.n_taps = m->n_taps;
  hp.tap_width = m->tap_width;
  hp.rel_hidden = m->head.rel_hidden;
  hp.masking = m->layers.back().f[LF_MASKING];
  hp.mlp_w1 = m->mlp_w1; hp.mlp_b1 = m->mlp_b1; hp.mlp_w2 = m->mlp_w2; hp.mlp_b2 = m->mlp_b2;
  hp.mlp_hidden = m->head.mlp_hidden;
  hp.mlp_act = m->head.mlp_act;
  hp.pool_final = pool_final ? 1 : 0;
  hp.signals = m->head.reserved[0];
  const int warps = 4;
  const size_t smem = static_cast<size_t>(warps) * (hp.feat + hp.n_taps * hp.tap_width) * 4;
  jg::heads_kernel<<<grid_for(n_windows, warps, ctx->num_sms, 16), warps * 32, smem, st>>>(hp);
  ctx->launches++;
  JG_CUDA(cudaGetLastError());
  return 0;
}

int jg_model_kernel_names(jg_model* m, char* buf, int32_t n) {
  if (!buf || n <= 0) return fail("jg_model_kernel_names: no buffer");
  std::vector<std::pair<std::string, int>> seen;
  if (m->last_resident) {
    std::snprintf(buf, static_cast<size_t>(n), "1 x jg::rs::stack_resident_kernel (%d conv layers)", static_cast<int>(m->layers.size()));
    return 0;
  }
  for (const Layer& L : m->layers) {
    if (L.f[LF_KIND] != 1 || !L.last_kernel[0]) continue;
    bool found = false;
    for (auto& pr : seen) if (pr.first == L.last_kernel) { pr.second++; found = true; }
    if (!found) seen.emplace_back(L.last_kernel, 1);
  }
  std::string out;
  for (auto& pr : seen) out += (out.empty() ? "" : " + ") + std::to_string(pr.second) + " x " + pr.first;
  std::snprintf(buf, static_cast<size_t>(n), "%s", out.c_str());
  return 0;
}

int jg_model_set_profiling(jg_model* m, int32_t on) {
  m->profiling = on != 0;
  m->prof_ms.assign(m->layers.size(), 0.0);
  m->prof_launches.assign(m->layers.size(), 0);
  m->prof_rows.assign(m->layers.size(), 0.0);
  for (auto& pr : m->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  m->prof_events.clear();
  m->prof_layer.clear();
  return 0;
}

int jg_model_get_profile(jg_model* m, int32_t n_layers, double* ms, int64_t* launches, double* windows) {
  JG_CUDA(cudaStreamSynchronize(m->ctx->stream));
  for (size_t i = 0; i < m->prof_events.size(); ++i) {
    float t = 0.0f;
    JG_CUDA(cudaEventElapsedTime(&t, m->prof_events[i].first, m->prof_events[i].second));
    m->prof_ms[m->prof_layer[i]] += t;
    m->prof_launches[m->prof_layer[i]] += 1;
    cudaEventDestroy(m->prof_events[i].first);
    cudaEventDestroy(m->prof_events[i].second);
  }
  m->prof_events.clear();
  m->prof_layer.clear();
  for (int l = 0; l < n_layers && l < static_cast<int>(m->layers.size()); ++l) {
    ms[l] = m->prof_ms[l];
    launches[l] = m->prof_launches[l];
    windows[l] = m->prof_rows[l];
  }
  return 0;
}

// ---- stage 4 -----------------------------------------------------------------------------------
int jg_aggregate_contigs(jg_ctx* ctx, const float* d_logits, const float* d_rel, const int64_t* d_offsets,
                         int64_t n_contigs, int64_t n_windows, int32_t n_cls, uint16_t* d_mean_h, uint16_t* d_var_h,
                         int32_t* d_consensus, int32_t* d_counts, uint16_t* d_entropy_h, uint16_t* d_energy_h,
                         int32_t* d_rel_pos, int32_t* d_frag_pred) {
  if (n_contigs <= 0) return 0;
  if (n_cls < 1 || n_cls > 32) return fail("n_cls must be in [1, 32]");
  JG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (n_windows <= 0) return 0;
  float* ent = nullptr;
  double* en = nullptr;
  JG_CUDA(cudaMallocAsync(&ent, n_windows * 4, st));
  JG_CUDA(cudaMallocAsync(&en, n_windows * 8, st));
  const long long* off = reinterpret_cast<const long long*>(d_offsets);
  jg::window_scalars_kernel<<<grid_for(n_windows, 256, ctx->num_sms), 256, 0, st>>>(d_logits, n_windows, n_cls, d_frag_pred, ent, en);
  jg::contig_moments_kernel<<<grid_for(n_contigs * n_cls, 128, ctx->num_sms, 16), 128, 0, st>>>(
      d_logits, off, n_contigs, n_cls, reinterpret_cast<__half*>(d_mean_h), reinterpret_cast<__half*>(d_var_h));
  jg::contig_summary_kernel<<<grid_for(n_contigs * 32, 256, ctx->num_sms, 16), 256, 0, st>>>(
      reinterpret_cast<const __half*>(d_mean_h), d_frag_pred, ent, en, d_rel, off, n_contigs, n_cls, d_consensus, d_counts,
      reinterpret_cast<__half*>(d_entropy_h), reinterpret_cast<__half*>(d_energy_h), d_rel_pos);
  ctx->launches += 3;
  JG_CUDA(cudaGetLastError());
  JG_CUDA(cudaFreeAsync(ent, st));
  JG_CUDA(cudaFreeAsync(en, st));
  return 0;
}

int jg_smooth_scores(jg_ctx* ctx, const float* d_logits, const int64_t* d_offsets, int64_t n_contigs,
                     int32_t n_cls, int32_t box, double* d_out) {
  if (n_contigs <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  dim3 grid(ctx->num_sms, static_cast<unsigned>(n_contigs < 64 ? n_contigs : 64));
  jg::smooth_scores_kernel<<<grid, 128, 0, ctx->stream>>>(d_logits, reinterpret_cast<const long long*>(d_offsets), n_contigs, n_cls, box, d_out);
  ctx->launches++;
  JG_CUDA(cudaGetLastError());
  return 0;
}

namespace {
int segment_launch(jg_ctx* ctx, const double* d_signal, const long long* d_offsets, int32_t n_single, int64_t total_points,
                   int32_t n_contigs, int32_t min_size, int32_t n_pen, int32_t* d_bkps, int32_t* d_nbkps) {
  JG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  double* work = nullptr;
  const size_t slots = static_cast<size_t>(total_points) + static_cast<size_t>(n_contigs);
  const size_t bytes = (2 + static_cast<size_t>(n_pen)) * slots * 8 + static_cast<size_t>(n_pen) * slots * 4 + 16;
  JG_CUDA(cudaMallocAsync(&work, bytes, st));
  jg::prefix_sums_kernel<<<n_contigs, 32, 0, st>>>(d_signal, d_offsets, n_single, static_cast<long long>(slots), work);
  jg::segment_scores_kernel<<<dim3(n_pen, n_contigs), 256, 32 * 8 + 32 * 4, st>>>(d_offsets, n_single, total_points,
                                                                                static_cast<long long>(slots), min_size, n_pen,
                                                                                work, d_bkps, d_nbkps);
  ctx->launches += 2;
  JG_CUDA(cudaGetLastError());
  JG_CUDA(cudaFreeAsync(work, st));
  return 0;
}
}  // namespace

int jg_segment_scores(jg_ctx* ctx, const double* d_signal, int32_t n, int32_t min_size, int32_t n_pen,
                      int32_t* d_bkps, int32_t* d_nbkps) {
  if (n <= 0 || n_pen <= 0) return 0;
  return segment_launch(ctx, d_signal, nullptr, n, n, 1, min_size, n_pen, d_bkps, d_nbkps);
}

int jg_segment_scores_batched(jg_ctx* ctx, const double* d_signal, const int64_t* d_offsets, int32_t n_contigs,
                              int64_t total_points, int32_t min_size, int32_t n_pen, int32_t* d_bkps, int32_t* d_nbkps) {
  if (n_contigs <= 0 || total_points <= 0 || n_pen <= 0) return 0;
  if (n_contigs > 65535) return fail("jg_segment_scores_batched: at most 65535 contigs per call");
  return segment_launch(ctx, d_signal, reinterpret_cast<const long long*>(d_offsets), 0, total_points, n_contigs, min_size,
                        n_pen, d_bkps, d_nbkps);
}

int jg_viterbi_decode(jg_ctx* ctx, const float* d_logits, const int64_t* d_offsets, int32_t n_contigs, int64_t n_windows,
                      int32_t n_cls, const double* d_costs, int32_t* d_path, int32_t* d_counts) {
  if (n_cls < 1 || n_cls > jg::kMaxCrfClasses) return fail("jg_viterbi_decode: n_cls must be in [1, 8]");
  if (n_contigs <= 0 || n_windows <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint8_t* bp = nullptr;
  JG_CUDA(cudaMallocAsync(&bp, static_cast<size_t>(n_windows) * n_cls, st));
  jg::viterbi_kernel<<<grid_for(n_contigs, 64, ctx->num_sms, 8), 64, 0, st>>>(
      d_logits, reinterpret_cast<const long long*>(d_offsets), n_contigs, n_cls, d_costs, bp, d_path, d_counts);
  ctx->launches += 1;
  JG_CUDA(cudaGetLastError());
  JG_CUDA(cudaFreeAsync(bp, st));
  return 0;
}

int jg_refine_contigs(jg_ctx* ctx, const float* d_logits, const int64_t* d_offsets, int64_t n_contigs, int64_t n_windows,
                      int32_t n_cls, const double* d_tau, int32_t merge_bp, int32_t merge_pv, int32_t mode, double merge_share,
                      uint8_t* d_label, double* d_margin, double* d_sums, int32_t* d_stats, double* d_total_weight) {
  if (n_cls < jg::kRefineClasses) return fail("jg_refine_contigs: the refinement layer needs the 6 score columns");
  if (mode < 0 || mode > 2) return fail("jg_refine_contigs: mode must be 0 (gated), 1 (weighted) or 2 (unweighted)");
  if (n_contigs <= 0 || n_windows <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  jg::refine_windows_kernel<<<grid_for(n_windows, 256, ctx->num_sms), 256, 0, st>>>(d_logits, n_windows, n_cls, d_tau, merge_bp,
                                                                                   merge_pv, d_label, d_margin);
  jg::refine_contigs_kernel<<<grid_for(n_contigs * 32, 256, ctx->num_sms, 16), 256, 0, st>>>(
      d_logits, n_cls, d_label, d_margin, reinterpret_cast<const long long*>(d_offsets), n_contigs, mode, merge_share, d_sums,
      d_stats, d_total_weight);
  ctx->launches += 2;
  JG_CUDA(cudaGetLastError());
  return 0;
}

int jg_legacy_reliability(jg_ctx* ctx, const float* d_embedding, int64_t n_windows, int32_t dim, const float* d_batch_mean,
                          const float* d_batch_std, const double* d_coef, double intercept, double cal_a, double cal_b,
                          const int64_t* d_offsets, int32_t n_contigs, double* d_window_p0, double* d_contig_mean) {
  if (n_windows <= 0) return 0;
  JG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  jg::legacy_reliability_kernel<<<grid_for(n_windows * 32, 256, ctx->num_sms, 8), 256, 0, st>>>(
      d_embedding, n_windows, dim, d_batch_mean, d_batch_std, d_coef, intercept, cal_a, cal_b, d_window_p0);
  ctx->launches += 1;
  if (n_contigs > 0 && d_contig_mean) {
    jg::segment_mean_f64_kernel<<<grid_for(n_contigs, 128, ctx->num_sms, 8), 128, 0, st>>>(
        d_window_p0, reinterpret_cast<const long long*>(d_offsets), n_contigs, d_contig_mean);
    ctx->launches += 1;
  }
  JG_CUDA(cudaGetLastError());
  return 0;
}

static_assert(sizeof(jg_sw_job) == sizeof(jg::SwJob), "jg_sw_job layout");
static_assert(sizeof(jg_sw_trace_job) == sizeof(jg::SwTraceJob), "jg_sw_trace_job layout");

int jg_sw_scan(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid, const jg_sw_job* d_jobs, int32_t n_jobs,
               int32_t threads, int32_t max_rows, int32_t max_cols, int32_t* d_out) {
  if (n_jobs <= 0) return 0;
  if (threads < 32 || threads > 1024 || threads % 32 != 0 || static_cast<long long>(threads) * jg::kSwRows < max_rows)
    return fail("jg_sw_scan: threads must be a multiple of 32 with threads * 16 >= max_rows");
  JG_CUDA(cudaSetDevice(ctx->device));
  const jg::SwSeq seq{d_codes, d_valid};
  const jg::SwScores sc{2, -100, 0, 100, 5};                       // utils/termini.py:113, 121-122
  const size_t smem = static_cast<size_t>(10) * threads * 4 + ((max_cols + 15) / 16) * 16;
  jg::sw_scan_kernel<<<n_jobs, threads, smem, ctx->stream>>>(seq, reinterpret_cast<const jg::SwJob*>(d_jobs), sc, d_out);
  ctx->launches += 1;
  JG_CUDA(cudaGetLastError());
  return 0;
}

int jg_sw_trace(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid, const jg_sw_trace_job* d_jobs, int32_t n_jobs,
                int32_t threads, int32_t max_rows, int32_t max_cols, uint8_t* d_scratch, int32_t* d_out) {
  if (n_jobs <= 0) return 0;
  if (threads < 32 || threads > 1024 || threads % 32 != 0 || static_cast<long long>(threads) * jg::kSwRows < max_rows)
    return fail("jg_sw_trace: threads must be a multiple of 32 with threads * 16 >= max_rows");
  JG_CUDA(cudaSetDevice(ctx->device));
  const jg::SwSeq seq{d_codes, d_valid};
  const jg::SwScores sc{2, -100, 0, 100, 5};
  const size_t smem = static_cast<size_t>(6) * threads * 4 + ((max_cols + 15) / 16) * 16;
  jg::sw_trace_kernel<<<n_jobs, threads, smem, ctx->stream>>>(seq, reinterpret_cast<const jg::SwTraceJob*>(d_jobs), sc, d_scratch, d_out);
  ctx->launches += 1;
  JG_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
