// Standalone hardware probe for the tcgen05 conv kernel (not part of the product library).
//   conv_probe <shape> <variant> [n_windows] [time_iters]
// shape:   0 = residual conv (Cin 128, Cout 128, k5 d3, SAME, shortcut + taps + pool)
//          1 = stem conv     (Cin 64,  Cout 128, k7 d1, VALID, raw tap)
//          2 = wide-dilation conv (Cin 64, Cout 64, k5 d8, SAME)
// variant: 0 = single-CTA kernel (conv_tc.cuh), 2 / 3 = CTA-pair kernel (conv_tc2.cuh), 4 = weights-stationary kernel (conv_ws.cuh)
// Compares conv_tc_kernel against conv_ref_kernel on the same random inputs and prints
// max |diff|; with time_iters > 0 also times the tensor-core kernel with CUDA events.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../jaeger_b200/csrc/conv_launch.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static uint32_t rng_state = 12345u;
static inline float frand() {  // uniform [-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static inline uint16_t f2h(float f) { const __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }   // fp16 storage
static inline float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

int main(int argc, char** argv) {
  const int shape = argc > 1 ? atoi(argv[1]) : 0;
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  const int n_win = argc > 3 ? atoi(argv[3]) : 2;
  const int iters = argc > 4 ? atoi(argv[4]) : 0;
  const int strip = argc > 5 ? atoi(argv[5]) : 0;  // 1: no taps/pool, 2: no activations/affine2, 4: no shortcut

  int cin, cout, k, dil, pad_left, l_in, l_out;
  const int frames = 6, period = 665, rpw = 4096;
  if (shape == 0) { cin = 128; cout = 128; k = 5; dil = 3; pad_left = 6; l_in = 659; l_out = 659; }
  else if (shape == 1) { cin = 64; cout = 128; k = 7; dil = 1; pad_left = 0; l_in = 665; l_out = 659; }
  else if (shape == 2) { cin = 64; cout = 64; k = 5; dil = 8; pad_left = 16; l_in = 640; l_out = 640; }
  else if (shape == 3) { cin = 128; cout = 256; k = 2; dil = 3; pad_left = 3; l_in = 659; l_out = 659; }
  else { cin = 128; cout = 64; k = 5; dil = 3; pad_left = 6; l_in = 659; l_out = 659; }

  const long long R = static_cast<long long>(n_win) * rpw;
  const long long plane = R + 2 * jg::kGuardRows;
  std::vector<uint16_t> hx(static_cast<size_t>(cin / 64) * plane * 64, 0), hsc(static_cast<size_t>(cout / 64) * plane * 64, 0);
  std::vector<uint8_t> hmask(R, 0), hscmask(R, 0);
  for (long long r = 0; r < R; ++r) {
    const int rw = static_cast<int>(r % rpw);
    const int f = rw / period, pos = rw % period;
    const bool in_frame_in = f < frames && pos < l_in;
    const bool in_frame_out = f < frames && pos < l_out;
    // knock a few rows out of the masks to exercise the masked paths
    const bool knocked = (r % 97) == 13;
    hmask[r] = in_frame_out && !knocked;
    hscmask[r] = in_frame_out && ((r % 89) != 7);
    if (in_frame_in)
      for (int c = 0; c < cin; ++c)
        hx[jg::kGuardRows * 64 + jg::act_index(r, c, plane)] = f2h(frand());
    if (hscmask[r])
      for (int c = 0; c < cout; ++c)
        hsc[jg::kGuardRows * 64 + jg::act_index(r, c, plane)] = f2h(frand());
  }
  const int ktot = k * cin;
  std::vector<uint16_t> hw(static_cast<size_t>(ktot) * cout);
  const float wscale = 1.0f / sqrtf(static_cast<float>(ktot));
  for (int t = 0; t < k; ++t)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        hw[jg::w_index(t, ci, co, cin, cout)] = f2h(frand() * wscale * 1.7f);
  std::vector<float> hpar(6 * cout);
  for (int c = 0; c < cout; ++c) {
    hpar[c] = getenv("JG_PROBE_FOLDED") ? 1.0f : 1.0f + 0.25f * frand();          // scale1
    hpar[cout + c] = 0.1f * frand();           // shift1
    hpar[2 * cout + c] = 1.0f + 0.25f * frand();
    hpar[3 * cout + c] = 0.1f * frand();
    hpar[4 * cout + c] = 0.05f * frand();      // bias
    hpar[5 * cout + c] = 0.3f * frand();       // sc_const
  }

  uint16_t *dx, *dsc, *dw, *dy_ref, *dy_tc;
  uint8_t *dmask, *dscmask;
  float *dpar, *dtap_ref, *dtap_tc, *dpool_ref, *dpool_tc;
  int* derr;
  const size_t ybytes = static_cast<size_t>(cout / 64) * plane * 128;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dsc, hsc.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&dy_ref, ybytes)); CK(cudaMalloc(&dy_tc, ybytes));
  CK(cudaMalloc(&dmask, R)); CK(cudaMalloc(&dscmask, R)); CK(cudaMalloc(&dpar, hpar.size() * 4));
  CK(cudaMalloc(&dtap_ref, n_win * cout * 4)); CK(cudaMalloc(&dtap_tc, n_win * cout * 4));
  CK(cudaMalloc(&dpool_ref, n_win * cout * 4)); CK(cudaMalloc(&dpool_tc, n_win * cout * 4));
  CK(cudaMalloc(&derr, 64)); CK(cudaMemset(derr, 0, 64));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsc, hsc.data(), hsc.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  // the CTA-pair kernel reads the same weights from its own (half-split) image
  std::vector<uint16_t> hw2(hw.size());
  for (int t = 0; t < k; ++t)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        hw2[jg::tc2::w2_index(t, ci, co, cin, cout, k)] = hw[jg::w_index(t, ci, co, cin, cout)];
  uint16_t* dw2; CK(cudaMalloc(&dw2, hw2.size() * 2));
  CK(cudaMemcpy(dw2, hw2.data(), hw2.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmask, hmask.data(), R, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dscmask, hscmask.data(), R, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpar, hpar.data(), hpar.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy_ref, 0, ybytes)); CK(cudaMemset(dy_tc, 0, ybytes));
  CK(cudaMemset(dtap_ref, 0, n_win * cout * 4)); CK(cudaMemset(dtap_tc, 0, n_win * cout * 4));
  std::vector<float> sentinel(static_cast<size_t>(n_win) * cout, -1.0e9f);
  CK(cudaMemcpy(dpool_ref, sentinel.data(), sentinel.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpool_tc, sentinel.data(), sentinel.size() * 4, cudaMemcpyHostToDevice));

  jg::ConvParams p{};
  p.x = reinterpret_cast<const jg::act_t*>(dx) + jg::kGuardRows * 64;
  p.sc = (shape == 0) ? reinterpret_cast<const jg::act_t*>(dsc) + jg::kGuardRows * 64 : nullptr;
  p.sc_mask = (shape == 0) ? dscmask : nullptr;
  p.sc_const = dpar + 5 * cout;
  p.out_mask = dmask;
  p.w = reinterpret_cast<const jg::act_t*>(dw);
  p.bias = dpar + 4 * cout;
  p.scale1 = dpar; p.shift1 = dpar + cout; p.scale2 = dpar + 2 * cout; p.shift2 = dpar + 3 * cout;
  p.x_plane = plane; p.y_plane = plane;
  p.n_tiles = static_cast<int>(R / jg::kTileM);
  p.rows_per_window = rpw;
  p.cin = cin; p.cout = cout; p.ntaps = k;
  int mn = 0, mx = 0;
  for (int t = 0; t < k; ++t) { p.shifts[t] = t * dil - pad_left; mn = p.shifts[t] < mn ? p.shifts[t] : mn; mx = p.shifts[t] > mx ? p.shifts[t] : mx; }
  p.halo_l = -mn; p.halo_r = mx;
  p.act1 = jg::ACT_GELU_TANH; p.act2 = jg::ACT_GELU_TANH;
  p.has_affine2 = (shape == 0);
  p.tap_mode = (shape == 0) ? 2 : (shape == 1 ? 1 : 0);
  p.pool_mode = (shape == 0) ? 1 : 0;
  if (strip & 1) { p.tap_mode = 0; p.pool_mode = 0; }
  if (strip & 2) { p.act1 = jg::ACT_NONE; p.act2 = jg::ACT_NONE; p.has_affine2 = 0; }
  if (strip & 4) { p.sc = nullptr; p.sc_mask = nullptr; }
  if (strip & 16) { p.has_affine2 = 0; p.act2 = jg::ACT_NONE; }
  if (strip & 32) { p.pool_mode = 0; }
  p.err = derr;
  p.folded = getenv("JG_PROBE_FOLDED") ? 1 : 0;   // with it, scale1 must be 1 (set below) and the specialised epilogues run
  (void)variant;

  int dev_sms = 0; CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  printf("shape %d variant %d windows %d tiles %d sms %d stages %d\n", shape, variant, n_win, p.n_tiles, dev_sms, jg::conv_tc_stages(p));

  jg::ConvParams pr = p;
  pr.y = reinterpret_cast<jg::act_t*>(dy_ref) + jg::kGuardRows * 64; pr.tap_sum = dtap_ref; pr.pool = dpool_ref;
  CK(jg::launch_conv_ref(pr, 0));
  CK(cudaDeviceSynchronize());

  jg::ConvParams pt = p;
  pt.y = reinterpret_cast<jg::act_t*>(dy_tc) + jg::kGuardRows * 64; pt.tap_sum = dtap_tc; pt.pool = dpool_tc;
  if (strip & 8) { pt.y = nullptr; }
  auto launch = [&](const jg::ConvParams& q) {
    return variant == 4 ? jg::launch_conv_ws(q, dev_sms, 0) : (variant >= 2 ? jg::launch_conv_tc2(q, dev_sms, 0, variant) : jg::launch_conv_tc(q, dev_sms, 0));
  };
  if (variant >= 2) pt.w = reinterpret_cast<const jg::act_t*>(dw2);
  if (variant == 4) {      // weights-stationary image: Wt[cout][tap * cin + ci]
    std::vector<uint16_t> hw3(hw.size());
    for (int t = 0; t < k; ++t)
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co)
          hw3[jg::ws::w3_index(t, ci, co, cin, k)] = hw[jg::w_index(t, ci, co, cin, cout)];
    uint16_t* dw3; CK(cudaMalloc(&dw3, hw3.size() * 2));
    CK(cudaMemcpy(dw3, hw3.data(), hw3.size() * 2, cudaMemcpyHostToDevice));
    pt.w = reinterpret_cast<const jg::act_t*>(dw3);
    if (strip & 8) pt.y = nullptr;
    printf("ws mode %d\n", jg::conv_ws_mode(pt));
    if (jg::conv_ws_mode(pt) < 0) { printf("layer not eligible for the weights-stationary kernel\n"); return 4; }
  }
  CK(launch(pt));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    int herr = -1; cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    printf("tc kernel failed: %s (err code %d)\n", cudaGetErrorString(se), herr);
    return 3;
  }

  std::vector<uint16_t> yr(ybytes / 2), yt(ybytes / 2);
  CK(cudaMemcpy(yr.data(), dy_ref, ybytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(yt.data(), dy_tc, ybytes, cudaMemcpyDeviceToHost));
  double maxdiff = 0, maxref = 0; size_t nbad = 0;
  for (size_t i = 0; i < yr.size(); ++i) {
    const double a = h2f(yr[i]), b = h2f(yt[i]);
    const double d = fabs(a - b);
    if (d > maxdiff) maxdiff = d;
    if (fabs(a) > maxref) maxref = fabs(a);
    if (d > 0.03 + 0.02 * fabs(a)) {
      if (nbad < 6 || (nbad % 4000037) == 0) {   // decode [group][row][64] -> (row, channel) ignoring the chunk swizzle
        const size_t g = i / (static_cast<size_t>(plane) * 64), rem = i % (static_cast<size_t>(plane) * 64);
        printf("  bad at group %zu row %lld chunk %zu elem %zu: ref %.4f got %.4f\n", g, static_cast<long long>(rem / 64) - jg::kGuardRows,
               (rem % 64) / 8, rem % 8, a, b);
      }
      ++nbad;
    }
  }
  std::vector<float> tr(n_win * cout), tt(n_win * cout), pr_(n_win * cout), pt_(n_win * cout);
  CK(cudaMemcpy(tr.data(), dtap_ref, tr.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tt.data(), dtap_tc, tt.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pr_.data(), dpool_ref, tr.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pt_.data(), dpool_tc, tr.size() * 4, cudaMemcpyDeviceToHost));
  double tapdiff = 0, tapmax = 0, pooldiff = 0;
  for (size_t i = 0; i < tr.size(); ++i) {
    tapdiff = fmax(tapdiff, fabs(tr[i] - tt[i])); tapmax = fmax(tapmax, fabs(tr[i]));
    pooldiff = fmax(pooldiff, fabs(pr_[i] - pt_[i]));
  }
  printf("y: max|ref| %.4f max|diff| %.5f bad %zu / %zu ; tap max|ref| %.3f max|diff| %.5f ; pool max|diff| %.5f\n",
         maxref, maxdiff, nbad, yr.size(), tapmax, tapdiff, pooldiff);
  const bool ok = nbad == 0 && tapdiff <= 1e-3 * (1.0 + tapmax) * 4 && pooldiff < 0.05;
  printf("RESULT shape %d variant %d: %s\n", shape, variant, ok ? "MATCH" : "MISMATCH");

  if (getenv("JG_TRACE") && variant == 4) {
    long long* ddbg; CK(cudaMalloc(&ddbg, 1024 * 8));
    jg::ConvParams pd = pt; pd.dbg = ddbg;
    for (int i = 0; i < 3; ++i) CK(launch(pt));
    CK(cudaMemset(ddbg, 0, 1024 * 8));
    CK(launch(pd));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(32);
    CK(cudaMemcpy(h.data(), ddbg, h.size() * 8, cudaMemcpyDeviceToHost));
    const double ns = double(h[15]);
    const double own = ns / 3.0;       // sub-tiles of epilogue group 0
    printf("ws CTA0: %lld cycles in %lld ns (%.0f MHz), %lld sub-tiles (%.0f / sub-tile; MMA floor %d)\n", h[0], h[16], 1e3 * h[0] / double(h[16]),
           h[15], h[0] / ns, 4 * k * (cin / 64) * 32);
    printf("  MMA warp per sub-tile: wait free acc %.0f, wait operands %.0f, issue %.0f | epilogue group 0 per own sub-tile: validity %.0f, slot %.0f, accumulator %.0f, math %.0f, pair barrier %.0f, other %.0f\n",
           h[1] / ns, h[2] / ns, h[13] / ns, h[7] / own, h[5] / own, h[3] / own, h[9] / own, h[11] / own,
           (h[14] - h[7] - h[5] - h[3] - h[9] - h[11]) / own);
  } else if (getenv("JG_TRACE")) {
    long long* ddbg; CK(cudaMalloc(&ddbg, 1024 * 8)); CK(cudaMemset(ddbg, 0, 1024 * 8));
    jg::ConvParams pd = pt; pd.dbg = ddbg;
    for (int i = 0; i < 3; ++i) CK(launch(pt));
    CK(launch(pd));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(1024);
    CK(cudaMemcpy(h.data(), ddbg, h.size() * 8, cudaMemcpyDeviceToHost));
    const long long t0 = h[0];
    printf("trace (cycles rel. to first MMA start): it  mma_start mma_issued | epi_arrive_wait tfull_ready epi_done\n");
    for (int i = 0; i < 16 && h[i * 8] != 0; ++i)
      printf("  %2d  %8lld %8lld | %8lld %8lld %8lld\n", i, h[i*8]-t0, h[i*8+1]-t0, h[i*8+2]-t0, h[i*8+3]-t0, h[i*8+4]-t0);
    if (variant >= 2 && h[600] != 0) {
      printf("pair epilogue trace, group 0 warp 0 (cycles rel. to tile start; prev = since previous tile start):\n  it  prev | rowvalid+sc  tfull | ld0 b0 ld1 b1 ld2 b2 ld3 b3 | end\n");
      for (int i = 0; i < 24 && h[600 + 16 * i] != 0; ++i) {
        const long long* t = &h[600 + 16 * i];
        printf("  %2d %6lld | %6lld %6lld |", i, i ? t[0] - h[600 + 16 * (i - 1)] : 0LL, t[1] - t[0], t[2] - t[0]);
        for (int k = 3; k <= 10; ++k) printf(" %5lld", t[k] - t[0]);
        printf(" | %6lld\n", t[11] - t[0]);
      }
    }
    const int my_tiles = p.n_tiles / dev_sms;   // per CTA (pair kernel: 128-row tiles of this CTA)
    printf("CTA0: total %lld cycles for %d tiles (%.0f /tile); first MMA starts at +%lld; waits: producer(free stage) %lld, MMA(free acc) %lld, MMA(operands) %lld, epi groups(wait MMA) %lld %lld %lld\n",
           h[523] - h[519], my_tiles, double(h[523] - h[519]) / my_tiles, t0 - h[519], h[520], h[521], h[522], h[524], h[525], h[526]);
  }
  if (iters > 0 && ok) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // warm up for ~1.5 s so the SM clock has left its idle state before timing
    {
      cudaEvent_t w0, w1; cudaEventCreate(&w0); cudaEventCreate(&w1);
      float wms = 0; CK(cudaEventRecord(w0));
      while (wms < 1500.0f) {
        for (int i = 0; i < 20; ++i) CK(launch(pt));
        CK(cudaEventRecord(w1)); CK(cudaEventSynchronize(w1));
        { int herr[8]; CK(cudaMemcpy(herr, derr, 32, cudaMemcpyDeviceToHost));
          if (herr[0]) { printf("STUCK WAIT code %d it %d cta %d parity %d warp %d (1 producer/EMPTY 2 MMA/TEMPTY 3 MMA/FULL 4 helper/SFREE 5 helper/VEMPTY 6 epi/VFULL 7 epi/SFULL 8 epi/SFREE 9 epi/TFULL)\n", herr[0], herr[1], herr[2], herr[3], herr[4]); return 5; } }
        cudaEventElapsedTime(&wms, w0, w1);
      }
    }
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(launch(pt));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    const double flops = 2.0 * static_cast<double>(R) * ktot * cout;
    printf("TIMING strip %d shape %d windows %d: %.3f ms/launch, %.1f TFLOP/s (rows incl. gaps), %.1f us/tile-wave\n",
           strip, shape, n_win, ms, flops / ms * 1e-9, ms * 1e3 / ((p.n_tiles + dev_sms - 1) / dev_sms));
  }
  return ok ? 0 : 1;
}
