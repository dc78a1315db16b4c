// Host-only check of the kernels' shared-memory / TMEM geometry functions over a sweep of model
// shapes (runs on the CPU: no kernel is launched).  Built and run by tests/test_geometry_host.py.
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <string>
#include "../../mujoco-mbrl_b200/csrc/rollout_tcf.cuh"
#include "../../mujoco-mbrl_b200/csrc/replay.cuh"
#include "../../mujoco-mbrl_b200/csrc/select.cuh"
using namespace mbrl;

#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d " #cond " ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++fails; } } while (0)

// Walks the weight-streaming kernel's operand image exactly as its producer / MMA warps do (parts,
// ring stages, K-step tiles, canonical K-major addressing with LBO = rows*16) and checks that every
// B-operand element the MMAs would read is the weight the math needs.
static float half_bits_to_float(uint16_t b) { __half h; std::memcpy(&h, &b, 2); return __half2float(h); }
static int check_tcw_stream(int O, int A, int U, size_t max_smem) {
  int fails = 0;
  TcwGeom g{};
  std::string why;
  if (!tcw_geometry(O, A, U, max_smem, &g, &why)) { std::printf("FAIL tcw_geometry(%d,%d,%d): %s\n", O, A, U, why.c_str()); return 1; }
  const int D = O + A;
  std::vector<float> W1((size_t)U * D), b1(U), W2((size_t)U * U), b2(U), W3((size_t)O * U);
  auto val = [](int a, int b, int c) { return (float)((a * 131 + b * 17 + c * 7) % 1021 - 510) / 64.0f; };  // exact in fp16
  for (int u = 0; u < U; ++u) { b1[u] = val(u, 1, 1); b2[u] = val(u, 2, 2); for (int d = 0; d < D; ++d) W1[(size_t)u * D + d] = val(u, d, 3); for (int k = 0; k < U; ++k) W2[(size_t)u * U + k] = val(u, k, 4); }
  for (int o = 0; o < O; ++o) for (int k = 0; k < U; ++k) W3[(size_t)o * U + k] = val(o, k, 5);
  std::vector<uint16_t> img;
  tcw_pack(g, true, W1.data(), b1.data(), W2.data(), b2.data(), W3.data(), &img);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(img.data());
  auto elem = [&](size_t tile_off, int rows, int n, int kk) {  // element (n, kk) of one K-step tile
    uint16_t b; std::memcpy(&b, base + tile_off + (size_t)(kk / 8) * rows * 16 + (size_t)n * 16 + (kk % 8) * 2, 2);
    return half_bits_to_float(b);
  };
  size_t off = 0;
  for (int part = 0; part < 5; ++part) {
    const int T = part < 2 ? g.Kx / 16 : (part < 4 ? g.KH + 1 : g.KH);
    const int tps = part < 4 ? g.tps_h : g.tps_y, tile = part < 4 ? g.tile_h : g.tile_y, rows = part < 4 ? g.Nc : g.Op;
    const int total = part < 2 ? g.p1_bytes : (part < 4 ? g.p2_bytes : g.p3_bytes);
    int consumed = 0;
    for (int t0 = 0; t0 < T; t0 += tps) {
      const int n_t = std::min(tps, T - t0);
      const int stage_bytes = std::min(tps * tile, total - consumed);  // what the producer copies into this stage
      if (stage_bytes != n_t * tile || stage_bytes > g.stage_bytes || stage_bytes % 16) { std::printf("FAIL stage size part %d\n", part); ++fails; }
      for (int i = 0; i < n_t; ++i) {
        const int ks = t0 + i;
        for (int n = 0; n < rows; ++n)
          for (int kk = 0; kk < 16; ++kk) {
            const int k = 16 * ks + kk;
            const float got = elem(off + (size_t)i * tile, rows, n, kk);
            float want = 0.f;
            const int c = part & 1, u = c * g.Nc + n;
            if (part < 2) {  // layer 1: k indexes the input tile [actions | 1 | pad | state]
              if (u < U) { if (k < A) want = W1[(size_t)u * D + O + k]; else if (k == A) want = b1[u]; else if (k >= g.Ka && k < g.Ka + O) want = W1[(size_t)u * D + k - g.Ka]; }
            } else if (part < 4) {
              if (u < U) { if (k < U) want = W2[(size_t)u * U + k]; else if (k == g.Np) want = b2[u]; }
            } else {
              if (n < O && k < U) want = W3[(size_t)n * U + k];
            }
            if (got != want && fails < 10) { std::printf("FAIL stream O=%d A=%d U=%d part %d ks %d n %d kk %d: got %g want %g\n", O, A, U, part, ks, n, kk, got, want); ++fails; }
          }
      }
      off += stage_bytes; consumed += stage_bytes;
    }
    if (consumed != total) { std::printf("FAIL part %d bytes\n", part); ++fails; }
  }
  if (off != (size_t)g.w_bytes) { std::printf("FAIL image size\n"); ++fails; }
  return fails;
}

int main() {
  const size_t max_smem = 232448;  // sharedMemPerBlockOptin of sm_100
  int fails = 0, fused_ok = 0, reg_ok = 0, total = 0, wide_ok = 0;
  for (int O : {5, 17, 40, 67, 128})
    for (int A : {1, 6, 8, 16, 21, 31})
      for (int U : {50, 100, 200, 256, 300, 448, 512}) fails += check_tcw_stream(O, A, U, max_smem);
  for (int O = 1; O <= 80; O += (O < 8 ? 1 : 7))
    for (int A = 1; A <= 31; A += (A < 8 ? 1 : 5))
      for (int U = 8; U <= 520; U += (U < 64 ? 7 : 13)) {
        ++total;
        TcfGeom g{};
        std::string why;
        if (tcf_geometry(O, A, U, max_smem, &g, &why)) {
          ++fused_ok;
          CHECK(g.Np % 16 == 0 && g.Np > U && g.Oy % 16 == 0 && g.Oy >= O && g.Ka % 16 == 0 && g.Ka > A && g.Ks % 16 == 0 && g.Ks >= O,
                "O=%d A=%d U=%d", O, A, U);
          CHECK(g.Na == g.Np + g.Oy && g.Na <= 256, "Na=%d", g.Na);
          CHECK(g.Np / 16 <= kTcfMaxKSteps, "K-steps %d", g.Np / 16);
          CHECK(kTcD2Col + g.Np <= 512 && g.Na <= kTcD2Col, "TMEM columns: D_A %d D_B at %d + %d", g.Na, kTcD2Col, g.Np);
          CHECK(g.waa_off % 16 == 0 && g.w1s_off % 16 == 0 && g.w2_off % 16 == 0 && g.w_bytes % 16 == 0, "operand image offsets");
          CHECK(g.xs_off % 128 == 0 && g.xa_off % 16 == 0 && g.bar_off % 8 == 0 && g.ms_off % 4 == 0, "tile / barrier offsets");
          CHECK(g.tab_off >= g.w_bytes && g.xs_off >= g.tab_off + (6 * g.Oy + 2 * kMaxAct + 2 * kTcRows) * 4, "table region");
          CHECK((size_t)g.smem_bytes <= max_smem && g.ms_floats >= 0 && g.ms_off + 4 * g.ms_floats == g.smem_bytes, "smem %d", g.smem_bytes);
          CHECK(g.xch_off == -1 || (g.xch_off % 4 == 0 && g.xch_off + kTcfXchSlots * 2 * kTcRows * 4 == g.ms_off), "control-term slots");
        } else {
          CHECK(!why.empty(), "rejected without a reason O=%d A=%d U=%d", O, A, U);
        }
        TcwGeom w{};
        std::string why_w;
        if (tcw_geometry(O, A, U, max_smem, &w, &why_w)) {
          ++wide_ok;
          CHECK(w.Np % 64 == 0 && w.Np >= U && w.Np <= 512 && w.Nc * 2 == w.Np && w.Nc % 32 == 0 && w.Nc <= 256, "Np=%d", w.Np);
          CHECK(w.Ka % 8 == 0 && w.Ka > A && w.Kx % 16 == 0 && w.Kx >= w.Ka + O && w.Op % 16 == 0 && w.Op >= O && w.Op <= 128, "K/N padding");
          CHECK(w.QA * 8 == w.Ka && w.SC * 8 == w.Kx - w.Ka && w.SC * 8 >= O, "input tile chunks");
          CHECK(kTcwAccCol + w.Nc <= 512 && w.Nc <= 256 && w.ycol >= w.Nc / 2 && w.ycol + w.Op <= kTcwAccCol, "TMEM columns: y at %d", w.ycol);
          CHECK(w.tps_h >= 1 && w.tps_y >= 1 && w.tps_h * w.tile_h <= w.stage_bytes && w.tps_y * w.tile_y <= w.stage_bytes, "stage tiles");
          CHECK(w.xa_off % 128 == 0 && w.xs_off % 128 == 0 && w.one_off % 128 == 0 && w.h2_off % 128 == 0 && w.ring_off % 128 == 0 && w.bar_off % 8 == 0, "offsets");
          CHECK(w.xa_off >= (6 * w.Op + 2 * kMaxAct + 18 * kTcRows) * 4 && w.xs_off > w.xa_off, "tables / tile order (LBO = xs - xa must be positive)");
          CHECK(w.stages >= 3 && w.stages <= kTcwMaxStages && (size_t)w.smem_bytes <= max_smem && w.ms_off + 4 * w.ms_floats == w.smem_bytes, "smem %d", w.smem_bytes);
          CHECK(w.w_bytes == 2 * w.p1_bytes + 2 * w.p2_bytes + w.p3_bytes && w.w_bytes % 16 == 0, "image size");
        } else {
          CHECK(U > 512 - 63 || !why_w.empty(), "wide geometry rejected O=%d A=%d U=%d: %s", O, A, U, why_w.c_str());
        }
        for (int H : {1, 20, 30, 50}) {
          const RegGeom r = replay_reg_geometry(O, A, U, H);
          if (!r.ok) continue;
          ++reg_ok;
          CHECK(r.Q * r.S <= kRegThreads && r.S * r.kpt2 >= U && r.kpt2 % 4 == 0 && r.kpt2 <= 40, "U=%d S=%d kpt2=%d Q=%d", U, r.S, r.kpt2, r.Q);
          CHECK(r.ldu >= U && r.ldu % 4 == 0 && r.kp1 >= O + A && r.kp1 % 8 == 0 && r.kp2 == r.S * r.kpt2, "padding");
          CHECK(r.ld3 >= 8 * r.k3 && 8 * r.k3 >= r.kp2 && r.o3 >= O && r.o3 % 4 == 0, "layer-3 layout");
          CHECK(r.x >= H * A && r.y == r.x + r.kp1 && r.h1 % 4 == 0 && r.h2 == r.h1 + r.kp2 && r.part % 4 == 0 && r.w1 % 4 == 0, "offsets");
          CHECK(r.total == r.w3 + r.o3 * r.ld3, "total");
          const ReplayLayout L = replay_layout(O, A, U, H, true);
          CHECK(L.total > 0 && L.w1 % 4 == 0 && L.part % 4 == 0, "smem-weight replay layout");
        }
      }
  // ---- peer-memory packet buffer (select.cuh): regions in order, 8-byte packets, 16-byte parity halves ----
  int layouts = 0;
  for (int world : {1, 2, 3, 4, 8, 16, 64})
    for (int slot : {1, 7, 73, 409, 2026, 13107, 16384})
      for (int pslots : {1, 30, 60, 300}) {
        ++layouts;
        const size_t idx = p2p_idx_off(world, slot), rank = p2p_rank_off(world, slot), part = p2p_part_off(world, slot);
        const size_t par = p2p_parity_words(world, slot, pslots), total = p2p_total_words(world, slot, pslots);
        CHECK(idx == 2 * (size_t)world * slot && rank == idx + 2 * (size_t)slot && part == rank + 6 * (size_t)world, "region order");
        CHECK(idx % 2 == 0 && rank % 2 == 0 && part % 2 == 0, "8-byte packets world=%d slot=%d", world, slot);
        CHECK(par >= part + 16 * (size_t)world * pslots && par % 4 == 0 && total == 2 * par, "parity halves world=%d slot=%d pslots=%d", world, slot, pslots);
      }
  // ---- select / refit launch helpers ----
  for (int n : {1, 31, 32, 33, 16384, 16385, 49152}) CHECK(select_padded(n) >= n && select_padded(n) % 32 == 0 && select_padded(n) < n + 32, "select_padded(%d)", n);
  CHECK(sizeof(uint32_t) * (size_t)select_padded(kSelectStageMax) + sizeof(uint32_t) * kSelectBins + 2048 <= max_smem, "top-k staging + histogram fit in shared memory");
  for (int k : {1, 31, 32, 33, 204, 1023, 1024, 1638, 2048, 2049, 13107, 131072}) {
    const int t = refit_threads(k);
    CHECK(t % 32 == 0 && t >= 32 && t <= kRefitThreads && (k > kRefitChunk || t >= std::min(k, kRefitThreads)), "refit_threads(%d) = %d", k, t);
  }
  for (int s = 1; s <= kTcfSpecs; ++s) {  // every compile-time geometry class is a geometry the run-time function produces
    const int O = s == 1 ? 17 : 5, A = tcf_spec_a(s), U = s == 1 ? 200 : 50;
    TcfGeom g{};
    std::string why;
    CHECK(tcf_geometry(O, A, U, max_smem, &g, &why) && tcf_matches_spec(g) == s, "geometry class %d", s);
  }
  {
    TcfGeom g{};
    std::string why;
    CHECK(tcf_geometry(24, 6, 200, max_smem, &g, &why) && tcf_matches_spec(g) == 1, "walker shape is class 1");
    CHECK(tcf_geometry(17, 7, 200, max_smem, &g, &why) && tcf_matches_spec(g) == 0, "another action count is no class");
  }
  std::printf("packet layouts checked: %d\n", layouts);
  std::printf("geometry sweep: %d shapes, fused tensor-core geometry accepted %d, weight-streaming geometry accepted %d, register replay accepted %d, failures %d\n",
              total, fused_ok, wide_ok, reg_ok, fails);
  return fails ? 1 : 0;
}
