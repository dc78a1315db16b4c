// Host-only check of the kernels' shared-memory / TMEM geometry functions over a sweep of model
// shapes (runs on the CPU: no kernel is launched).  Built and run by tests/test_geometry_host.py.
#include <cstdio>
#include <string>
#include "../../mujoco-mbrl_b200/csrc/rollout_tcf.cuh"
#include "../../mujoco-mbrl_b200/csrc/replay.cuh"
using namespace mbrl;

#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d " #cond " ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++fails; } } while (0)

int main() {
  const size_t max_smem = 232448;  // sharedMemPerBlockOptin of sm_100
  int fails = 0, fused_ok = 0, reg_ok = 0, total = 0;
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
        } else {
          CHECK(!why.empty(), "rejected without a reason O=%d A=%d U=%d", O, A, U);
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
  std::printf("geometry sweep: %d shapes, fused tensor-core geometry accepted %d, register replay accepted %d, failures %d\n",
              total, fused_ok, reg_ok, fails);
  return fails ? 1 : 0;
}
