// selftest.cpp -- exercises the C++ host mirror (goal_gx.hpp) the way GoalPrimal drives the reference:
// build the discretisation, set a solution, compute_jacob / compute_resid, update states, localize.
// Prints checksums that tests/test_gpu_host_mirror.py compares with the oracle.
//   usage: gx_selftest <model> <n>      (Kuhn cube with n cells per side)
#include <cmath>
#include <cstdio>
#include <cstring>

#include "goal_gx.hpp"

static void kuhn(int n, std::vector<double>& co, std::vector<gx::LO>& cn) {
  int const s = n + 1;
  for (int k = 0; k <= n; ++k) for (int j = 0; j <= n; ++j) for (int i = 0; i <= n; ++i) {
    co.push_back((double)i / n); co.push_back((double)j / n); co.push_back((double)k / n);
  }
  int const perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  int const stride[3] = {1, s, s * s};
  for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
    int const base = i + s * j + s * s * k;
    for (auto const& p : perms) {
      int v[4] = {base, 0, 0, 0};
      for (int t = 0; t < 3; ++t) v[t + 1] = v[t] + stride[p[t]];
      int inv = 0;
      for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) inv += p[a] > p[b];
      if (inv & 1) std::swap(v[1], v[2]);
      cn.insert(cn.end(), v, v + 4);
    }
  }
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s <neohookean|J2> <n>\n", argv[0]); return 2; }
  try {
    std::vector<double> co; std::vector<gx::LO> cn;
    kuhn(std::atoi(argv[2]), co, cn);
    gx::Disc disc(co, cn, argv[1], {{1000.0, 0.25, 100.0, 10.0, 1.0}});
    gx::Primal primal(&disc);
    gx::States states(&disc);
    size_t const nn = disc.get_num_nodes();
    std::vector<double> u(3 * nn), p(nn);
    for (size_t a = 0; a < nn; ++a) {
      double const* x = &co[3 * a];
      u[3 * a] = 0.004 * x[0] + 2e-3 * std::sin(6.283185307179586 * x[1]);
      u[3 * a + 1] = -0.0016 * x[1] + 2e-3 * std::sin(6.283185307179586 * x[2]);
      u[3 * a + 2] = -0.0016 * x[2] + 2e-3 * std::sin(6.283185307179586 * x[0]);
      p[a] = std::cos(3.0 * x[0] + 2.0 * x[1] - x[2]);
    }
    primal.set_solution(u, p);
    primal.compute_jacob();
    auto const& g = primal.get_sol_info()->ghost;
    double sR = 0, sA = 0;
    for (double v : g.R) sR += v * v;
    for (double v : g.dRdu) sA += v * v;
    std::printf("nodes %d elems %d nnz %lld\n", disc.get_num_nodes(), disc.get_num_elems(), (long long)disc.get_nnz());
    std::printf("jacob |R|^2 %.17e |A|^2 %.17e\n", sR, sA);
    primal.compute_resid();
    sR = 0;
    for (double v : primal.get_sol_info()->ghost.R) sR += v * v;
    std::printf("resid |R|^2 %.17e\n", sR);
    std::vector<double> sig;
    states.get("sigma", sig);
    double sS = 0;
    for (double v : sig) sS += v * v;
    std::printf("sigma |s|^2 %.17e\n", sS);
    states.update();
    gx::NestedAdjoint adj(&disc);
    std::vector<double> zu(3 * nn, 1.0), zp(nn, 1.0);
    adj.localize(zu, zp, zp);  // z == 1: must reproduce the plain residual
    double sE = 0;
    for (double v : adj.get_sol_info()->ghost.R) sE += v * v;
    std::printf("localize(z=1) |R|^2 %.17e\n", sE);
    // ---- the steps either side of the assembly: functional + dMdu, traction term, solution update, size field
    {
      primal.set_solution(u, p);
      primal.compute_resid();  // leaves the ghost R on the device and the sigma state "max vm" reads
      gx::Functional avm(&disc, "avg vm"), ks(&disc, "max vm", 0, 0.05);
      std::vector<double> dMdu;
      avm.compute_adjoint_rhs(dMdu);
      double sD = 0;
      for (double v : dMdu) sD += v * v;
      ks.compute();
      std::printf("functional avg_vm %.17e |dMdu|^2 %.17e max_vm %.17e\n", avm.get_value(), sD, ks.get_value());
      // traction (0.3, -1, 0.25) on the faces of the first element: three sides sharing nodes
      std::vector<gx::LO> sides = {cn[1], cn[2], cn[3], cn[0], cn[3], cn[2], cn[0], cn[1], cn[3]};
      std::vector<double> T;
      for (int k = 0; k < 3; ++k) T.insert(T.end(), {0.3, -1.0, 0.25});
      gx::set_tbcs(&disc, sides, T);
      std::vector<double> R(4 * nn);
      if (gx_fetch(disc.ctx, R.data(), nullptr)) gx::fail(gx_last_error(disc.ctx));
      sR = 0;
      for (double v : R) sR += v * v;
      std::printf("resid+tbcs |R|^2 %.17e\n", sR);
      std::vector<double> du(4 * nn);
      for (size_t k = 0; k < du.size(); ++k) du[k] = 1e-4 * std::sin(0.37 * (double)k);
      gx::add_soln(&disc, du);
      primal.compute_resid();
      sR = 0;
      for (double v : primal.get_sol_info()->ghost.R) sR += v * v;
      std::printf("resid(u+du) |R|^2 %.17e\n", sR);
      std::vector<double> eta(disc.get_num_elems()), vs;
      for (size_t e = 0; e < eta.size(); ++e) eta[e] = 1e-5 * (1.0 + (double)(e % 7));
      gx::get_iso_target_size(&disc, eta, 2 * disc.get_num_elems(), vs);
      double sV = 0;
      for (double v : vs) sV += v;
      std::printf("size field sum %.17e\n", sV);
    }
    try {  // error behaviour: unknown state name -> fail()
      std::vector<double> bad;
      states.get("no_such_state", bad);
      std::printf("ERROR: missing failure\n");
      return 1;
    } catch (std::runtime_error const& e) { std::printf("fail() ok: %s\n", e.what()); }
  } catch (std::exception const& e) { std::fprintf(stderr, "FAILED: %s\n", e.what()); return 1; }
  return 0;
}
