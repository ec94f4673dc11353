// Test-only: the per-point / per-registration device math of lm.cuh compiled for the HOST
// (the functions are __host__ __device__), so the CPU suite can compare it with the oracle
// without a GPU.  Not part of liblisreg.so.
#include "lm.cuh"
using namespace lisreg;
extern "C" {
int hc_corner_coeff(const float* q, const float* nb15, float* raw5) {
  float4 nb[5]; for (int j = 0; j < 5; j++) nb[j] = make_float4(nb15[3*j], nb15[3*j+1], nb15[3*j+2], 0.f);
  float raw[5] = {0,0,0,0,0}; bool ok = corner_coeff(q[0], q[1], q[2], nb, raw);
  for (int i = 0; i < 5; i++) raw5[i] = raw[i]; return ok ? 1 : 0;
}
int hc_surf_coeff(const float* q, const float* nb15, float* raw5) {
  float4 nb[5]; for (int j = 0; j < 5; j++) nb[j] = make_float4(nb15[3*j], nb15[3*j+1], nb15[3*j+2], 0.f);
  float raw[5] = {0,0,0,0,0}; bool ok = surf_coeff(q[0], q[1], q[2], nb, raw);
  for (int i = 0; i < 5; i++) raw5[i] = raw[i]; return ok ? 1 : 0;
}
void hc_jacobi3(const float* A, float* W, float* V) { float a[9]; int ir[3], ic[3]; for (int i = 0; i < 9; i++) a[i] = A[i]; jacobi_eigen<3>(a, W, V, ir, ic); }
void hc_jacobi3_reg(const float* A, float* W, float* V) {
  float w[3], v[9]; jacobi_eigen3(A[0], A[1], A[2], A[4], A[5], A[8], w, v);
  for (int i = 0; i < 3; i++) W[i] = w[i]; for (int i = 0; i < 9; i++) V[i] = v[i];
}
void hc_jacobi6(const float* A, float* W, float* V) { float a[36]; int ir[6], ic[6]; for (int i = 0; i < 36; i++) a[i] = A[i]; jacobi_eigen<6>(a, W, V, ir, ic); }
int hc_qr6(const float* A, const float* b, float* x) { float a[36]; for (int i = 0; i < 36; i++) a[i] = A[i]; for (int i = 0; i < 6; i++) x[i] = b[i]; return qr_solve<6>(a, x); }
void hc_plane(const float* A15, float* x) { float a[15]; const float b[5] = {-1,-1,-1,-1,-1}; for (int i = 0; i < 15; i++) a[i] = A15[i]; colpiv_qr_solve_5x3(a, b, x); }
void hc_state_refresh(const float* pose, float* T12, float* trig6) {
  RegState s; for (int i = 0; i < 6; i++) s.pose[i] = pose[i]; state_refresh(s);
  for (int i = 0; i < 12; i++) T12[i] = s.T[i]; for (int i = 0; i < 6; i++) trig6[i] = s.trig[i];
}
// one LMOptimization tail from the 29 sums; pose in/out
void hc_solve_tail(float* pose, int iter, int degenerate_in, const double* sums, const lisreg_lm_params* p, lisreg_lm_iter* log, int* out_flags) {
  RegState st; memset(&st, 0, sizeof(st));
  for (int i = 0; i < 6; i++) st.pose[i] = pose[i];
  st.iter = iter; st.degenerate = degenerate_in;
  LmParamsDev d; memset(&d, 0, sizeof(d));
  d.max_iters = p->max_iters; d.early_exit = p->early_exit; d.gate = p->sqdist_gate; d.conv_rot = p->conv_rot_deg; d.conv_trans = p->conv_trans_cm;
  d.min_sel = p->min_sel; d.degenerate_eig = p->degenerate_eig; d.rot_tol = p->rot_tolerance; d.z_tol = p->z_tolerance;
  SolveScratch sc; lm_solve_tail(st, d, sums, log, sc);
  for (int i = 0; i < 6; i++) pose[i] = st.pose[i];
  out_flags[0] = st.done; out_flags[1] = st.converged; out_flags[2] = st.degenerate; out_flags[3] = st.any_small;
}
}
// host build of the grid index (counting sort) + knn5_grid, to test the search logic on the CPU
#include <vector>
extern "C" void hc_knn5(const float* map4, int n, const float* q4, int nq, float h, float gate, int* idx, float* sqd) {
  GridDev g; float mn[3] = {1e30f,1e30f,1e30f}, mx[3] = {-1e30f,-1e30f,-1e30f};
  for (int i = 0; i < n; i++) for (int d = 0; d < 3; d++) { mn[d] = fminf(mn[d], map4[4*i+d]); mx[d] = fmaxf(mx[d], map4[4*i+d]); }
  g.h = h; g.inv_h = 1.0f / h;
  g.nx = (int)floorf((mx[0]-mn[0])/h)+2; g.ny = (int)floorf((mx[1]-mn[1])/h)+2; g.nz = (int)floorf((mx[2]-mn[2])/h)+2;
  g.ox = mn[0]-0.5f*h; g.oy = mn[1]-0.5f*h; g.oz = mn[2]-0.5f*h; g.n = n; g.ncells = g.nx*g.ny*g.nz;
  std::vector<uint32_t> start(g.ncells + 1, 0); std::vector<int> cid(n);
  for (int i = 0; i < n; i++) {
    int cx = cell_coord(map4[4*i], g.ox, g.inv_h), cy = cell_coord(map4[4*i+1], g.oy, g.inv_h), cz = cell_coord(map4[4*i+2], g.oz, g.inv_h);
    cid[i] = (cz * g.ny + cy) * g.nx + cx; start[cid[i] + 1]++;
  }
  for (int c = 0; c < g.ncells; c++) start[c+1] += start[c];
  std::vector<uint32_t> fill(start.begin(), start.end() - 1);
  std::vector<float4> sorted(n);
  for (int i = 0; i < n; i++) { float4 p = make_float4(map4[4*i], map4[4*i+1], map4[4*i+2], 0.f); memcpy(&p.w, &i, 4); sorted[fill[cid[i]]++] = p; }
  g.cell_start = start.data(); g.pts = sorted.data();
  for (int i = 0; i < nq; i++) {
    knn_key best[5];
    knn5_grid(g, q4[4*i], q4[4*i+1], q4[4*i+2], gate, best);
    for (int j = 0; j < 5; j++) {
      float d = knn_key_d(best[j]); int pos = knn_key_pos(best[j]); bool ok = pos >= 0 && d < gate;
      int oi = 0; if (ok) memcpy(&oi, &sorted[pos].w, 4);
      idx[5*i+j] = ok ? oi : -1; sqd[5*i+j] = ok ? d : FLT_MAX;
    }
  }
}
