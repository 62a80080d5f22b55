// D3Q15 lattice and the moment-space two-relaxation-time ("BGK") collision.
//
// Replaces (B200-native, from scratch):
//   /root/reference/include/d3q15_bgk.h:11-29     lattice vectors and weights
//   /root/reference/include/velocity_set.h:57-108 ND/NV/HALO/CS2
//   /root/reference/src/AmrSim.cpp:25-107         Collide (and :487-580 CoarseCollide body)
//   /root/reference/src/AmrSim.cpp:879-927        equilibrium
//   /root/reference/src/AmrSim.cpp:957-971        density / velocity moments
//   /root/reference/src/AmrSim.cpp:1033-1073      DELTA, MODE_MATRIX, MODE_MATRIX_INVERSE
//
// Two collision implementations share one interface:
//   CollideFast    -- the product path.  Exploits the structure of the moment
//                     basis (pair sums/differences, only the 10 hydrodynamic
//                     rows, one reciprocal): ~140 fp64 ops per cell instead of
//                     ~900 for the dense 15x15 products, so the kernel stays
//                     HBM-bound.  Rounding differs from the reference at the
//                     1e-15 level (reassociation + FMA contraction only).
//   CollideLiteral -- the reference's operation order term by term.  Compiled
//                     in a translation unit built with --fmad=false it is
//                     bit-identical to a non-FMA CPU build; used for parity
//                     checks (LBX_COLLIDE_LITERAL).
#pragma once
#include <cstdint>

namespace lbx {

constexpr int NV = 15;   // velocities            (velocity_set.h NV)
constexpr int ND = 3;    // dimensions            (velocity_set.h ND)
constexpr int HALO = 2;  // ghost width of DistFn (d3q15_bgk.h:11)

// lattice vectors: 0 rest; 1..6 = +-x,+-y,+-z; 7..14 = (+-1,+-1,+-1)
__host__ __device__ constexpr int cx(int p) { return p == 1 ? 1 : p == 2 ? -1 : p < 7 ? 0 : (p < 11 ? 1 : -1); }
__host__ __device__ constexpr int cy(int p) { return p == 3 ? 1 : p == 4 ? -1 : p < 7 ? 0 : (((p - 7) >> 1) & 1 ? -1 : 1); }
__host__ __device__ constexpr int cz(int p) { return p == 5 ? 1 : p == 6 ? -1 : p < 7 ? 0 : ((p - 7) & 1 ? -1 : 1); }

// weights as rationals: 2/9, 1/9 x6, 1/72 x8
__host__ __device__ constexpr int w_num(int p) { return p == 0 ? 2 : 1; }
__host__ __device__ constexpr int w_den(int p) { return p < 7 ? 9 : 72; }

// ---- moment basis as exact rationals (row m, velocity p) --------------------
// row 0: 1; rows 1..3: c_a; rows 4..9: c_a c_b - delta_ab/3 (xx,xy,xz,yy,yz,zz);
// row 10: -2 rest / 1 axis / -2 diagonal; rows 11..13: row10*c_a; row 14: cx cy cz
__host__ __device__ constexpr int cc(int a, int p) { return a == 0 ? cx(p) : a == 1 ? cy(p) : cz(p); }
__host__ __device__ constexpr int g10(int p) { return (p >= 1 && p <= 6) ? 1 : -2; }
__host__ __device__ constexpr int qa(int m) { return m == 4 || m == 5 || m == 6 ? 0 : (m == 7 || m == 8 ? 1 : 2); }
__host__ __device__ constexpr int qb(int m) { return m == 4 ? 0 : (m == 5 || m == 7 ? 1 : 2); }
__host__ __device__ constexpr int m_den(int m, int) { return (m == 4 || m == 7 || m == 9) ? 3 : 1; }
__host__ __device__ constexpr int m_num(int m, int p) {
  return m == 0 ? 1
       : m <= 3 ? cc(m - 1, p)
       : m <= 9 ? ((m == 4 || m == 7 || m == 9) ? 3 * cc(qa(m), p) * cc(qb(m), p) - 1
                                                 : cc(qa(m), p) * cc(qb(m), p))
       : m == 10 ? g10(p)
       : m <= 13 ? g10(p) * cc(m - 11, p)
       : cx(p) * cy(p) * cz(p);
}
// norms N_m = sum_p w_p M[m][p]^2 = {1, 1/3 x3, 2/9, 1/9, 1/9, 2/9, 1/9, 2/9, 2, 2/3 x3, 1/9}
__host__ __device__ constexpr int n_num(int m) { return (m == 4 || m == 7 || m == 9 || m == 10 || (m >= 11 && m <= 13)) ? 2 : 1; }
__host__ __device__ constexpr int n_den(int m) {
  return m == 0 ? 1 : m <= 3 ? 3 : (m == 4 || m == 7 || m == 9) ? 9 : m <= 9 ? 9 : m == 10 ? 1 : m <= 13 ? 3 : 9;
}
// M[m][p] and Minv[p][m] = w_p M[m][p] / N_m, each ONE correctly rounded division
__host__ __device__ constexpr double mode_entry(int m, int p) { return (double)m_num(m, p) / (double)m_den(m, p); }
__host__ __device__ constexpr double inv_entry(int p, int m) {
  return (double)(w_num(p) * m_num(m, p) * n_den(m)) / (double)(w_den(p) * m_den(m, p) * n_num(m));
}

// ---- equilibrium (src/AmrSim.cpp:879-927), same expression tree ------------
__device__ __forceinline__ void equilibrium_cell(double rho, double ux, double uy, double uz, double* f) {
  const double CS2 = 1.0 / 3.0;
  const double rw0 = rho * 2.0 / 9.0, rw1 = rho / 9.0, rw2 = rho / 72.0;
  const double u2x = ux * ux, u2y = uy * uy, u2z = uz * uz;
  const double ucx = ux / CS2, ucy = uy / CS2, ucz = uz / CS2;
  const double qx = u2x / (2.0 * CS2 * CS2), qy = u2y / (2.0 * CS2 * CS2), qz = u2z / (2.0 * CS2 * CS2);
  const double uv = ucx * ucy, vw = ucy * ucz, uw = ucx * ucz;
  const double ms = (u2x + u2y + u2z) / (2.0 * CS2);
  const double ms2 = (u2x + u2y + u2z) * (1 - CS2) / (2.0 * CS2 * CS2);
  f[0] = rw0 * (1.0 - ms);
  f[1] = rw1 * (1.0 - ms + ucx + qx);
  f[2] = rw1 * (1.0 - ms - ucx + qx);
  f[3] = rw1 * (1.0 - ms + ucy + qy);
  f[4] = rw1 * (1.0 - ms - ucy + qy);
  f[5] = rw1 * (1.0 - ms + ucz + qz);
  f[6] = rw1 * (1.0 - ms - ucz + qz);
  f[7] = rw2 * (1.0 + ucx + ucy + ucz + uv + vw + uw + ms2);
  f[8] = rw2 * (1.0 + ucx + ucy - ucz + uv - vw - uw + ms2);
  f[9] = rw2 * (1.0 + ucx - ucy + ucz - uv - vw + uw + ms2);
  f[10] = rw2 * (1.0 + ucx - ucy - ucz - uv + vw - uw + ms2);
  f[11] = rw2 * (1.0 - ucx + ucy + ucz - uv + vw - uw + ms2);
  f[12] = rw2 * (1.0 - ucx + ucy - ucz - uv - vw + uw + ms2);
  f[13] = rw2 * (1.0 - ucx - ucy + ucz + uv - vw - uw + ms2);
  f[14] = rw2 * (1.0 - ucx - ucy - ucz + uv + vw + uw + ms2);
}

// ---- literal collision / moments -------------------------------------------
// Term-by-term restatement; structural zeros are skipped (adding +-0 never
// changes an accumulator that started at +0) and rows 10..14 are not formed
// because they are overwritten with 0 before use (src/AmrSim.cpp:90-94).
struct CollideLiteral {
  __device__ __forceinline__ static void moments(const double* f, double& rho, double& ux, double& uy, double& uz) {
    double mode[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
#pragma unroll
      for (int p = 0; p < NV; ++p)
        if (m_num(m, p) != 0) acc += f[p] * mode_entry(m, p);
      mode[m] = acc;
    }
    rho = mode[0];
    ux = mode[1] / mode[0];
    uy = mode[2] / mode[0];
    uz = mode[3] / mode[0];
  }

  __device__ __forceinline__ static void collide(double* f, double omega_s, double omega_b) {
    double mode[10];
#pragma unroll
    for (int m = 0; m < 10; ++m) {
      double acc = 0.0;
#pragma unroll
      for (int p = 0; p < NV; ++p)
        if (m_num(m, p) != 0) acc += f[p] * mode_entry(m, p);
      mode[m] = acc;
    }
    const double rho = mode[0];
    double v[3], usq = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      v[a] = mode[a + 1] / rho;
      usq += v[a] * v[a];
    }
    double S[3][3] = {{mode[4], mode[5], mode[6]}, {mode[5], mode[7], mode[8]}, {mode[6], mode[8], mode[9]}};
    double TrS = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) TrS += S[a][a];
#pragma unroll
    for (int a = 0; a < 3; ++a) S[a][a] -= (TrS / 3);
    TrS -= omega_b * (TrS - rho * usq);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double dlt = (a == b) ? 1.0 / NV : 0.0;   // sic: 1/NMODES (src/AmrSim.cpp:1033)
        S[a][b] -= omega_s * (S[a][b] - rho * (v[a] * v[b] - usq * dlt));
      }
      S[a][a] += (TrS / 3);
    }
    mode[4] = S[0][0]; mode[5] = S[0][1]; mode[6] = S[0][2];
    mode[7] = S[1][1]; mode[8] = S[1][2]; mode[9] = S[2][2];
#pragma unroll
    for (int p = 0; p < NV; ++p) {
      double fp = 0.0;
#pragma unroll
      for (int m = 0; m < 10; ++m)
        if (m_num(m, p) != 0) fp += mode[m] * inv_entry(p, m);
      f[p] = fp;
    }
  }
};

// ---- fast collision / moments ----------------------------------------------
struct CollideFast {
  __device__ __forceinline__ static void moments(const double* f, double& rho, double& ux, double& uy, double& uz) {
    const double a = f[7] + f[8], b = f[9] + f[10], c = f[11] + f[12], d = f[13] + f[14];
    const double a_ = f[7] - f[8], b_ = f[9] - f[10], c_ = f[11] - f[12], d_ = f[13] - f[14];
    const double ab = a + b, cd = c + d;
    rho = f[0] + (((f[1] + f[2]) + (f[3] + f[4])) + (f[5] + f[6])) + (ab + cd);
    const double jx = (f[1] - f[2]) + (ab - cd);
    const double jy = (f[3] - f[4]) + ((a + c) - (b + d));
    const double jz = (f[5] - f[6]) + ((a_ + b_) + (c_ + d_));
    // the reference divides each component by rho (src/AmrSim.cpp:970); keep true
    // divisions here -- this kernel is far from compute-bound
    ux = jx / rho;
    uy = jy / rho;
    uz = jz / rho;
  }

  __device__ __forceinline__ static void collide(double* f, double omega_s, double omega_b) {
    // pair sums / differences
    const double s12 = f[1] + f[2], d12 = f[1] - f[2];
    const double s34 = f[3] + f[4], d34 = f[3] - f[4];
    const double s56 = f[5] + f[6], d56 = f[5] - f[6];
    const double a = f[7] + f[8], a_ = f[7] - f[8];       // (+,+,*)
    const double b = f[9] + f[10], b_ = f[9] - f[10];     // (+,-,*)
    const double c = f[11] + f[12], c_ = f[11] - f[12];   // (-,+,*)
    const double d = f[13] + f[14], d_ = f[13] - f[14];   // (-,-,*)
    const double ab = a + b, cd = c + d, sd = ab + cd;
    const double apb_ = a_ + b_, cpd_ = c_ + d_;
    // conserved moments (rows 0..3)
    const double rho = f[0] + ((s12 + s34) + s56) + sd;
    const double jx = d12 + (ab - cd);
    const double jy = d34 + ((a + c) - (b + d));
    const double jz = d56 + (apb_ + cpd_);
    // second moments sum_p c_a c_b f_p; rows 4..9 are these minus rho/3 on the diagonal
    const double pxx = s12 + sd, pyy = s34 + sd, pzz = s56 + sd;
    double sxy = (a + d) - (b + c);
    double sxz = apb_ - cpd_;
    double syz = (a_ - b_) + (c_ - d_);
    const double third = 1.0 / 3.0;
    const double P = (pxx + pyy) + pzz;
    double TrS = P - rho;                    // m4 + m7 + m9
    const double Pt = P * third;
    double sxx = pxx - Pt, syy = pyy - Pt, szz = pzz - Pt;   // traceless diagonal
    const double ir = 1.0 / rho;
    const double ux = jx * ir, uy = jy * ir, uz = jz * ir;
    const double usq = (ux * ux + uy * uy) + uz * uz;
    // relax the trace (omega_b) and the traceless part (omega_s); the reference's
    // DELTA is diag(1/15), not 1/3 (src/AmrSim.cpp:1033-1035) -- reproduced
    TrS -= omega_b * (TrS - rho * usq);
    const double du = usq * (1.0 / NV);
    sxx -= omega_s * (sxx - rho * (ux * ux - du));
    syy -= omega_s * (syy - rho * (uy * uy - du));
    szz -= omega_s * (szz - rho * (uz * uz - du));
    sxy -= omega_s * (sxy - jx * uy);
    sxz -= omega_s * (sxz - jx * uz);
    syz -= omega_s * (syz - jy * uz);
    const double t3 = TrS * third;
    sxx += t3; syy += t3; szz += t3;        // = post-collision m4, m7, m9
    const double T = (sxx + syy) + szz;
    // back-projection with ghost modes = 0 (columns 0..9 of the inverse)
    const double r9 = rho * (1.0 / 9.0), T6 = T * (1.0 / 6.0);
    f[0] = 2.0 * r9 - 2.0 * T6;
    const double ax = r9 + (0.5 * sxx - T6), ay = r9 + (0.5 * syy - T6), az = r9 + (0.5 * szz - T6);
    const double hx = jx * third, hy = jy * third, hz = jz * third;
    f[1] = ax + hx; f[2] = ax - hx;
    f[3] = ay + hy; f[4] = ay - hy;
    f[5] = az + hz; f[6] = az - hz;
    const double base = (r9 + T * third) * 0.125;          // rho/72 + T/24
    const double A = hx * 0.125, B = hy * 0.125, C = hz * 0.125;
    const double X = sxy * 0.125, Y = sxz * 0.125, Z = syz * 0.125;
    const double bpX = base + X, bmX = base - X, ApB = A + B, AmB = A - B;
    const double e_pp = bpX + ApB, e_mm = bpX - ApB, e_pm = bmX + AmB, e_mp = bmX - AmB;
    const double YpZ = Y + Z, YmZ = Y - Z;
    const double g_pp = C + YpZ, g_mm = C - YpZ, g_pm = C + YmZ, g_mp = C - YmZ;
    f[7] = e_pp + g_pp;  f[8] = e_pp - g_pp;
    f[9] = e_pm + g_pm;  f[10] = e_pm - g_pm;
    f[11] = e_mp + g_mp; f[12] = e_mp - g_mp;
    f[13] = e_mm + g_mm; f[14] = e_mm - g_mm;
  }
};

}  // namespace lbx
