// elem_math.h -- element / material arithmetic shared by the host planner and the device
// kernels of libsvlgpu (FP64 everywhere).  Written from the reference's published element
// formulation; each routine names the reference code whose results it reproduces
// (paths relative to the reference's 02-Run_Process/).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define SVL_HD __host__ __device__ __forceinline__
#else
#define SVL_HD inline
#endif

namespace svl {

// 2-point Gauss abscissa exactly as tabulated by the reference (15-digit literal, weight 1,
// x fastest): 04-Elements/11-Integration/GaussQuadrature.cpp:311-316, 733-742.
constexpr double kGauss = 0.577350269189626;

// local node signs: lin3DHexa8.cpp:791-798 (VTK hexahedron), lin2DQuad4.cpp:673-676
SVL_HD double hx(int i) { return ((i + 1) & 2) ? 1.0 : -1.0; }          // -,+,+,-,-,+,+,-
SVL_HD double hy(int i) { return (i & 2) ? 1.0 : -1.0; }                // -,-,+,+,-,-,+,+
SVL_HD double hz(int i) { return (i & 4) ? 1.0 : -1.0; }                // -,-,-,-,+,+,+,+

// Physical shape-function gradients of the trilinear hex at Gauss point g.
// X[8][3] node coordinates; returns |det J|.  Reproduces ComputeJacobianMatrix +
// ComputeStrainDisplacementMatrix (lin3DHexa8.cpp:755-785, 810-853).
SVL_HD double hex8_grad(const double (*X)[3], int g, double (*dN)[3], double *N) {
    const double r = (g & 1) ? kGauss : -kGauss, s = (g & 2) ? kGauss : -kGauss,
                 t = (g & 4) ? kGauss : -kGauss;
    double dl[8][3];
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double a = 1.0 + hx(i) * r, b = 1.0 + hy(i) * s, c = 1.0 + hz(i) * t;
        dl[i][0] = 0.125 * hx(i) * b * c;
        dl[i][1] = 0.125 * hy(i) * a * c;
        dl[i][2] = 0.125 * hz(i) * a * b;
        if (N) N[i] = 0.125 * a * b * c;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = 0; q < 3; q++) J[p][q] += dl[i][p] * X[i][q];
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    double Ji[3][3];
    Ji[0][0] = c00 * id; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][0] = c01 * id; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][0] = c02 * id; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int q = 0; q < 3; q++)
            dN[i][q] = Ji[q][0] * dl[i][0] + Ji[q][1] * dl[i][1] + Ji[q][2] * dl[i][2];
    return fabs(det);
}

// Bilinear quad at Gauss point g (lin2DQuad4.cpp:648-711); returns |det J|.
SVL_HD double quad4_grad(const double (*X)[2], int g, double (*dN)[2], double *N) {
    const double r = (g & 1) ? kGauss : -kGauss, s = (g & 2) ? kGauss : -kGauss;
    double dl[4][2], J[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double qx = ((i + 1) & 2) ? 1.0 : -1.0, qy = (i & 2) ? 1.0 : -1.0;
        dl[i][0] = 0.25 * qx * (1.0 + qy * s);
        dl[i][1] = 0.25 * qy * (1.0 + qx * r);
        if (N) N[i] = 0.25 * (1.0 + qx * r) * (1.0 + qy * s);
        J[0][0] += dl[i][0] * X[i][0]; J[0][1] += dl[i][0] * X[i][1];
        J[1][0] += dl[i][1] * X[i][0]; J[1][1] += dl[i][1] * X[i][1];
    }
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    const double Ji[2][2] = {{J[1][1] * id, -J[0][1] * id}, {-J[1][0] * id, J[0][0] * id}};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        dN[i][0] = Ji[0][0] * dl[i][0] + Ji[0][1] * dl[i][1];
        dN[i][1] = Ji[1][0] * dl[i][0] + Ji[1][1] * dl[i][1];
    }
    return fabs(det);
}

// Isotropic elasticity constants: Elastic3DLinear.cpp:18-20 / Elastic2DPlaneStrain.cpp:18-20.
struct Iso { double c1, c2, c3; };
SVL_HD Iso iso_from_E_nu(double E, double nu) {
    Iso m;
    m.c1 = E * (1.0 - nu) / (1.0 - 2.0 * nu) / (1.0 + nu);
    m.c2 = E * nu / (1.0 - 2.0 * nu) / (1.0 + nu);
    m.c3 = E / (2.0 * (1.0 + nu));
    return m;
}
// sigma = C eps, Voigt [11,22,33,12,23,13], engineering shear (Elastic3DLinear.cpp:97-100)
SVL_HD void iso_stress3(const Iso &m, const double e[6], double s[6]) {
    s[0] = m.c1 * e[0] + m.c2 * e[1] + m.c2 * e[2];
    s[1] = m.c2 * e[0] + m.c1 * e[1] + m.c2 * e[2];
    s[2] = m.c2 * e[0] + m.c2 * e[1] + m.c1 * e[2];
    s[3] = m.c3 * e[3]; s[4] = m.c3 * e[4]; s[5] = m.c3 * e[5];
}
SVL_HD void iso_stress2(const Iso &m, const double e[3], double s[3]) {
    s[0] = m.c1 * e[0] + m.c2 * e[1];
    s[1] = m.c2 * e[0] + m.c1 * e[1];
    s[2] = m.c3 * e[2];
}

// J2 radial return with linear mixed hardening: Plastic3DJ2.cpp:206-259 (+ CommitState
// :163-170: the explicit path commits after every update).  st = eps_p[6] | q[6] | alpha.
struct J2Par { double K, G, H, beta, Sy; };
// returns true when the step was plastic (the state changed)
SVL_HD bool j2_return_map(const J2Par &p, const double ee[6], double st[13], double sig[6]) {
    double e[6] = {ee[0], ee[1], ee[2], 0.5 * ee[3], 0.5 * ee[4], 0.5 * ee[5]};
    const double tr = e[0] + e[1] + e[2];
    double str[6], xi[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const double dev = e[i] - ((i < 3) ? 1.0 / 3.0 * tr : 0.0);
        str[i] = 2.0 * p.G * (dev - st[i]);
        xi[i] = str[i] - st[6 + i];
    }
    const double nrm = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2] +
                            2.0 * (xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]));
    const double f = nrm - sqrt(2.0 / 3.0) * (p.Sy + st[12] * p.beta * p.H);
    const double kt = p.K * tr;
    if (f <= 0.0) {
#pragma unroll
        for (int i = 0; i < 6; i++) sig[i] = ((i < 3) ? kt : 0.0) + str[i];
        return false;
    } else {
        const double dg = f / (2.0 * p.G + 2.0 / 3.0 * p.H);
        st[12] += sqrt(2.0 / 3.0) * dg;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double n = xi[i] / nrm;
            st[6 + i] += 2.0 / 3.0 * (1.0 - p.beta) * p.H * dg * n;
            st[i] += dg * n;
            sig[i] = ((i < 3) ? kt : 0.0) + str[i] - 2.0 * p.G * dg * n;
        }
        return true;
    }
}

// ---- host-side class tables -------------------------------------------------------------
// K_e = sum_gp w |J| B^T C B (lin3DHexa8.cpp:288-318) for an isotropic C, row-major 24x24.
inline void hex8_stiffness(const double (*X)[3], const Iso &m, double *K) {
    for (int i = 0; i < 576; i++) K[i] = 0.0;
    for (int g = 0; g < 8; g++) {
        double d[8][3];
        const double w = hex8_grad(X, g, d, nullptr);
        for (int j = 0; j < 8; j++)
            for (int b = 0; b < 3; b++) {
                // strain of unit displacement (node j, comp b), then stress
                double e[6] = {0, 0, 0, 0, 0, 0}, s[6];
                if (b == 0) { e[0] = d[j][0]; e[3] = d[j][1]; e[5] = d[j][2]; }
                if (b == 1) { e[1] = d[j][1]; e[3] = d[j][0]; e[4] = d[j][2]; }
                if (b == 2) { e[2] = d[j][2]; e[4] = d[j][1]; e[5] = d[j][0]; }
                iso_stress3(m, e, s);
                for (int i = 0; i < 8; i++) {
                    K[(3 * i + 0) * 24 + 3 * j + b] += w * (d[i][0] * s[0] + d[i][1] * s[3] + d[i][2] * s[5]);
                    K[(3 * i + 1) * 24 + 3 * j + b] += w * (d[i][1] * s[1] + d[i][0] * s[3] + d[i][2] * s[4]);
                    K[(3 * i + 2) * 24 + 3 * j + b] += w * (d[i][2] * s[2] + d[i][1] * s[4] + d[i][0] * s[5]);
                }
            }
    }
}
// Element mass (lin3DHexa8.cpp:242-285): consistent 8x8 scalar block m[i][j] = sum w rho |J| N_i N_j
// (identical for the 3 components); lumped = row sums.
inline void hex8_mass_nodes(const double (*X)[3], double rho, double m[8][8]) {
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) m[i][j] = 0.0;
    for (int g = 0; g < 8; g++) {
        double d[8][3], N[8];
        const double w = rho * hex8_grad(X, g, d, N);
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) m[i][j] += w * N[i] * N[j];
    }
}
inline void quad4_stiffness(const double (*X)[2], double th, const Iso &m, double *K) {
    for (int i = 0; i < 64; i++) K[i] = 0.0;
    for (int g = 0; g < 4; g++) {
        double d[4][2];
        const double w = th * quad4_grad(X, g, d, nullptr);
        for (int j = 0; j < 4; j++)
            for (int b = 0; b < 2; b++) {
                double e[3] = {0, 0, 0}, s[3];
                if (b == 0) { e[0] = d[j][0]; e[2] = d[j][1]; }
                if (b == 1) { e[1] = d[j][1]; e[2] = d[j][0]; }
                iso_stress2(m, e, s);
                for (int i = 0; i < 4; i++) {
                    K[(2 * i + 0) * 8 + 2 * j + b] += w * (d[i][0] * s[0] + d[i][1] * s[2]);
                    K[(2 * i + 1) * 8 + 2 * j + b] += w * (d[i][1] * s[1] + d[i][0] * s[2]);
                }
            }
    }
}
inline void quad4_mass_nodes(const double (*X)[2], double th, double rho, double m[4][4]) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = 0.0;
    for (int g = 0; g < 4; g++) {
        double d[4][2], N[4];
        const double w = rho * th * quad4_grad(X, g, d, N);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] += w * N[i] * N[j];
    }
}

}  // namespace svl
