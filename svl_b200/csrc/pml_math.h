// pml_math.h -- element matrices of the mixed displacement-stress PML (host side, FP64).
//
// Reproduces PML3DHexa8::Compute{Mass,Damping,Stiffness,PML}Matrix (04-Elements/10-Hexahedron/
// PML3DHexa8.cpp:214-285, 430-568, 288-427, 572-713; stretching :952-996) and
// PML2DQuad4::Compute{Mass,Stiffness,Damping}Matrix (04-Elements/06-Quadrilateral/PML2DQuad4.cpp:233-292,
// 296-377, 380-464; stretching :655-691).  Node state: 3-D [u1 u2 u3 s11 s22 s33 s12 s23 s13], 2-D
// [u1 u2 s11 s22 s12].
//
// Formulation used here: with a_c = alpha_c, b_c = beta_c the stretch functions along axis c, the scalar
// weights of the four matrices are the coefficients of the polynomial  prod_c (a_c + s b_c)  in s
// (s^0 -> M, s^1 -> C, s^2 -> K, s^3 -> G), and the weight of the u-sigma coupling along axis c is the
// coefficient one order lower of the same product with factor c left out.
#pragma once
#include <cmath>
#include "elem_math.h"

namespace svl {

template <int ND> struct PmlLayout;
template <> struct PmlLayout<3> {
    static constexpr int npe = 8, ndofn = 9, nmat = 4;
    // stress slot coupled to displacement component a through the derivative along c
    static int sidx(int a, int c) { return a == c ? 3 + a : (a + c == 1 ? 6 : (a + c == 3 ? 7 : 8)); }
};
template <> struct PmlLayout<2> {
    static constexpr int npe = 4, ndofn = 5, nmat = 3;
    static int sidx(int a, int c) { return a == c ? 2 + a : 4; }
};

// out[q] (q = 0 M, 1 C, 2 K, 3 G (3-D only)) row-major (npe*ndofn)^2, may be null.
// par: 3-D [n, L, R, x0(3), npml(3)], 2-D [th, n, L, R, x0(2), npml(2)]  (Driver.hpp:1288-1305, 1203-1219)
template <int ND>
inline void pml_element_matrices(const double *Xflat, double E, double nu, double rho, const double *par,
                                 double *const out[4]) {
    using Lay = PmlLayout<ND>;
    constexpr int npe = Lay::npe, nn = Lay::ndofn, n = npe * nn;
    for (int q = 0; q < Lay::nmat; q++)
        if (out[q]) for (int i = 0; i < n * n; i++) out[q][i] = 0.0;
    const double th = (ND == 2) ? par[0] : 1.0;
    const double *pp = (ND == 2) ? par + 1 : par;
    const double mexp = pp[0], L = pp[1], R = pp[2];
    const double *x0 = pp + 3, *npml = pp + 3 + ND;
    const double mu = E / (2.0 * (1.0 + nu));
    const double lambda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
    const double Vp = std::sqrt((lambda + 2.0 * mu) / rho);
    const double bref = L / 10.0;
    const double a0 = (mexp + 1.0) * bref / 2.0 / L * std::log(1.0 / R);
    const double b0 = (mexp + 1.0) * Vp / 2.0 / L * std::log(1.0 / R);
    // compliance of the stress block (normal-normal diagonal, normal-normal off-diagonal, shear)
    const double cn = (ND == 3) ? (lambda + mu) / mu / (3.0 * lambda + 2.0 * mu) : (lambda + 2.0 * mu) / 4.0 / mu / (lambda + mu);
    const double co = (ND == 3) ? lambda / 2.0 / mu / (3.0 * lambda + 2.0 * mu) : lambda / 4.0 / mu / (lambda + mu);
    const double cs = 1.0 / mu;
    double X[npe][ND];
    for (int i = 0; i < npe; i++) for (int c = 0; c < ND; c++) X[i][c] = Xflat[ND * i + c];
    const int ngp = (ND == 3) ? 8 : 4;
    for (int g = 0; g < ngp; g++) {
        double dN[npe][ND], N[npe], wdet;
        if constexpr (ND == 3) wdet = hex8_grad(X, g, dN, N);
        else wdet = th * quad4_grad(X, g, dN, N);
        double al[ND], be[ND];
        for (int c = 0; c < ND; c++) {
            double xg = 0.0;
            for (int i = 0; i < npe; i++) xg += N[i] * X[i][c];
            const double pw = std::pow((xg - x0[c]) * npml[c] / L, mexp);
            al[c] = 1.0 + a0 * pw;
            be[c] = b0 * pw;
        }
        // chi[q]: coefficients of prod_c (al_c + s be_c); phi[c][q-1]: same without factor c
        double chi[4] = {1.0, 0.0, 0.0, 0.0};
        for (int c = 0; c < ND; c++)
            for (int q = c + 1; q >= 0; q--) chi[q] = chi[q] * al[c] + (q > 0 ? chi[q - 1] * be[c] : 0.0);
        double phi[ND][4];
        for (int c = 0; c < ND; c++) {
            double pc[4] = {1.0, 0.0, 0.0, 0.0};
            int deg = 0;
            for (int d = 0; d < ND; d++) {
                if (d == c) continue;
                deg++;
                for (int q = deg; q >= 0; q--) pc[q] = pc[q] * al[d] + (q > 0 ? pc[q - 1] * be[d] : 0.0);
            }
            for (int q = 0; q < 4; q++) phi[c][q] = pc[q];
        }
        for (int q = 0; q < Lay::nmat; q++) {
            double *A = out[q];
            if (!A) continue;
            const double ch = chi[q];
            for (int j = 0; j < npe; j++)
                for (int k = 0; k < npe; k++) {
                    const double S = N[j] * N[k] * wdet;
                    double *B = A + (size_t)(nn * j) * n + nn * k;
                    for (int a = 0; a < ND; a++) B[a * n + a] += rho * ch * S;
                    for (int a = 0; a < ND; a++) {
                        B[(ND + a) * n + ND + a] += -ch * cn * S;
                        for (int b = 0; b < ND; b++) if (b != a) B[(ND + a) * n + ND + b] += ch * co * S;
                    }
                    for (int a = 2 * ND; a < nn; a++) B[a * n + a] += -ch * cs * S;
                    if (q == 0) continue;
                    for (int a = 0; a < ND; a++)
                        for (int c = 0; c < ND; c++) {
                            const int s = Lay::sidx(a, c);
                            B[a * n + s] += phi[c][q - 1] * dN[j][c] * N[k] * wdet;
                            B[s * n + a] += phi[c][q - 1] * dN[k][c] * N[j] * wdet;
                        }
                }
        }
    }
}

}  // namespace svl
