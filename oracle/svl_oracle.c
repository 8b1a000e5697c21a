/*
 * svl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see svl_oracle.h).
 *
 * CPU restatement of SeismoVLAB/SVL's explicit hot path.  Citations are
 * relative to /root/reference/02-Run_Process/.  Compile with
 * -ffp-contract=off so that no FMA contraction changes the rounding relative
 * to the reference's plain C++ arithmetic.
 */
#include "svl_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* 04-Elements/11-Integration/GaussQuadrature.cpp:311-316, 733-742:
 * 15-digit literal, weights 1, x fastest.                                      */
#define GP 0.577350269189626

static const double HX[8] = {-1, 1, 1, -1, -1, 1, 1, -1};   /* lin3DHexa8.cpp:791-798 */
static const double HY[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
static const double HZ[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
static const double QX[4] = {-1, 1, 1, -1};                  /* lin2DQuad4.cpp:673-676 */
static const double QY[4] = {-1, -1, 1, 1};

/* ------------------------------------------------------------------------ */
/* materials                                                                 */
/* ------------------------------------------------------------------------ */
/* 02-Materials/01-Linear/Elastic3DLinear.cpp:18-27 */
void svlo_elastic3d_C(double E, double nu, double C[36]) {
    double c1 = E * (1.0 - nu) / (1.0 - 2.0 * nu) / (1.0 + nu);
    double c2 = E * nu / (1.0 - 2.0 * nu) / (1.0 + nu);
    double c3 = E / (2.0 * (1.0 + nu));
    memset(C, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[6 * i + j] = (i == j) ? c1 : c2;
    C[21] = C[28] = C[35] = c3;
}
/* 02-Materials/01-Linear/Elastic2DPlaneStrain.cpp:18-26 */
void svlo_planestrain_C(double E, double nu, double C[9]) {
    double c1 = E * (1.0 - nu) / (1.0 - 2.0 * nu) / (1.0 + nu);
    double c2 = E * nu / (1.0 - 2.0 * nu) / (1.0 + nu);
    double c3 = E / (2.0 * (1.0 + nu));
    double t[9] = {c1, c2, 0, c2, c1, 0, 0, 0, c3};
    memcpy(C, t, sizeof t);
}

/* 02-Materials/02-NonLinear/Plastic3DJ2.cpp:206-259 (UpdateState cond==1) and
 * :163-170 (CommitState).  state = eps_p(6) | backstress(6) | alpha.          */
void svlo_j2_update(const double par[6], const double e_in[6], double st[13], double sig[6]) {
    const double K = par[0], G = par[1], H = par[3], beta = par[4], Sy = par[5];
    double e[6] = {e_in[0], e_in[1], e_in[2], 0.5 * e_in[3], 0.5 * e_in[4], 0.5 * e_in[5]};
    double tr = e[0] + e[1] + e[2];
    double one[6] = {1, 1, 1, 0, 0, 0};
    double s_tr[6], xi[6];
    for (int i = 0; i < 6; i++) {
        double dev = e[i] - 1.0 / 3.0 * tr * one[i];
        s_tr[i] = 2.0 * G * (dev - st[i]);
        xi[i] = s_tr[i] - st[6 + i];
    }
    double nrm = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2] +
                      2.0 * (xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]));   /* :326-329 */
    double f = nrm - sqrt(2.0 / 3.0) * (Sy + st[12] * beta * H);
    if (f <= 0) {
        for (int i = 0; i < 6; i++) sig[i] = K * tr * one[i] + s_tr[i];
    } else {
        double dg = f / (2.0 * G + 2.0 / 3.0 * H);
        st[12] += sqrt(2.0 / 3.0) * dg;
        for (int i = 0; i < 6; i++) {
            double n = xi[i] / nrm;
            st[6 + i] += 2.0 / 3.0 * (1.0 - beta) * H * dg * n;
            st[i] += dg * n;
            sig[i] = K * tr * one[i] + s_tr[i] - 2.0 * G * dg * n;
        }
    }
}

/* Same update plus the consistent tangent (Plastic3DJ2.cpp:221-224 elastic, :254-256 plastic; identical expressions in
 * PlasticPlaneStrainJ2.cpp:245,262): D = 1 (x) 1 on the normal components, I = diag(1,1,1,1/2,1/2,1/2), Id = I - D/3,
 *   elastic: K D + 2G Id;   plastic: K D + 2G (Id - n n^T / (1 + H/(3G))) - 4 G^2 dgamma / |xi| (Id - n n^T)              */
void svlo_j2_update_tangent(const double par[6], const double e_in[6], double st[13], double sig[6], double Ct[36]) {
    const double K = par[0], G = par[1], H = par[3], beta = par[4], Sy = par[5];
    double e[6] = {e_in[0], e_in[1], e_in[2], 0.5 * e_in[3], 0.5 * e_in[4], 0.5 * e_in[5]};
    double tr = e[0] + e[1] + e[2];
    double one[6] = {1, 1, 1, 0, 0, 0};
    double s_tr[6], xi[6], D[36], Id[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            D[6 * i + j] = one[i] * one[j];
            Id[6 * i + j] = (i == j ? (i < 3 ? 1.0 : 0.5) : 0.0) - 1.0 / 3.0 * D[6 * i + j];
        }
    for (int i = 0; i < 6; i++) {
        double dev = e[i] - 1.0 / 3.0 * tr * one[i];
        s_tr[i] = 2.0 * G * (dev - st[i]);
        xi[i] = s_tr[i] - st[6 + i];
    }
    double nrm = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2] +
                      2.0 * (xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]));
    double f = nrm - sqrt(2.0 / 3.0) * (Sy + st[12] * beta * H);
    if (f <= 0) {
        for (int i = 0; i < 6; i++) sig[i] = K * tr * one[i] + s_tr[i];
        for (int i = 0; i < 36; i++) Ct[i] = K * D[i] + 2.0 * G * Id[i];
    } else {
        double dg = f / (2.0 * G + 2.0 / 3.0 * H), n[6];
        st[12] += sqrt(2.0 / 3.0) * dg;
        for (int i = 0; i < 6; i++) {
            n[i] = xi[i] / nrm;
            st[6 + i] += 2.0 / 3.0 * (1.0 - beta) * H * dg * n[i];
            st[i] += dg * n[i];
            sig[i] = K * tr * one[i] + s_tr[i] - 2.0 * G * dg * n[i];
        }
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++)
                Ct[6 * i + j] = K * D[6 * i + j] + 2.0 * G * (Id[6 * i + j] - 1.0 / (1.0 + H / G / 3.0) * n[i] * n[j]) -
                                4.0 * G * G * dg / nrm * (Id[6 * i + j] - n[i] * n[j]);
    }
}

/* ------------------------------------------------------------------------ */
/* hex8 kinematics: lin3DHexa8.cpp:755-785 (J), :810-853 (B)                 */
/* ------------------------------------------------------------------------ */
static void hex8_dN_local(double r, double s, double t, double dN[8][3], double N[8]) {
    for (int i = 0; i < 8; i++) {
        double a = 1.0 + HX[i] * r, b = 1.0 + HY[i] * s, c = 1.0 + HZ[i] * t;
        dN[i][0] = 1.0 / 8.0 * HX[i] * b * c;
        dN[i][1] = 1.0 / 8.0 * HY[i] * a * c;
        dN[i][2] = 1.0 / 8.0 * HZ[i] * a * b;
        if (N) N[i] = 1.0 / 8.0 * a * b * c;
    }
}
static double inv3(const double J[9], double Ji[9]) {
    double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    double id = 1.0 / det;
    Ji[0] = c00 * id; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * id; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * id;
    Ji[3] = c01 * id; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * id; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * id;
    Ji[6] = c02 * id; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * id; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * id;
    return det;
}
/* dNdx[i][b] = d N_i / d x_b at (r,s,t); returns |det J|; N optional           */
static double hex8_grad(const double *X, double r, double s, double t, double dNdx[8][3], double N[8]) {
    double dN[8][3], J[9] = {0}, Ji[9];
    hex8_dN_local(r, s, t, dN, N);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            double v = 0;
            for (int i = 0; i < 8; i++) v += dN[i][a] * X[3 * i + b];
            J[3 * a + b] = v;
        }
    double det = inv3(J, Ji);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 3; b++)
            dNdx[i][b] = Ji[3 * b + 2] * dN[i][2] + Ji[3 * b + 1] * dN[i][1] + Ji[3 * b + 0] * dN[i][0];
    return fabs(det);
}
static void hex8_gp(int g, double *r, double *s, double *t) {
    *r = (g & 1) ? GP : -GP; *s = (g & 2) ? GP : -GP; *t = (g & 4) ? GP : -GP;
}
/* strain Voigt [11,22,33,12,23,13] engineering shear: lin3DHexa8.cpp:845-850, :721-741 */
void svlo_hex8_strain(const double *X, const double *U, double eps[8][6]) {
    for (int g = 0; g < 8; g++) {
        double r, s, t, d[8][3];
        hex8_gp(g, &r, &s, &t);
        hex8_grad(X, r, s, t, d, NULL);
        double e[6] = {0};
        for (int i = 0; i < 8; i++) {
            const double *u = U + 3 * i;
            e[0] += d[i][0] * u[0];
            e[1] += d[i][1] * u[1];
            e[2] += d[i][2] * u[2];
            e[3] += d[i][1] * u[0] + d[i][0] * u[1];
            e[4] += d[i][2] * u[1] + d[i][1] * u[2];
            e[5] += d[i][2] * u[0] + d[i][0] * u[2];
        }
        memcpy(eps[g], e, sizeof e);
    }
}
/* f = sum_gp w |J| B^T sigma : lin3DHexa8.cpp:382-412 */
void svlo_hex8_force(const double *X, const double sig[8][6], double f[24]) {
    memset(f, 0, 24 * sizeof(double));
    for (int g = 0; g < 8; g++) {
        double r, s, t, d[8][3];
        hex8_gp(g, &r, &s, &t);
        double wd = 1.0 * hex8_grad(X, r, s, t, d, NULL);
        const double *sg = sig[g];
        for (int i = 0; i < 8; i++) {
            f[3 * i + 0] += wd * (d[i][0] * sg[0] + d[i][1] * sg[3] + d[i][2] * sg[5]);
            f[3 * i + 1] += wd * (d[i][1] * sg[1] + d[i][0] * sg[3] + d[i][2] * sg[4]);
            f[3 * i + 2] += wd * (d[i][2] * sg[2] + d[i][1] * sg[4] + d[i][0] * sg[5]);
        }
    }
}
/* lin3DHexa8.cpp:242-285: consistent mass then optional row-sum lumping        */
void svlo_hex8_mass(const double *X, double rho, int lumped, double M[576]) {
    memset(M, 0, 576 * sizeof(double));
    for (int g = 0; g < 8; g++) {
        double r, s, t, d[8][3], N[8];
        hex8_gp(g, &r, &s, &t);
        double wd = 1.0 * rho * hex8_grad(X, r, s, t, d, N);
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++)
                for (int c = 0; c < 3; c++) M[(3 * i + c) * 24 + 3 * j + c] += wd * N[i] * N[j];
    }
    if (lumped)
        for (int i = 0; i < 24; i++)
            for (int j = 0; j < 24; j++)
                if (i != j) { M[i * 24 + i] += M[i * 24 + j]; M[i * 24 + j] = 0.0; }
}
static void hex8_B(const double d[8][3], double B[6][24]) {
    memset(B, 0, 6 * 24 * sizeof(double));
    for (int i = 0; i < 8; i++) {
        B[0][3 * i] = d[i][0]; B[1][3 * i + 1] = d[i][1]; B[2][3 * i + 2] = d[i][2];
        B[3][3 * i] = d[i][1]; B[3][3 * i + 1] = d[i][0];
        B[4][3 * i + 1] = d[i][2]; B[4][3 * i + 2] = d[i][1];
        B[5][3 * i] = d[i][2]; B[5][3 * i + 2] = d[i][0];
    }
}
/* lin3DHexa8.cpp:288-318 */
static void hex8_stiffness_gp(const double *X, const double *Cgp, int stride, double K[576]);
void svlo_hex8_stiffness(const double *X, const double C[36], double K[576]) { hex8_stiffness_gp(X, C, 0, K); }
/* K = sum_gp w |J| B^T C_gp B with one tangent per Gauss point (stride 36) or one for all (stride 0): lin3DHexa8.cpp:288-318 */
static void hex8_stiffness_gp(const double *X, const double *Cgp, int stride, double K[576]) {
    memset(K, 0, 576 * sizeof(double));
    for (int g = 0; g < 8; g++) {
        const double *C = Cgp + (size_t)stride * g;
        double r, s, t, d[8][3], B[6][24], CB[6][24];
        hex8_gp(g, &r, &s, &t);
        double wd = 1.0 * hex8_grad(X, r, s, t, d, NULL);
        hex8_B(d, B);
        for (int a = 0; a < 6; a++)
            for (int j = 0; j < 24; j++) {
                double v = 0;
                for (int b = 0; b < 6; b++) v += C[6 * a + b] * B[b][j];
                CB[a][j] = v;
            }
        for (int i = 0; i < 24; i++)
            for (int j = 0; j < 24; j++) {
                double v = 0;
                for (int a = 0; a < 6; a++) v += B[a][i] * CB[a][j];
                K[i * 24 + j] += wd * v;
            }
    }
}

/* ------------------------------------------------------------------------ */
/* quad4 kinematics: lin2DQuad4.cpp:648-711                                  */
/* ------------------------------------------------------------------------ */
static double quad4_grad(const double *X, double r, double s, double dNdx[4][2], double N[4]) {
    double dN[4][2], J[4] = {0};
    for (int i = 0; i < 4; i++) {
        dN[i][0] = 1.0 / 4.0 * QX[i] * (1.0 + QY[i] * s);
        dN[i][1] = 1.0 / 4.0 * QY[i] * (1.0 + QX[i] * r);
        if (N) N[i] = 1.0 / 4.0 * (1.0 + QX[i] * r) * (1.0 + QY[i] * s);
    }
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            double v = 0;
            for (int i = 0; i < 4; i++) v += dN[i][a] * X[2 * i + b];
            J[2 * a + b] = v;
        }
    double det = J[0] * J[3] - J[1] * J[2];
    double Ji[4] = {J[3] / det, -J[1] / det, -J[2] / det, J[0] / det};
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 2; b++) dNdx[i][b] = Ji[2 * b + 1] * dN[i][1] + Ji[2 * b + 0] * dN[i][0];
    return fabs(det);
}
static void quad4_gp(int g, double *r, double *s) { *r = (g & 1) ? GP : -GP; *s = (g & 2) ? GP : -GP; }
/* Voigt [11,22,12]: lin2DQuad4.cpp:705-708 */
void svlo_quad4_strain(const double *X, const double *U, double eps[4][3]) {
    for (int g = 0; g < 4; g++) {
        double r, s, d[4][2];
        quad4_gp(g, &r, &s);
        quad4_grad(X, r, s, d, NULL);
        double e[3] = {0};
        for (int i = 0; i < 4; i++) {
            e[0] += d[i][0] * U[2 * i];
            e[1] += d[i][1] * U[2 * i + 1];
            e[2] += d[i][1] * U[2 * i] + d[i][0] * U[2 * i + 1];
        }
        memcpy(eps[g], e, sizeof e);
    }
}
/* lin2DQuad4.cpp:383-412 */
void svlo_quad4_force(const double *X, double th, const double sig[4][3], double f[8]) {
    memset(f, 0, 8 * sizeof(double));
    for (int g = 0; g < 4; g++) {
        double r, s, d[4][2];
        quad4_gp(g, &r, &s);
        double wd = 1.0 * th * quad4_grad(X, r, s, d, NULL);
        for (int i = 0; i < 4; i++) {
            f[2 * i + 0] += wd * (d[i][0] * sig[g][0] + d[i][1] * sig[g][2]);
            f[2 * i + 1] += wd * (d[i][1] * sig[g][1] + d[i][0] * sig[g][2]);
        }
    }
}
/* lin2DQuad4.cpp:243-286 */
void svlo_quad4_mass(const double *X, double th, double rho, int lumped, double M[64]) {
    memset(M, 0, 64 * sizeof(double));
    for (int g = 0; g < 4; g++) {
        double r, s, d[4][2], N[4];
        quad4_gp(g, &r, &s);
        double wd = 1.0 * rho * th * quad4_grad(X, r, s, d, N);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                for (int c = 0; c < 2; c++) M[(2 * i + c) * 8 + 2 * j + c] += wd * N[i] * N[j];
    }
    if (lumped)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++)
                if (i != j) { M[i * 8 + i] += M[i * 8 + j]; M[i * 8 + j] = 0.0; }
}
/* lin2DQuad4.cpp:289-319 */
static void quad4_stiffness_gp(const double *X, double th, const double *Cgp, int stride, double K[64]);
void svlo_quad4_stiffness(const double *X, double th, const double C[9], double K[64]) { quad4_stiffness_gp(X, th, C, 0, K); }
static void quad4_stiffness_gp(const double *X, double th, const double *Cgp, int stride, double K[64]) {
    memset(K, 0, 64 * sizeof(double));
    for (int g = 0; g < 4; g++) {
        const double *C = Cgp + (size_t)stride * g;
        double r, s, d[4][2], B[3][8] = {{0}}, CB[3][8];
        quad4_gp(g, &r, &s);
        double wd = 1.0 * th * quad4_grad(X, r, s, d, NULL);
        for (int i = 0; i < 4; i++) {
            B[0][2 * i] = d[i][0]; B[1][2 * i + 1] = d[i][1];
            B[2][2 * i] = d[i][1]; B[2][2 * i + 1] = d[i][0];
        }
        for (int a = 0; a < 3; a++)
            for (int j = 0; j < 8; j++) {
                double v = 0;
                for (int b = 0; b < 3; b++) v += C[3 * a + b] * B[b][j];
                CB[a][j] = v;
            }
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) {
                double v = 0;
                for (int a = 0; a < 3; a++) v += B[a][i] * CB[a][j];
                K[i * 8 + j] += wd * v;
            }
    }
}

/* ------------------------------------------------------------------------ */
/* PML matrices: PML3DHexa8.cpp:214-285 (M), 430-568 (C), 288-427 (K),        */
/* 572-713 (G), 952-996 (stretching);  PML2DQuad4.cpp:233-292, 296-377,       */
/* 380-464, 655-691.                                                          */
/* ------------------------------------------------------------------------ */
void svlo_pml3d_matrices(const double *X, double E, double nu, double rho, const double p[9],
                         double *M, double *C, double *K, double *G) {
    const int n = 72;
    double *out[4] = {M, C, K, G};
    for (int q = 0; q < 4; q++) if (out[q]) memset(out[q], 0, n * n * sizeof(double));
    const double mp = p[0], L = p[1], R = p[2], x0[3] = {p[3], p[4], p[5]}, np[3] = {p[6], p[7], p[8]};
    const double mu = E / (2.0 * (1.0 + nu));                 /* Elastic3DLinear::GetShearModulus */
    const double lambda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
    for (int g = 0; g < 8; g++) {
        double r, s, t, d[8][3], N[8];
        hex8_gp(g, &r, &s, &t);
        double D = hex8_grad(X, r, s, t, d, N);
        double xg[3] = {0, 0, 0};
        for (int i = 0; i < 8; i++) for (int c = 0; c < 3; c++) xg[c] += N[i] * X[3 * i + c];
        double Vp = sqrt((lambda + 2.0 * mu) / rho);
        double bp = L / 10.0;
        double a0 = (mp + 1.0) * bp / 2.0 / L * log(1.0 / R);
        double b0 = (mp + 1.0) * Vp / 2.0 / L * log(1.0 / R);
        double al[3], be[3];
        for (int c = 0; c < 3; c++) {
            double pw = pow((xg[c] - x0[c]) * np[c] / L, mp);
            al[c] = 1.0 + a0 * pw;
            be[c] = b0 * pw;
        }
        const double ax = al[0], ay = al[1], az = al[2], bx = be[0], by = be[1], bz = be[2];
        /* chi and phi_{x,y,z} for M, C, K, G */
        double chi[4] = {ax * ay * az, ax * ay * bz + ax * by * az + bx * ay * az,
                         ax * by * bz + bx * by * az + bx * ay * bz, bx * by * bz};
        double phi[4][3] = {{0, 0, 0},
                            {ay * az, ax * az, ax * ay},
                            {ay * bz + az * by, ax * bz + az * bx, ax * by + ay * bx},
                            {by * bz, bx * bz, bx * by}};
        for (int q = 0; q < 4; q++) {
            double *A = out[q];
            if (!A) continue;
            double ch = chi[q];
            double d1 = -ch * (lambda + mu) / mu / (3.0 * lambda + 2.0 * mu);
            double d2 = -ch / mu;
            double o1 = ch * lambda / 2.0 / mu / (3.0 * lambda + 2.0 * mu);
            for (int j = 0; j < 8; j++)
                for (int k = 0; k < 8; k++) {
                    double S = N[j] * N[k] * 1.0 * D;
                    double *B = A + (9 * j) * n + 9 * k;
#define AT(a, b) B[(a) * n + (b)]
                    AT(0, 0) += rho * ch * S; AT(1, 1) += rho * ch * S; AT(2, 2) += rho * ch * S;
                    AT(3, 3) += d1 * S; AT(4, 4) += d1 * S; AT(5, 5) += d1 * S;
                    AT(6, 6) += d2 * S; AT(7, 7) += d2 * S; AT(8, 8) += d2 * S;
                    AT(3, 4) += o1 * S; AT(3, 5) += o1 * S; AT(4, 5) += o1 * S;
                    AT(4, 3) += o1 * S; AT(5, 3) += o1 * S; AT(5, 4) += o1 * S;
                    if (q > 0) {
                        double wD = 1.0 * D;
                        double gxj = d[j][0] * N[k] * phi[q][0] * wD, gyj = d[j][1] * N[k] * phi[q][1] * wD,
                               gzj = d[j][2] * N[k] * phi[q][2] * wD;
                        double gxk = d[k][0] * N[j] * phi[q][0] * wD, gyk = d[k][1] * N[j] * phi[q][1] * wD,
                               gzk = d[k][2] * N[j] * phi[q][2] * wD;
                        AT(0, 3) += gxj; AT(0, 6) += gyj; AT(0, 8) += gzj;
                        AT(1, 4) += gyj; AT(1, 6) += gxj; AT(1, 7) += gzj;
                        AT(2, 5) += gzj; AT(2, 8) += gxj; AT(2, 7) += gyj;
                        AT(3, 0) += gxk; AT(6, 0) += gyk; AT(8, 0) += gzk;
                        AT(4, 1) += gyk; AT(6, 1) += gxk; AT(7, 1) += gzk;
                        AT(5, 2) += gzk; AT(8, 2) += gxk; AT(7, 2) += gyk;
                    }
#undef AT
                }
        }
    }
}

void svlo_pml2d_matrices(const double *X, double E, double nu, double rho, const double p[8],
                         double *M, double *C, double *K) {
    const int n = 20;
    double *out[3] = {M, C, K};
    for (int q = 0; q < 3; q++) if (out[q]) memset(out[q], 0, n * n * sizeof(double));
    const double th = p[0], mp = p[1], L = p[2], R = p[3], x0[2] = {p[4], p[5]}, np[2] = {p[6], p[7]};
    const double mu = E / (2.0 * (1.0 + nu));
    const double lambda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
    for (int g = 0; g < 4; g++) {
        double r, s, d[4][2], N[4];
        quad4_gp(g, &r, &s);
        double D = quad4_grad(X, r, s, d, N);
        double xg[2] = {0, 0};
        for (int i = 0; i < 4; i++) for (int c = 0; c < 2; c++) xg[c] += N[i] * X[2 * i + c];
        double Vp = sqrt((lambda + 2.0 * mu) / rho);
        double bp = L / 10.0;
        double a0 = (mp + 1.0) * bp / 2.0 / L * log(1.0 / R);
        double b0 = (mp + 1.0) * Vp / 2.0 / L * log(1.0 / R);
        double al[2], be[2];
        for (int c = 0; c < 2; c++) {
            double pw = pow((xg[c] - x0[c]) * np[c] / L, mp);
            al[c] = 1.0 + a0 * pw;
            be[c] = b0 * pw;
        }
        const double ax = al[0], ay = al[1], bx = be[0], by = be[1];
        double chi[3] = {ax * ay, ax * by + ay * bx, bx * by};
        double phx[3] = {0, ay, by}, phy[3] = {0, ax, bx};
        for (int q = 0; q < 3; q++) {
            double *A = out[q];
            if (!A) continue;
            double ch = chi[q];
            double d1 = -ch * (lambda + 2.0 * mu) / 4.0 / mu / (lambda + mu);
            double d2 = -ch / mu;
            double o1 = ch * lambda / 4.0 / mu / (lambda + mu);
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 4; k++) {
                    double S = N[j] * N[k] * th * 1.0 * D;
                    double *B = A + (5 * j) * n + 5 * k;
#define AT(a, b) B[(a) * n + (b)]
                    AT(0, 0) += rho * ch * S; AT(1, 1) += rho * ch * S;
                    AT(2, 2) += d1 * S; AT(3, 3) += d1 * S; AT(4, 4) += d2 * S;
                    AT(2, 3) += o1 * S; AT(3, 2) += o1 * S;
                    if (q > 0) {
                        double w = th * 1.0 * D;
                        AT(0, 2) += phx[q] * d[j][0] * N[k] * w; AT(2, 0) += phx[q] * d[k][0] * N[j] * w;
                        AT(0, 4) += phy[q] * d[j][1] * N[k] * w; AT(4, 0) += phy[q] * d[k][1] * N[j] * w;
                        AT(1, 3) += phy[q] * d[j][1] * N[k] * w; AT(3, 1) += phy[q] * d[k][1] * N[j] * w;
                        AT(1, 4) += phx[q] * d[j][0] * N[k] * w; AT(4, 1) += phx[q] * d[k][0] * N[j] * w;
                    }
#undef AT
                }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* DRM element forces: lin3DHexa8.cpp:660-718, lin2DQuad4.cpp:564-614         */
/* ------------------------------------------------------------------------ */
void svlo_hex8_drm_force(const double *X, const double C[36], double rho, int lumped,
                         const uint8_t ext[8], const double Uo[24], const double Vo[24],
                         const double Ao[24], double f[24]) {
    double M[576], K[576];
    (void)Vo;                                     /* FREE damping: C_e = 0 */
    svlo_hex8_mass(X, rho, lumped, M);
    svlo_hex8_stiffness(X, C, K);
    for (int i = 0; i < 24; i++) {
        double v = 0;
        for (int j = 0; j < 24; j++) {
            if (ext[i / 3] == ext[j / 3]) continue;
            v += M[i * 24 + j] * Ao[j];
        }
        double w = 0;
        for (int j = 0; j < 24; j++) {
            if (ext[i / 3] == ext[j / 3]) continue;
            w += K[i * 24 + j] * Uo[j];
        }
        f[i] = v + w;
    }
}
void svlo_quad4_drm_force(const double *X, double th, const double C[9], double rho, int lumped,
                          const uint8_t ext[4], const double Uo[8], const double Vo[8],
                          const double Ao[8], double f[8]) {
    double M[64], K[64];
    (void)Vo;
    svlo_quad4_mass(X, th, rho, lumped, M);
    svlo_quad4_stiffness(X, th, C, K);
    for (int i = 0; i < 8; i++) {
        double v = 0, w = 0;
        for (int j = 0; j < 8; j++) {
            if (ext[i / 2] == ext[j / 2]) continue;
            v += M[i * 8 + j] * Ao[j];
        }
        for (int j = 0; j < 8; j++) {
            if (ext[i / 2] == ext[j / 2]) continue;
            w += K[i * 8 + j] * Uo[j];
        }
        f[i] = v + w;
    }
}

/* ------------------------------------------------------------------------ */
/* analysis level                                                            */
/* ------------------------------------------------------------------------ */
typedef struct { int i, j; double v; } trip;
static int trip_cmp(const void *a, const void *b) {
    const trip *x = (const trip *)a, *y = (const trip *)b;
    if (x->i != y->i) return x->i < y->i ? -1 : 1;
    if (x->j != y->j) return x->j < y->j ? -1 : 1;
    return 0;
}
typedef struct { int n, cap; trip *t; } tlist;
static void tl_push(tlist *l, int i, int j, double v) {
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 1024; l->t = (trip *)realloc(l->t, l->cap * sizeof(trip)); }
    l->t[l->n].i = i; l->t[l->n].j = j; l->t[l->n].v = v; l->n++;
}
/* COO -> CSR with duplicate summation (Eigen setFromTriplets semantics)        */
typedef struct { int n; int *ptr, *col; double *val; } csr;
static csr tl_to_csr(tlist *l, int n) {
    csr A; A.n = n;
    qsort(l->t, l->n, sizeof(trip), trip_cmp);
    int m = 0;
    for (int k = 0; k < l->n; k++) {
        if (m > 0 && l->t[m - 1].i == l->t[k].i && l->t[m - 1].j == l->t[k].j) l->t[m - 1].v += l->t[k].v;
        else l->t[m++] = l->t[k];
    }
    A.ptr = (int *)calloc(n + 1, sizeof(int)); A.col = (int *)malloc((m + 1) * sizeof(int));
    A.val = (double *)malloc((m + 1) * sizeof(double));
    for (int k = 0; k < m; k++) A.ptr[l->t[k].i + 1]++;
    for (int i = 0; i < n; i++) A.ptr[i + 1] += A.ptr[i];
    for (int k = 0; k < m; k++) { A.col[k] = l->t[k].j; A.val[k] = l->t[k].v; }
    return A;
}
static void csr_free(csr *A) { free(A->ptr); free(A->col); free(A->val); }

static int elem_nn(int kind) { return (kind == SVLO_LIN3DHEXA8 || kind == SVLO_PML3DHEXA8) ? 8 : (kind == SVLO_ZEROLENGTH1D) ? 2 : 4; }
/* ZeroLength1D::ComputeLocalAxes (04-Elements/01-Zero/ZeroLength1D.cpp:318-352): -1 on node i, +1 on node j, direction dir */
static void zl_axes(int ndim, int dir, double a[6]) {
    for (int i = 0; i < 2 * ndim; i++) a[i] = 0.0;
    a[dir] = -1.0; a[ndim + dir] = 1.0;
}
/* PlasticPlaneStrainJ2 (02-Materials/02-NonLinear/PlasticPlaneStrainJ2.cpp:227-278): the 3-D return map on the embedded
 * tensor [0, e11, e22, 0, e12/2, 0] (:235), stress read back from slots 1, 2, 4 (:118-122)                            */
static void j2ps_update(const double par[6], const double eps[3], double st[13], double sig[3]) {
    double e6[6] = {0.0, eps[0], eps[1], 0.0, eps[2], 0.0}, s6[6];
    svlo_j2_update(par, e6, st, s6);
    sig[0] = s6[1]; sig[1] = s6[2]; sig[2] = s6[4];
}
/* initial tangent K*D + 2G(I - D/3) restricted to slots 1, 2, 4 (PlasticPlaneStrainJ2.cpp:150-166) */
static void j2ps_C0(const double par[6], double C[9]) {
    const double K = par[0], G = par[1];
    const double a = K + 2.0 * G * (1.0 - 1.0 / 3.0), b = K + 2.0 * G * (0.0 - 1.0 / 3.0), c = 2.0 * G * 0.5;
    double t[9] = {a, b, 0, b, a, 0, 0, 0, c};
    memcpy(C, t, sizeof t);
}

static int elem_is_pml(int kind) { return kind == SVLO_PML3DHEXA8 || kind == SVLO_PML2DQUAD4; }

typedef struct {
    int nd;                /* element dofs                                         */
    int dofs[72];          /* total dofs (Element::GetTotalDegreeOfFreedom)        */
    double X[24];
    double *Kpml;          /* PML: nd*nd stiffness                                 */
    double sig[8][6];      /* stored Gauss-point stress (solid)                    */
    double st[8][13];      /* J2 state                                             */
    double Ct[8][36];      /* Gauss-point tangents of the plastic materials (6x6 or 3x3 in the first 9), Newton only */
    int has_Ct;
} elem_rt;

static void elem_setup(const svlo_model *m, int e, elem_rt *rt) {
    int kind = m->elem_kind[e], nn = elem_nn(kind);
    rt->nd = 0; rt->Kpml = NULL; rt->has_Ct = 0;
    memset(rt->sig, 0, sizeof rt->sig); memset(rt->st, 0, sizeof rt->st);
    for (int i = 0; i < nn; i++) {
        int nd = m->elem_conn[8 * e + i];
        for (int c = 0; c < m->ndim; c++) rt->X[m->ndim * i + c] = m->coords[m->ndim * nd + c];
        /* lin3DHexa8.cpp:129-147 / PML3DHexa8.cpp GetTotalDegreeOfFreedom        */
        for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++) rt->dofs[rt->nd++] = m->totaldof[p];
    }
}

/* element mass/damping matrices -> triplets with |m_ij| > mtol filter
 * (Assembler.cpp:660-697, 116-158)                                             */
static void elem_MC(const svlo_model *m, int e, const elem_rt *rt, double *Me, double *Ce) {
    int kind = m->elem_kind[e];
    const double *mp = m->mat_par + 8 * m->elem_mat[e];
    const double *at = m->elem_attr + 10 * e;
    int nd = rt->nd;
    memset(Ce, 0, nd * nd * sizeof(double));
    if (kind == SVLO_LIN3DHEXA8) {
        svlo_hex8_mass(rt->X, mp[2], m->lumped, Me);
    } else if (kind == SVLO_LIN2DQUAD4) {
        svlo_quad4_mass(rt->X, at[0], mp[2], m->lumped, Me);
    } else if (kind == SVLO_PML3DHEXA8) {
        svlo_pml3d_matrices(rt->X, mp[0], mp[1], mp[2], at, Me, Ce, NULL, NULL);
    } else if (kind == SVLO_ZEROLENGTH1D) {
        /* M = 0 (ZeroLength1D.cpp:176-190); C = eta a a^T (:212-231, Viscous1DLinear.cpp GetDamping) */
        double a[6];
        zl_axes(m->ndim, (int)at[0], a);
        memset(Me, 0, nd * nd * sizeof(double));
        for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) Ce[i * nd + j] = mp[0] * a[i] * a[j];
        return;
    } else {
        svlo_pml2d_matrices(rt->X, mp[0], mp[1], mp[2], at, Me, Ce, NULL);
    }
    if (!elem_is_pml(kind)) {
        /* Rayleigh: C_e = am*M_e + ak*K0_e  (lin3DHexa8.cpp:354-366)            */
        double am = m->elem_am ? m->elem_am[e] : 0.0, ak = m->elem_ak ? m->elem_ak[e] : 0.0;
        if (am != 0.0 || ak != 0.0) {
            double K0[576], Cm[36];
            int mk = m->mat_kind[m->elem_mat[e]];
            if (kind == SVLO_LIN3DHEXA8) {
                if (mk == SVLO_ELASTIC3DLINEAR) svlo_elastic3d_C(mp[0], mp[1], Cm);
                else { /* Plastic3DJ2 initial tangent K*D + 2G(I - D/3): Plastic3DJ2.cpp:148-160 */
                    memset(Cm, 0, sizeof Cm);
                    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
                        Cm[6 * i + j] = mp[0] + 2.0 * mp[1] * ((i == j ? 1.0 : 0.0) - 1.0 / 3.0);
                    Cm[21] = Cm[28] = Cm[35] = 2.0 * mp[1] * 0.5;
                }
                svlo_hex8_stiffness(rt->X, Cm, K0);
            } else {
                if (mk == SVLO_PLASTICPLANESTRAINJ2) j2ps_C0(mp, Cm);
                else svlo_planestrain_C(mp[0], mp[1], Cm);
                svlo_quad4_stiffness(rt->X, at[0], Cm, K0);
            }
            for (int i = 0; i < nd * nd; i++) Ce[i] += am * Me[i] + ak * K0[i];
        }
    }
}

/* dense LDL^T without pivoting (EigenSolver.hpp:88 SimplicialLDLT semantics)   */
static int ldlt_factor(double *A, int n) {
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k] * A[k * n + k];
        if (d == 0.0 || d != d) return 1;
        A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double v = A[i * n + j];
            for (int k = 0; k < j; k++) v -= A[i * n + k] * A[j * n + k] * A[k * n + k];
            A[i * n + j] = v / d;
        }
    }
    return 0;
}
static void ldlt_solve(const double *A, int n, double *b) {
    for (int i = 0; i < n; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= A[i * n + k] * b[k]; b[i] = v; }
    for (int i = 0; i < n; i++) b[i] /= A[i * n + i];
    for (int i = n - 1; i >= 0; i--) { double v = b[i]; for (int k = i + 1; k < n; k++) v -= A[k * n + i] * b[k]; b[i] = v; }
}

/* The same factorisation and substitutions restricted to the skyline of A (lo[i] = first non-zero column of row i, hi[i] =
 * last row that reaches column i): without pivoting the factor keeps the profile, and the skipped terms are exact zeros, so
 * the results equal ldlt_factor / ldlt_solve bit for bit.  Used by the Newton iteration, which refactors every pass. */
static void skyline_profile(const double *A, int n, int *lo, int *hi) {
    for (int i = 0; i < n; i++) {
        int k = 0;
        while (k < i && A[i * n + k] == 0.0 && A[k * n + i] == 0.0) k++;
        lo[i] = k;
        hi[i] = i;
    }
    for (int i = 0; i < n; i++)
        for (int k = lo[i]; k < i; k++) if (hi[k] < i) hi[k] = i;
}
static int ldlt_factor_sky(double *A, int n, const int *lo) {
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = lo[j]; k < j; k++) d -= A[j * n + k] * A[j * n + k] * A[k * n + k];
        if (d == 0.0 || d != d) return 1;
        A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            if (lo[i] > j) continue;
            double v = A[i * n + j];
            for (int k = (lo[i] > lo[j] ? lo[i] : lo[j]); k < j; k++) v -= A[i * n + k] * A[j * n + k] * A[k * n + k];
            A[i * n + j] = v / d;
        }
    }
    return 0;
}
static void ldlt_solve_sky(const double *A, int n, const int *lo, const int *hi, double *b) {
    for (int i = 0; i < n; i++) { double v = b[i]; for (int k = lo[i]; k < i; k++) v -= A[i * n + k] * b[k]; b[i] = v; }
    for (int i = 0; i < n; i++) b[i] /= A[i * n + i];
    for (int i = n - 1; i >= 0; i--) {
        double v = b[i];
        for (int k = i + 1; k <= hi[i]; k++) if (lo[k] <= i) v -= A[k * n + i] * b[k];
        b[i] = v;
    }
}

/* Mesh::GetTotalToFreeMatrix (06-Mesh/Mesh.cpp:328-381) as a CSR over total dofs */
static csr build_T(const svlo_model *m) {
    tlist l = {0, 0, NULL};
    for (int nd = 0; nd < m->n_nodes; nd++)
        for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++) {
            int fr = m->freedof[p], to = m->totaldof[p];
            if (fr > -1) tl_push(&l, to, fr, 1.0);
            if (fr < -1)
                for (int c = 0; c < m->n_cons; c++)
                    if (m->cons_tag[c] == fr)
                        for (int q = m->cons_ptr[c]; q < m->cons_ptr[c + 1]; q++)
                            tl_push(&l, m->cons_slave[c], m->cons_master[q], m->cons_factor[q]);
        }
    csr T = tl_to_csr(&l, m->n_total);
    free(l.t);
    return T;
}

static void material_update(const svlo_model *m, int e, elem_rt *rt, const double *Utot) {
    int kind = m->elem_kind[e], mk = m->mat_kind[m->elem_mat[e]];
    const double *mp = m->mat_par + 8 * m->elem_mat[e];
    double ue[24];
    for (int i = 0; i < rt->nd; i++) ue[i] = Utot[rt->dofs[i]];
    if (kind == SVLO_LIN3DHEXA8) {
        double eps[8][6];
        svlo_hex8_strain(rt->X, ue, eps);
        if (mk == SVLO_ELASTIC3DLINEAR) {
            double C[36];
            svlo_elastic3d_C(mp[0], mp[1], C);
            for (int g = 0; g < 8; g++)
                for (int a = 0; a < 6; a++) {
                    double v = 0;
                    for (int b = 0; b < 6; b++) v += C[6 * a + b] * eps[g][b];
                    rt->sig[g][a] = v;
                }
        } else {
            for (int g = 0; g < 8; g++) svlo_j2_update_tangent(mp, eps[g], rt->st[g], rt->sig[g], rt->Ct[g]);
            rt->has_Ct = 1;
        }
    } else if (kind == SVLO_LIN2DQUAD4) {
        double eps[4][3], C[9];
        svlo_quad4_strain(rt->X, ue, eps);
        if (mk == SVLO_PLASTICPLANESTRAINJ2) {
            for (int g = 0; g < 4; g++) {
                /* PlasticPlaneStrainJ2.cpp:235 embedding; tangent = rows / columns 1, 2, 4 (:151-159) */
                double e6[6] = {0.0, eps[g][0], eps[g][1], 0.0, eps[g][2], 0.0}, s6[6], C6[36];
                svlo_j2_update_tangent(mp, e6, rt->st[g], s6, C6);
                rt->sig[g][0] = s6[1]; rt->sig[g][1] = s6[2]; rt->sig[g][2] = s6[4];
                const int ix[3] = {1, 2, 4};
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) rt->Ct[g][3 * a + b] = C6[6 * ix[a] + ix[b]];
            }
            rt->has_Ct = 1;
            return;
        }
        svlo_planestrain_C(mp[0], mp[1], C);
        for (int g = 0; g < 4; g++)
            for (int a = 0; a < 3; a++) {
                double v = 0;
                for (int b = 0; b < 3; b++) v += C[3 * a + b] * eps[g][b];
                rt->sig[g][a] = v;
            }
    }
}

static void elem_fint(const svlo_model *m, int e, const elem_rt *rt, const double *U, double *fe) {
    int kind = m->elem_kind[e];
    if (kind == SVLO_LIN3DHEXA8) svlo_hex8_force(rt->X, rt->sig, fe);
    else if (kind == SVLO_LIN2DQUAD4) {
        double s3[4][3];
        for (int g = 0; g < 4; g++) for (int a = 0; a < 3; a++) s3[g][a] = rt->sig[g][a];
        svlo_quad4_force(rt->X, m->elem_attr[10 * e], s3, fe);
    } else if (kind == SVLO_ZEROLENGTH1D) {
        /* stress(0) * localAxes with Viscous1DLinear::GetStress() == 0 (ZeroLength1D.cpp:241-256): the dashpot acts
         * through the damping matrix only                                       */
        for (int i = 0; i < rt->nd; i++) fe[i] = 0.0;
    } else {
        /* PML: f = K_e u_e, PML3DHexa8.cpp:716-743, PML2DQuad4.cpp:474-494       */
        int nd = rt->nd;
        for (int i = 0; i < nd; i++) {
            double v = 0;
            for (int j = 0; j < nd; j++) v += rt->Kpml[i * nd + j] * U[rt->dofs[j]];
            fe[i] = v;
        }
    }
}

int svlo_mass_diagonal(const svlo_model *m, double *Md) {
    memset(Md, 0, m->n_total * sizeof(double));
    int off = 0;
    for (int q = 0; q < m->n_mass; q++) {
        int nd = m->mass_node[q];
        for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++, off++)
            if (fabs(m->mass_val[off]) > m->mtol) Md[m->totaldof[p]] += m->mass_val[off];
    }
    double *Me = (double *)malloc(72 * 72 * sizeof(double)), *Ce = (double *)malloc(72 * 72 * sizeof(double));
    for (int e = 0; e < m->n_elem; e++) {
        elem_rt rt; elem_setup(m, e, &rt);
        elem_MC(m, e, &rt, Me, Ce);
        for (int i = 0; i < rt.nd; i++)
            if (fabs(Me[i * rt.nd + i]) > m->mtol) Md[rt.dofs[i]] += Me[i * rt.nd + i];
    }
    free(Me); free(Ce);
    return 0;
}

int svlo_internal_force(const svlo_model *m, const double *U, double *F) {
    memset(F, 0, m->n_total * sizeof(double));
    for (int e = 0; e < m->n_elem; e++) {
        elem_rt rt; elem_setup(m, e, &rt);
        double fe[72];
        double *Kp = NULL;
        if (elem_is_pml(m->elem_kind[e])) {
            const double *mp = m->mat_par + 8 * m->elem_mat[e];
            Kp = (double *)malloc(rt.nd * rt.nd * sizeof(double));
            if (m->elem_kind[e] == SVLO_PML3DHEXA8) svlo_pml3d_matrices(rt.X, mp[0], mp[1], mp[2], m->elem_attr + 10 * e, NULL, NULL, Kp, NULL);
            else svlo_pml2d_matrices(rt.X, mp[0], mp[1], mp[2], m->elem_attr + 10 * e, NULL, NULL, Kp);
            rt.Kpml = Kp;
        } else material_update(m, e, &rt, U);
        elem_fint(m, e, &rt, U, fe);
        for (int i = 0; i < rt.nd; i++)
            if (fabs(fe[i]) > m->ftol) F[rt.dofs[i]] += fe[i];
        free(Kp);
    }
    return 0;
}

/* element tangent stiffness for the implicit integrators (Assembler::ComputeStiffnessMatrix, Assembler.cpp:70-113):
 * linear materials only                                                         */
static int elem_K(const svlo_model *m, int e, const elem_rt *rt, double *Ke) {
    int kind = m->elem_kind[e], mk = m->mat_kind[m->elem_mat[e]];
    const double *mp = m->mat_par + 8 * m->elem_mat[e];
    double Cm[36];
    memset(Ke, 0, (size_t)rt->nd * rt->nd * sizeof(double));
    if (kind == SVLO_LIN3DHEXA8 && mk == SVLO_ELASTIC3DLINEAR) {
        svlo_elastic3d_C(mp[0], mp[1], Cm); svlo_hex8_stiffness(rt->X, Cm, Ke); return 0;
    }
    if (kind == SVLO_LIN2DQUAD4 && mk == SVLO_ELASTIC2DPLANESTRAIN) {
        svlo_planestrain_C(mp[0], mp[1], Cm); svlo_quad4_stiffness(rt->X, m->elem_attr[10 * e], Cm, Ke); return 0;
    }
    if (kind == SVLO_LIN3DHEXA8 && mk == SVLO_PLASTIC3DJ2) {
        if (rt->has_Ct) { hex8_stiffness_gp(rt->X, &rt->Ct[0][0], 36, Ke); return 0; }
        memset(Cm, 0, sizeof Cm);                 /* virgin material: K D + 2G (I - D/3), Plastic3DJ2.cpp:28-31 */
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Cm[6 * i + j] = mp[0] + 2.0 * mp[1] * ((i == j ? 1.0 : 0.0) - 1.0 / 3.0);
        Cm[21] = Cm[28] = Cm[35] = 2.0 * mp[1] * 0.5;
        svlo_hex8_stiffness(rt->X, Cm, Ke); return 0;
    }
    if (kind == SVLO_LIN2DQUAD4 && mk == SVLO_PLASTICPLANESTRAINJ2) {
        if (rt->has_Ct) { quad4_stiffness_gp(rt->X, m->elem_attr[10 * e], &rt->Ct[0][0], 36, Ke); return 0; }
        j2ps_C0(mp, Cm); svlo_quad4_stiffness(rt->X, m->elem_attr[10 * e], Cm, Ke); return 0;
    }
    if (kind == SVLO_ZEROLENGTH1D) return 0;      /* Viscous1DLinear::GetTangentStiffness() == 0 */
    if (elem_is_pml(kind) && rt->Kpml) { memcpy(Ke, rt->Kpml, (size_t)rt->nd * rt->nd * sizeof(double)); return 0; }
    return 1;
}

/* integrator: 0 CentralDifference (10-Integrators/02-CentralDifference/CentralDifference.cpp),
 *             1 NewmarkBeta, average acceleration (10-Integrators/03-Newmark/NewmarkBeta.cpp:21-36 Initialize,
 *               :64-79 ComputeNewStep, :106-121 ComputeEffectiveForce, :124-133 ComputeEffectiveStiffness),
 *             2 ExtendedNewmarkBeta (10-Integrators/03-Newmark/ExtendedNewmarkBeta.cpp): NewmarkBeta plus the PML history
 *               term, Keff += dt/3 G, rhs -= G (Ubar + dt U + dt^2/6 V), Ubar += dt U + dt/3 dU + dt^2/6 V with
 *               G = Assembler::ComputePMLHistoryMatrix (Assembler.cpp:162-205; PML3DHexa8::ComputePMLMatrix,
 *               PML2DQuad4 returns an empty matrix)                                                                 */
/* algorithm: NULL = Linear (09-Algorithms/01-Linear/Linear.cpp:22-56); otherwise NewtonRaphson (09-Algorithms/02-Newton/
 * NewtonRaphson.cpp:22-56) with tolerance, iteration cap and convergence test `flag` of Algorithm::ComputeConvergence
 * (Algorithm.cpp:122-186): 1 |Feff|, 2 |du|, 3 |du.*Feff|, 4 |Feff| / |Feff_0|, 5 |du| / |dU_0|, 6 |du.*Feff| / |du.*Feff|_0,
 * 7 |du| / |dU|, 8 none (one iteration).  NewmarkBeta only.
 * Like the reference, every Newton iteration calls Material::UpdateState on the LIVE state.          */
typedef struct { double tol; int nmax, flag; } svlo_newton;
static int run_dynamic_alg(const svlo_model *m, int integrator, const svlo_newton *alg, int nt, int field, int n_rec,
                           const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads);
static int run_dynamic(const svlo_model *m, int integrator, int nt, int field, int n_rec,
                       const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads) {
    return run_dynamic_alg(m, integrator, NULL, nt, field, n_rec, rec_dofs, out, Ufinal, nthreads);
}
static int run_dynamic_alg(const svlo_model *m, int integrator, const svlo_newton *alg, int nt, int field, int n_rec,
                           const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads) {
    const int nT = m->n_total, nF = m->n_free, nE = m->n_elem;
    const double dt = m->dt;
    int rc = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    elem_rt *rt = (elem_rt *)malloc((size_t)nE * sizeof(elem_rt));
    double *U = (double *)calloc(nT, sizeof(double)), *V = (double *)calloc(nT, sizeof(double)),
           *A = (double *)calloc(nT, sizeof(double)), *Up = (double *)calloc(nT, sizeof(double));
    double *Fint = (double *)calloc(nT, sizeof(double)), *Fext = (double *)calloc(nT, sizeof(double)),
           *Ftmp = (double *)calloc(nT, sizeof(double)), *rhs = (double *)calloc(nT, sizeof(double)),
           *dUt = (double *)calloc(nT, sizeof(double)), *Utr = (double *)calloc(nT, sizeof(double));
    double *Feff = (double *)calloc(nF + 1, sizeof(double)), *dU = (double *)calloc(nF + 1, sizeof(double));
    double *fe_all = (double *)malloc((size_t)nE * 72 * sizeof(double));
    double *SM = (double *)calloc(nT, sizeof(double)), *Reac = (double *)calloc(nT, sizeof(double));
    /* Node::IsFixed: any restrained dof (Driver.hpp:338-341); rows of R are kept for all dofs of such nodes */
    unsigned char *node_fixed = (unsigned char *)calloc(m->n_nodes, 1), *is_fixed_dof = (unsigned char *)calloc(nT, 1);
    for (int nd = 0; nd < m->n_nodes; nd++) {
        for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++) if (m->freedof[p] == -1) node_fixed[nd] = 1;
        if (node_fixed[nd]) for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++) is_fixed_dof[m->totaldof[p]] = 1;
    }
    if (m->U0) memcpy(U, m->U0, nT * sizeof(double));
    if (m->V0) memcpy(V, m->V0, nT * sizeof(double));
    if (m->A0) memcpy(A, m->A0, nT * sizeof(double));
    /* CentralDifference::Initialize :61 */
    for (int i = 0; i < nT; i++) Up[i] = U[i] - dt * V[i] + dt * dt / 2.0 * A[i];

    /* M, C in total space (Assembler.cpp:47-67, 116-158) */
    tlist lM = {0, 0, NULL}, lC = {0, 0, NULL};
    {
        int off = 0;
        for (int q = 0; q < m->n_mass; q++) {
            int nd = m->mass_node[q];
            for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++, off++)
                if (fabs(m->mass_val[off]) > m->mtol) tl_push(&lM, m->totaldof[p], m->totaldof[p], m->mass_val[off]);
        }
        double *Me = (double *)malloc(72 * 72 * sizeof(double)), *Ce = (double *)malloc(72 * 72 * sizeof(double));
        for (int e = 0; e < nE; e++) {
            elem_setup(m, e, &rt[e]);
            elem_MC(m, e, &rt[e], Me, Ce);
            int nd = rt[e].nd;
            for (int j = 0; j < nd; j++)
                for (int i = 0; i < nd; i++) {
                    if (fabs(Me[i * nd + j]) > m->mtol) tl_push(&lM, rt[e].dofs[i], rt[e].dofs[j], Me[i * nd + j]);
                    if (fabs(Ce[i * nd + j]) > m->mtol) tl_push(&lC, rt[e].dofs[i], rt[e].dofs[j], Ce[i * nd + j]);
                }
            if (elem_is_pml(m->elem_kind[e])) {
                const double *mp = m->mat_par + 8 * m->elem_mat[e];
                rt[e].Kpml = (double *)malloc(nd * nd * sizeof(double));
                if (m->elem_kind[e] == SVLO_PML3DHEXA8) svlo_pml3d_matrices(rt[e].X, mp[0], mp[1], mp[2], m->elem_attr + 10 * e, NULL, NULL, rt[e].Kpml, NULL);
                else svlo_pml2d_matrices(rt[e].X, mp[0], mp[1], mp[2], m->elem_attr + 10 * e, NULL, NULL, rt[e].Kpml);
            }
        }
        free(Me); free(Ce);
    }
    /* CentralDifference: Keff = M/dt^2 + C/2dt (:70) and Kminus = M/dt^2 - C/2dt (:217).
     * NewmarkBeta: Keff = K + 4/dt^2 M + 2/dt C (NewmarkBeta.cpp:128-130); Kminus carries M, lC stays C           */
    tlist lKp = {0, 0, NULL}, lKm = {0, 0, NULL};
    tlist lG = {0, 0, NULL};
    if (integrator >= 1) {
        double *Ke = (double *)malloc(72 * 72 * sizeof(double));
        for (int e = 0; e < nE && !rc; e++) {
            if (elem_K(m, e, &rt[e], Ke)) { rc = 4; break; }
            int nd = rt[e].nd;
            for (int j = 0; j < nd; j++)
                for (int i = 0; i < nd; i++)
                    if (fabs(Ke[i * nd + j]) > 1e-12 /* ktol, Driver.hpp:1805 default */) tl_push(&lKp, rt[e].dofs[i], rt[e].dofs[j], Ke[i * nd + j]);
        }
        if (integrator == 2) {
            for (int e = 0; e < nE; e++) {
                if (m->elem_kind[e] != SVLO_PML3DHEXA8) continue;
                const double *mp = m->mat_par + 8 * m->elem_mat[e];
                int nd = rt[e].nd;
                svlo_pml3d_matrices(rt[e].X, mp[0], mp[1], mp[2], m->elem_attr + 10 * e, NULL, NULL, NULL, Ke);
                for (int j = 0; j < nd; j++)
                    for (int i = 0; i < nd; i++)
                        if (fabs(Ke[i * nd + j]) > 1e-12) {
                            tl_push(&lG, rt[e].dofs[i], rt[e].dofs[j], Ke[i * nd + j]);
                            tl_push(&lKp, rt[e].dofs[i], rt[e].dofs[j], dt / 3.0 * Ke[i * nd + j]);
                        }
            }
        }
        free(Ke);
        for (int k = 0; k < lM.n; k++) {
            tl_push(&lKp, lM.t[k].i, lM.t[k].j, 4.0 / dt / dt * lM.t[k].v);
            tl_push(&lKm, lM.t[k].i, lM.t[k].j, lM.t[k].v);
        }
        for (int k = 0; k < lC.n; k++) tl_push(&lKp, lC.t[k].i, lC.t[k].j, 2.0 / dt * lC.t[k].v);
    } else {
    for (int k = 0; k < lM.n; k++) {
        tl_push(&lKp, lM.t[k].i, lM.t[k].j, 1.0 / dt / dt * lM.t[k].v);
        tl_push(&lKm, lM.t[k].i, lM.t[k].j, 1.0 / dt / dt * lM.t[k].v);
    }
    for (int k = 0; k < lC.n; k++) {
        tl_push(&lKp, lC.t[k].i, lC.t[k].j, 1.0 / 2.0 / dt * lC.t[k].v);
        tl_push(&lKm, lC.t[k].i, lC.t[k].j, -(1.0 / 2.0 / dt * lC.t[k].v));
    }
    }
    csr Cs = tl_to_csr(&lC, nT), Gs = tl_to_csr(&lG, nT);
    double *Ubar = (double *)calloc(nT, sizeof(double)), *Gtmp = (double *)calloc(nT, sizeof(double));
    csr Kp = tl_to_csr(&lKp, nT), Km = tl_to_csr(&lKm, nT);
    csr T = build_T(m);
    /* Keff_free = T' Keff T (:224-231) */
    tlist lF = {0, 0, NULL};
    for (int i = 0; i < nT; i++)
        for (int p = Kp.ptr[i]; p < Kp.ptr[i + 1]; p++) {
            int j = Kp.col[p];
            for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++)
                for (int b = T.ptr[j]; b < T.ptr[j + 1]; b++)
                    tl_push(&lF, T.col[a], T.col[b], T.val[a] * T.val[b] * Kp.val[p]);
        }
    csr Kf = tl_to_csr(&lF, nF);
    /* split free dofs into diagonal rows and a coupled block */
    int *cidx = (int *)malloc((nF + 1) * sizeof(int));
    int nc = 0;
    double *Kdiag = (double *)calloc(nF + 1, sizeof(double));
    for (int i = 0; i < nF; i++) {
        int coupled = 0;
        for (int p = Kf.ptr[i]; p < Kf.ptr[i + 1]; p++) {
            if (Kf.col[p] == i) Kdiag[i] = Kf.val[p];
            else if (Kf.val[p] != 0.0) coupled = 1;
        }
        cidx[i] = coupled ? nc++ : -1;
    }
    double *Kc = NULL, *bc = NULL;
    int *sky_lo = NULL, *sky_hi = NULL;
    if (nc > 0) {
        Kc = (double *)calloc((size_t)nc * nc, sizeof(double));
        bc = (double *)calloc(nc, sizeof(double));
        for (int i = 0; i < nF; i++) {
            if (cidx[i] < 0) continue;
            for (int p = Kf.ptr[i]; p < Kf.ptr[i + 1]; p++)
                if (cidx[Kf.col[p]] >= 0) Kc[(size_t)cidx[i] * nc + cidx[Kf.col[p]]] = Kf.val[p];
        }
        /* envelope form of the same factorisation (bit-identical to the dense one, see skyline_profile): the coupled block of a
         * PML model is banded in the pre-processor's numbering, which keeps mid-size cases (10^4 unknowns) within minutes */
        sky_lo = (int *)malloc((nc + 1) * sizeof(int)); sky_hi = (int *)malloc((nc + 1) * sizeof(int));
        skyline_profile(Kc, nc, sky_lo, sky_hi);
        if (ldlt_factor_sky(Kc, nc, sky_lo)) { rc = 2; goto done; }
    }
    if (rc) goto done;

    for (int k = 1; k < nt; k++) {            /* DynamicAnalysis.cpp:36 */
        /* --- Fint: Assembler.cpp:239-269 (ascending element order, ftol filter) */
#pragma omp parallel for schedule(static)
        for (int e = 0; e < nE; e++) elem_fint(m, e, &rt[e], U, fe_all + (size_t)72 * e);
        memset(Fint, 0, nT * sizeof(double));
        for (int e = 0; e < nE; e++) {
            const double *fe = fe_all + (size_t)72 * e;
            for (int i = 0; i < rt[e].nd; i++)
                if (fabs(fe[i]) > m->ftol) Fint[rt[e].dofs[i]] += fe[i];
        }
        /* --- Fext: Assembler.cpp:290-489 */
        memset(Fext, 0, nT * sizeof(double));
        for (int l = 0; l < m->n_pload; l++) {
            memset(Ftmp, 0, nT * sizeof(double));
            int ntl = m->pl_nt[l];
            double amp = (ntl == 1) ? m->pl_series[m->pl_sptr[l]] : m->pl_series[m->pl_sptr[l] + k];
            for (int q = m->pl_ptr[l]; q < m->pl_ptr[l + 1]; q++) {
                int nd = m->pl_nodes[q], c = 0;
                for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1] && c < 3; p++, c++)
                    if (c < m->ndim) Ftmp[m->totaldof[p]] = amp * m->pl_dir[3 * l + c];   /* assign: :330,347 */
            }
            for (int i = 0; i < nT; i++) Fext[i] += m->pl_factor[l] * Ftmp[i];
        }
        if (m->n_drm_elem > 0) {
            memset(Ftmp, 0, nT * sizeof(double));
            int nf = 3 * m->ndim;
            for (int q = 0; q < m->n_drm_elem; q++) {
                int e = m->drm_elem[q], nn = elem_nn(m->elem_kind[e]), nd = m->ndim;
                double Uo[24], Vo[24], Ao[24], f[24];
                uint8_t ext[8];
                for (int i = 0; i < nn; i++) {
                    int node = m->elem_conn[8 * e + i], li = -1;
                    for (int z = 0; z < m->n_drm_node; z++) if (m->drm_node[z] == node) { li = z; break; }
                    ext[i] = (li >= 0) ? m->drm_ext[li] : 0;
                    for (int c = 0; c < nd; c++) {
                        double sgn = (li >= 0 && m->drm_ext[li]) ? -1.00 : 1.0;   /* Driver.hpp:1714-1716 */
                        const double *row = (li >= 0) ? m->drm_field + ((size_t)li * m->drm_nt + k) * nf : NULL;
                        Uo[nd * i + c] = row ? sgn * row[c] : 0.0;
                        Vo[nd * i + c] = row ? sgn * row[nd + c] : 0.0;
                        Ao[nd * i + c] = row ? sgn * row[2 * nd + c] : 0.0;
                    }
                }
                const double *mp = m->mat_par + 8 * m->elem_mat[e];
                if (m->elem_kind[e] == SVLO_LIN3DHEXA8) {
                    double C[36]; svlo_elastic3d_C(mp[0], mp[1], C);
                    svlo_hex8_drm_force(rt[e].X, C, mp[2], m->lumped, ext, Uo, Vo, Ao, f);
                } else {
                    double C[9]; svlo_planestrain_C(mp[0], mp[1], C);
                    svlo_quad4_drm_force(rt[e].X, m->elem_attr[10 * e], C, mp[2], m->lumped, ext, Uo, Vo, Ao, f);
                }
                for (int i = 0; i < rt[e].nd; i++) Ftmp[rt[e].dofs[i]] += f[i];
            }
            for (int i = 0; i < nT; i++) Fext[i] += m->drm_factor * Ftmp[i];
        }
        if (integrator >= 1) {
            /* Fext + Fbar - Fint + M (4/dt V + A - 4/dt^2 dU) + C (V - 2/dt dU) with dU = 0 (Linear.cpp:25):
             * NewmarkBeta.cpp:116-118                                               */
            for (int i = 0; i < nT; i++) Ftmp[i] = 4.0 / dt * V[i] + A[i];
            for (int i = 0; i < nT; i++) {
                double v = 0, w = 0;
                for (int p = Km.ptr[i]; p < Km.ptr[i + 1]; p++) v += Km.val[p] * Ftmp[Km.col[p]];
                for (int p = Cs.ptr[i]; p < Cs.ptr[i + 1]; p++) w += Cs.val[p] * V[Cs.col[p]];
                rhs[i] = Fext[i] - Fint[i] + v + w;
            }
            if (integrator == 2) {                 /* - G (Ubar + dt U + dt^2/6 V): ExtendedNewmarkBeta.cpp ComputeEffectiveForce */
                for (int i = 0; i < nT; i++) Gtmp[i] = Ubar[i] + dt * U[i] + dt * dt / 6.0 * V[i];
                for (int i = 0; i < nT; i++) {
                    double g = 0;
                    for (int p = Gs.ptr[i]; p < Gs.ptr[i + 1]; p++) g += Gs.val[p] * Gtmp[Gs.col[p]];
                    rhs[i] -= g;
                }
            }
        } else {
        /* --- Feff = T'(Fext - Fint + Kminus (U-Up)) : CentralDifference.cpp:217-220 */
        for (int i = 0; i < nT; i++) Ftmp[i] = U[i] - Up[i];
        for (int i = 0; i < nT; i++) {
            double v = 0;
            for (int p = Km.ptr[i]; p < Km.ptr[i + 1]; p++) v += Km.val[p] * Ftmp[Km.col[p]];
            rhs[i] = Fext[i] - Fint[i] + v;
        }
        }
        memset(Feff, 0, nF * sizeof(double));
        for (int i = 0; i < nT; i++)
            for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++) Feff[T.col[a]] += T.val[a] * rhs[i];
        /* --- support motion (Linear.cpp:40 -> CentralDifference::ComputeSupportMotionVector, CentralDifference.cpp:189-202):
         * SupportMotion = sum factor (g(k) - g(k-1)) with g(k) = Xo[k] if k < size else Xo[0] (Assembler.cpp:493-533,
         * Node.cpp:228-247);  Feff -= T' (Keff SupportMotion) with Keff the integrator's `K` member (:70)            */
        if (m->n_sup > 0) {
            if (integrator != 0) { rc = 5; goto done; }    /* the Newmark variants are not restated */
            memset(SM, 0, nT * sizeof(double));
            for (int q = 0; q < m->n_sup; q++) {
                const double *xo = m->sup_series + m->sup_ptr[q];
                const int sz = m->sup_ptr[q + 1] - m->sup_ptr[q];
                double dg = (k < sz) ? xo[k] : xo[0];
                if (k > 0) dg -= (k - 1 < sz) ? xo[k - 1] : xo[0];
                SM[m->sup_dof[q]] += m->sup_factor[q] * dg;
            }
            for (int i = 0; i < nT; i++) {
                double v = 0;
                for (int p = Kp.ptr[i]; p < Kp.ptr[i + 1]; p++) v += Kp.val[p] * SM[Kp.col[p]];
                for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++) Feff[T.col[a]] -= T.val[a] * v;
            }
        }
        if (alg && integrator == 1) {
            /* ---- NewtonRaphson::ComputeNewIncrement: iterate on the increment dU of this step ------------------------- */
            double *Kn = (double *)malloc((size_t)nF * nF * sizeof(double)), *du = (double *)malloc((nF + 1) * sizeof(double));
            double *Ke = (double *)malloc(72 * 72 * sizeof(double));
            int *sky_lo = (int *)malloc((nF + 1) * sizeof(int)), *sky_hi = (int *)malloc((nF + 1) * sizeof(int));
            double normfactor = 0.0, residual = 0.0;
            int it = 0;
            memset(dU, 0, (nF + 1) * sizeof(double));
            memset(dUt, 0, nT * sizeof(double));
            do {
                if (it > 0) {
                    /* effective force with the accumulated increment (NewmarkBeta.cpp:116-118) and the stresses of the
                     * last UpdateState                                              */
#pragma omp parallel for schedule(static)
                    for (int e = 0; e < nE; e++) elem_fint(m, e, &rt[e], U, fe_all + (size_t)72 * e);
                    memset(Fint, 0, nT * sizeof(double));
                    for (int e = 0; e < nE; e++) {
                        const double *fe = fe_all + (size_t)72 * e;
                        for (int i = 0; i < rt[e].nd; i++)
                            if (fabs(fe[i]) > m->ftol) Fint[rt[e].dofs[i]] += fe[i];
                    }
                    for (int i = 0; i < nT; i++) { Ftmp[i] = 4.0 / dt * V[i] + A[i] - 4.0 / dt / dt * dUt[i]; Utr[i] = V[i] - 2.0 / dt * dUt[i]; }
                    for (int i = 0; i < nT; i++) {
                        double v = 0, w = 0;
                        for (int p = Km.ptr[i]; p < Km.ptr[i + 1]; p++) v += Km.val[p] * Ftmp[Km.col[p]];
                        for (int p = Cs.ptr[i]; p < Cs.ptr[i + 1]; p++) w += Cs.val[p] * Utr[Cs.col[p]];
                        rhs[i] = Fext[i] - Fint[i] + v + w;
                    }
                    memset(Feff, 0, nF * sizeof(double));
                    for (int i = 0; i < nT; i++)
                        for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++) Feff[T.col[a]] += T.val[a] * rhs[i];
                }
                /* Keff = T' (K_t + 4/dt^2 M + 2/dt C) T from the CURRENT tangents (NewmarkBeta.cpp:124-133) */
                memset(Kn, 0, (size_t)nF * nF * sizeof(double));
                for (int p = 0; p < lM.n; p++) {          /* element + nodal masses, already filtered */
                    int i = lM.t[p].i, j = lM.t[p].j;
                    for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++)
                        for (int b = T.ptr[j]; b < T.ptr[j + 1]; b++)
                            Kn[(size_t)T.col[a] * nF + T.col[b]] += T.val[a] * T.val[b] * 4.0 / dt / dt * lM.t[p].v;
                }
                for (int p = 0; p < lC.n; p++) {
                    int i = lC.t[p].i, j = lC.t[p].j;
                    for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++)
                        for (int b = T.ptr[j]; b < T.ptr[j + 1]; b++)
                            Kn[(size_t)T.col[a] * nF + T.col[b]] += T.val[a] * T.val[b] * 2.0 / dt * lC.t[p].v;
                }
                for (int e = 0; e < nE; e++) {
                    if (elem_K(m, e, &rt[e], Ke)) { rc = 4; break; }
                    int nd = rt[e].nd;
                    for (int j = 0; j < nd; j++)
                        for (int i = 0; i < nd; i++) {
                            if (!(fabs(Ke[i * nd + j]) > 1e-12)) continue;
                            int gi = rt[e].dofs[i], gj = rt[e].dofs[j];
                            for (int a = T.ptr[gi]; a < T.ptr[gi + 1]; a++)
                                for (int b = T.ptr[gj]; b < T.ptr[gj + 1]; b++)
                                    Kn[(size_t)T.col[a] * nF + T.col[b]] += T.val[a] * T.val[b] * Ke[i * nd + j];
                        }
                }
                if (rc) break;
                skyline_profile(Kn, nF, sky_lo, sky_hi);
                if (ldlt_factor_sky(Kn, nF, sky_lo)) { rc = 2; break; }
                memcpy(du, Feff, nF * sizeof(double));
                ldlt_solve_sky(Kn, nF, sky_lo, sky_hi, du);
                double n_du = 0, n_dU = 0, n_F = 0, n_E = 0;
                for (int i = 0; i < nF; i++) {
                    dU[i] += du[i]; n_du += du[i] * du[i]; n_dU += dU[i] * dU[i]; n_F += Feff[i] * Feff[i];
                    n_E += (du[i] * Feff[i]) * (du[i] * Feff[i]);      /* norm of the componentwise product, :137-141 */
                }
                n_du = sqrt(n_du); n_dU = sqrt(n_dU); n_F = sqrt(n_F); n_E = sqrt(n_E);
                switch (alg->flag) {                       /* Algorithm.cpp:122-186 */
                case 1: residual = n_F; break;
                case 2: residual = n_du; break;
                case 3: residual = n_E; break;
                case 4: if (it) residual = n_F / normfactor; else { normfactor = n_F; residual = n_F; } break;
                case 5: if (it) residual = n_du / normfactor; else { normfactor = n_dU; residual = n_du; } break;
                case 6: if (it) residual = n_E / normfactor; else { normfactor = n_E; residual = n_E; } break;
                case 7: residual = n_du / n_dU; break;
                case 8: residual = -1.0; break;            /* "maximum number of iterations": one pass, :179-182 */
                default: residual = 0.0; break;            /* undefined test: Residual stays 0.0, :127,183-185 */
                }
                /* UpdateStatesIncrements (Algorithm.cpp:18-56) */
                for (int i = 0; i < nT; i++) {
                    double v = 0;
                    for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++) v += T.val[a] * dU[T.col[a]];
                    dUt[i] = v;
                    Utr[i] = U[i] + v;
                }
#pragma omp parallel for schedule(static)
                for (int e = 0; e < nE; e++) material_update(m, e, &rt[e], Utr);
                it++;
            } while (residual > alg->tol && it < alg->nmax);
            free(Kn); free(du); free(Ke); free(sky_lo); free(sky_hi);
            if (rc) goto done;
            for (int i = 0; i < nT; i++) {                 /* NewmarkBeta.cpp:73-76 */
                U[i] += dUt[i];
                A[i] = 4.0 / dt / dt * dUt[i] - 4.0 / dt * V[i] - A[i];
                V[i] = 2.0 / dt * dUt[i] - V[i];
            }
            for (int i = 0; i < nT; i++) if (U[i] != U[i]) rc = 3;
            const double *srcn = field == 0 ? U : field == 1 ? V : A;
            for (int q = 0; q < n_rec; q++) out[(size_t)(k - 1) * n_rec + q] = srcn[rec_dofs[q]];
            continue;
        }
        /* --- solve (EigenSolver.cpp:19-60) */
        for (int i = 0; i < nF; i++)
            if (cidx[i] < 0) dU[i] = Feff[i] / Kdiag[i]; else bc[cidx[i]] = Feff[i];
        if (nc > 0) {
            ldlt_solve_sky(Kc, nc, sky_lo, sky_hi, bc);
            for (int i = 0; i < nF; i++) if (cidx[i] >= 0) dU[i] = bc[cidx[i]];
        }
        /* --- dU_total = T dU ; UpdateStatesIncrements: Algorithm.cpp:18-56 */
        for (int i = 0; i < nT; i++) {
            double v = 0;
            for (int a = T.ptr[i]; a < T.ptr[i + 1]; a++) v += T.val[a] * dU[T.col[a]];
            dUt[i] = v;
            Utr[i] = U[i] + v;
        }
#pragma omp parallel for schedule(static)
        for (int e = 0; e < nE; e++) material_update(m, e, &rt[e], Utr);
        /* CentralDifference.cpp:135: dU = T dU_free + SupportMotion -- added AFTER UpdateStatesIncrements, which only sees
         * T dU_free (Algorithm.cpp:23): the element stresses lag the support displacement by one step */
        if (m->n_sup > 0) for (int i = 0; i < nT; i++) dUt[i] += SM[i];
        if (integrator >= 1) {
            /* NewmarkBeta.cpp:73-76; ExtendedNewmarkBeta updates the PML history first */
            for (int i = 0; i < nT; i++) {
                if (integrator == 2) Ubar[i] = Ubar[i] + dt * U[i] + dt / 3.0 * dUt[i] + dt * dt / 6.0 * V[i];
                U[i] += dUt[i];
                A[i] = 4.0 / dt / dt * dUt[i] - 4.0 / dt * V[i] - A[i];
                V[i] = 2.0 / dt * dUt[i] - V[i];
            }
        } else
        /* --- CentralDifference.cpp:138-148 */
        for (int i = 0; i < nT; i++) {
            V[i] = 1.0 / 2.0 / dt * (U[i] + dUt[i] - Up[i]);
            A[i] = 1.0 / dt / dt * (dUt[i] - U[i] + Up[i]);
            Up[i] = U[i];
            U[i] += dUt[i];
        }
        for (int i = 0; i < nT; i++) if (U[i] != U[i]) rc = 3;
        if (field == 3) {
            /* DynamicAnalysis::UpdateDomain (DynamicAnalysis.cpp:130-150): R = ComputeDynamicInternalForceVector - Fext(k) - Fbar
             * (CentralDifference.cpp:155-171; NewmarkBeta has the same body) with the node states of this step.
             * Assembler.cpp:272-287, 568-619: node inertia forces Mass .* A of every node that carries a point mass, plus
             * f_int + C_e V_e + M_e A_e of every element that HAS A FIXED NODE (Element.cpp:53-64, lin3DHexa8.cpp:415-436);
             * no tolerance filter.  Only fixed nodes keep their rows (DynamicAnalysis.cpp:136-149).                     */
            memset(Reac, 0, nT * sizeof(double));
            int off = 0;
            for (int q = 0; q < m->n_mass; q++) {
                int nd = m->mass_node[q];
                for (int p = m->node_ptr[nd]; p < m->node_ptr[nd + 1]; p++, off++) Reac[m->totaldof[p]] += m->mass_val[off] * A[m->totaldof[p]];
            }
            double *Me = (double *)malloc(72 * 72 * sizeof(double)), *Ce = (double *)malloc(72 * 72 * sizeof(double));
            for (int e = 0; e < nE; e++) {
                int nn = elem_nn(m->elem_kind[e]), fixed = 0;
                for (int i = 0; i < nn && !fixed; i++) fixed = node_fixed[m->elem_conn[8 * e + i]];
                if (!fixed) continue;
                double fe[72];
                elem_fint(m, e, &rt[e], U, fe);
                elem_MC(m, e, &rt[e], Me, Ce);
                int nd = rt[e].nd;
                for (int i = 0; i < nd; i++) {
                    double v = fe[i];
                    for (int j = 0; j < nd; j++) v += Ce[i * nd + j] * V[rt[e].dofs[j]] + Me[i * nd + j] * A[rt[e].dofs[j]];
                    Reac[rt[e].dofs[i]] += v;
                }
            }
            free(Me); free(Ce);
            for (int i = 0; i < nT; i++) Reac[i] = is_fixed_dof[i] ? Reac[i] - Fext[i] : 0.0;
        }
        const double *src = field == 0 ? U : field == 1 ? V : field == 2 ? A : Reac;
        for (int q = 0; q < n_rec; q++) out[(size_t)(k - 1) * n_rec + q] = src[rec_dofs[q]];
    }
    if (Ufinal) memcpy(Ufinal, U, nT * sizeof(double));
done:
    for (int e = 0; e < nE; e++) free(rt[e].Kpml);
    free(rt); free(U); free(V); free(A); free(Up); free(Fint); free(Fext); free(Ftmp); free(rhs);
    free(dUt); free(Utr); free(Feff); free(dU); free(fe_all); free(lM.t); free(lC.t); free(lKp.t);
    free(lKm.t); free(lF.t); csr_free(&Kp); csr_free(&Km); csr_free(&T); csr_free(&Kf); csr_free(&Cs); csr_free(&Gs); free(lG.t); free(Ubar); free(Gtmp);
    free(cidx); free(Kdiag); free(Kc); free(bc); free(sky_lo); free(sky_hi); free(SM); free(Reac); free(node_fixed); free(is_fixed_dof);
    return rc;
}

int svlo_run_central_difference(const svlo_model *m, int nt, int field, int n_rec,
                                const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads) {
    return run_dynamic(m, 0, nt, field, n_rec, rec_dofs, out, Ufinal, nthreads);
}
int svlo_run_newmark(const svlo_model *m, int nt, int field, int n_rec,
                     const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads) {
    return run_dynamic(m, 1, nt, field, n_rec, rec_dofs, out, Ufinal, nthreads);
}
int svlo_run_extended_newmark(const svlo_model *m, int nt, int field, int n_rec,
                              const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads) {
    return run_dynamic(m, 2, nt, field, n_rec, rec_dofs, out, Ufinal, nthreads);
}
/* NewmarkBeta + NewtonRaphson (plastic materials): tol / nmax / flag = the JSON's cnvgtol / nstep / cnvgtest (Driver.hpp:1792-1795) */
int svlo_run_newmark_newton(const svlo_model *m, double tol, int nmax, int flag, int nt, int field, int n_rec,
                            const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads) {
    svlo_newton alg = {tol, nmax, flag};
    return run_dynamic_alg(m, 1, &alg, nt, field, n_rec, rec_dofs, out, Ufinal, nthreads);
}
