// EGM08 spherical-harmonic gravity field (device), used by egm08_force_kernel only (graph_kernels.cu).
#pragma once
#include "sixdof_device.cuh"

namespace b200 {

// libs/nox-py/python/elodin/egm08.py (EGM08.compute_field) in the operation order of
// oracle/sixdof_oracle.c:eff_gravity_egm08 — column by column (m outer, l = m..L inner), one running accumulator per
// component, every IEEE operation explicit — so the result is bit-identical to the oracle in both math modes.
// tab = [C | S | n1 | n2 | nq1 | nq2] each (L+1)^2 row-major [l][m], then diag[L+1], offc[L+1] (built at create by
// sixdof_abi.cu:egm08_tables with the oracle's formulas).  Returns the force (mass included): ~30 k instructions per
// evaluation at degree 64 and a 130-entry local array — which is why it lives in its own kernel and the body kernels
// only add its result.
static __device__ __forceinline__ Vec3 egm08_field(const double *__restrict__ tab, int L, double mu, double r_ref, Vec3 p, double mass)
{
    using namespace ex;
    const int n = L + 1;
    const double *C = tab, *S = C + n * n, *n1 = S + n * n, *n2 = n1 + n * n, *nq1 = n2 + n * n, *nq2 = nq1 + n * n;
    const double *diag = nq2 + n * n, *offc = diag + n;
    const double r = sqr(add(add(mul(p.x, p.x), mul(p.y, p.y)), mul(p.z, p.z)));
    const double s = div(p.x, r), t = div(p.y, r), u = div(p.z, r);
    double w[130]; // w[l] = rho_{l+1} / r_ref, rho_l = (mu / r) (r_ref / r)^l by repeated multiplication; w[L] = 0
    {
        double rho = div(mu, r);
        const double q = div(r_ref, r);
        for (int l = 1; l <= L; ++l) { rho = mul(rho, q); w[l - 1] = div(rho, r_ref); }
        w[L] = div(0.0, r_ref);
    }
    double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    double im_prev = 0.0, rm_prev = 0.0, im = 0.0, rm = 1.0;
    for (int m = 0; m <= L; ++m) {
        if (m > 0) {
            const double i_new = add(mul(s, im), mul(t, rm)), r_new = sub(mul(s, rm), mul(t, im));
            im_prev = im; rm_prev = rm; im = i_new; rm = r_new;
        }
        const double rm1 = m == 0 ? 0.0 : rm_prev, im1 = m == 0 ? 0.0 : im_prev;
        const double mp = m == L ? 0.0 : (double)(m + 1);
        double A0 = 0.0, A1 = 0.0, B0 = 0.0, B1 = 0.0;
        for (int l = m; l <= L; ++l) {
            double Al;
            if (l == m) Al = diag[m];
            else if (l == m + 1) Al = mul(offc[l], u);
            else Al = sub(mul(mul(u, n1[l * n + m]), A0), mul(n2[l * n + m], A1));
            A1 = A0; A0 = Al;
            double Bl = 0.0;
            if (m + 1 <= L) {
                if (l == m + 1) Bl = diag[m + 1];
                else if (l == m + 2) Bl = mul(offc[l], u);
                else if (l > m + 2) Bl = sub(mul(mul(u, n1[l * n + m + 1]), B0), mul(n2[l * n + m + 1], B1));
            }
            double Bn = 0.0;
            if (m + 1 <= L && l + 1 <= L) {
                const int l1 = l + 1;
                if (l1 == m + 1) Bn = diag[m + 1];
                else if (l1 == m + 2) Bn = mul(offc[l1], u);
                else Bn = sub(mul(mul(u, n1[l1 * n + m + 1]), Bl), mul(n2[l1 * n + m + 1], B0));
            }
            B1 = B0; B0 = Bl;
            const double wl = w[l];
            const double c = C[l * n + m], sv = S[l * n + m];
            const double ee = add(mul(c, rm1), mul(sv, im1)), ff = sub(mul(sv, rm1), mul(c, im1)), dd = add(mul(c, rm), mul(sv, im));
            a1 = add(a1, mul(mul(mul(wl, Al), mp), ee));
            a2 = add(a2, mul(mul(mul(wl, Al), mp), ff));
            a3 = add(a3, mul(mul(mul(mul(wl, Bl), mp), nq1[l * n + m]), dd));
            a4 = add(a4, mul(mul(mul(mul(mul(wl, Bn), mp), nq2[l * n + m]), dd), -1.0));
        }
    }
    return Vec3{mul(mass, add(a1, mul(s, a4))), mul(mass, add(a2, mul(t, a4))), mul(mass, add(a3, mul(u, a4)))};
}

} // namespace b200
