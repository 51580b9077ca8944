// EGM08 spherical-harmonic gravity field (device), used by egm08_force_kernel only (graph_kernels.cu).
#pragma once
#include "sixdof_device.cuh"

namespace b200 {

// libs/nox-py/python/elodin/egm08.py (EGM08.compute_field) in the operation order of
// oracle/sixdof_oracle.c:eff_gravity_egm08 — column by column (m outer, l = m..L inner), one running accumulator per
// component, every IEEE operation explicit — so the result is bit-identical to the oracle in both math modes.
// tab = the term stream sixdof_abi.cu:egm08_tables builds at create (eight doubles per (m, l) term in consumption
// order: the warp-uniform loads of an evaluation walk 64 contiguous bytes per term).  The oracle evaluates the B
// recursion twice per term (B_l and B_{l+1}); B_{l+1} of one term is bit for bit B_l of the next, so it is carried
// instead.  Returns the force (mass included): ~35 FP64 instructions per term, 2145 terms at degree 64, and a local
// array of the L+1 radial factors — which is why it lives in its own kernel and the body kernels only add its result.
static __device__ __forceinline__ Vec3 egm08_field(const double *__restrict__ tab, int L, double mu, double r_ref, Vec3 p, double mass)
{
    using namespace ex;
    const double r = sqr(add(add(mul(p.x, p.x), mul(p.y, p.y)), mul(p.z, p.z)));
    const double s = div(p.x, r), t = div(p.y, r), u = div(p.z, r);
    double w[129]; // w[l] = rho_{l+1} / r_ref, rho_l = (mu / r) (r_ref / r)^l by repeated multiplication; w[L] = 0
    {
        double rho = div(mu, r);
        const double q = div(r_ref, r);
        for (int l = 1; l <= L; ++l) { rho = mul(rho, q); w[l - 1] = div(rho, r_ref); }
        w[L] = div(0.0, r_ref);
    }
    const double4 *rec = reinterpret_cast<const double4 *>(tab); // two per term (cudaMalloc'd: 256-byte aligned)
    double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    double im_prev = 0.0, rm_prev = 0.0, im = 0.0, rm = 1.0;
    for (int m = 0; m <= L; ++m) {
        if (m > 0) {
            const double i_new = add(mul(s, im), mul(t, rm)), r_new = sub(mul(s, rm), mul(t, im));
            im_prev = im; rm_prev = rm; im = i_new; rm = r_new;
        }
        const double rm1 = m == 0 ? 0.0 : rm_prev, im1 = m == 0 ? 0.0 : im_prev;
        const double mp = m == L ? 0.0 : (double)(m + 1);
        // one term: Al = A_l at order m, Bl / Bn = B_l / B_{l+1} at order m+1, rc = C, S, nq1, nq2 of (l, m)
        auto term = [&](double Al, double Bl, double Bn, const double4 rc, double wl) {
            const double ee = add(mul(rc.x, rm1), mul(rc.y, im1)), ff = sub(mul(rc.y, rm1), mul(rc.x, im1)), dd = add(mul(rc.x, rm), mul(rc.y, im));
            const double wa = mul(mul(wl, Al), mp);
            a1 = add(a1, mul(wa, ee));
            a2 = add(a2, mul(wa, ff));
            a3 = add(a3, mul(mul(mul(mul(wl, Bl), mp), rc.z), dd));
            a4 = sub(a4, mul(mul(mul(mul(wl, Bn), mp), rc.w), dd)); // the oracle adds the product times -1.0: the same value
        };
        // l = m: A_m = diag[m], B_m = 0, B_{m+1} = diag[m+1] (0 beyond degree L)
        double4 ra = rec[0];
        double A1 = 0.0, A0 = ra.x, Bl = 0.0, Bn = m < L ? ra.z : 0.0;
        term(A0, Bl, Bn, rec[1], w[m]);
        rec += 2;
        if (m == L) break;
        // l = m + 1: A = offc[l] u, B_{l+1} = offc[l+1] u
        ra = rec[0];
        A1 = A0; A0 = mul(ra.x, u);
        Bl = Bn; Bn = m + 1 < L ? mul(ra.z, u) : 0.0;
        term(A0, Bl, Bn, rec[1], w[m + 1]);
        rec += 2;
        if (m + 1 == L) continue;
        // m + 2 <= l < L: both three-term recursions
#pragma unroll 2
        for (int l = m + 2; l < L; ++l) {
            ra = rec[0];
            const double Al = sub(mul(mul(u, ra.x), A0), mul(ra.y, A1));
            A1 = A0; A0 = Al;
            const double Bq = sub(mul(mul(u, ra.z), Bn), mul(ra.w, Bl)); // B_{l+1} from B_l (= the carried Bn) and B_{l-1}
            Bl = Bn; Bn = Bq;
            term(Al, Bl, Bn, rec[1], w[l]);
            rec += 2;
        }
        // l = L: B_{L+1} lies beyond the table
        ra = rec[0];
        const double Al = sub(mul(mul(u, ra.x), A0), mul(ra.y, A1));
        term(Al, Bn, 0.0, rec[1], w[L]);
        rec += 2;
    }
    return Vec3{mul(mass, add(a1, mul(s, a4))), mul(mass, add(a2, mul(t, a4))), mul(mass, add(a3, mul(u, a4)))};
}

} // namespace b200
