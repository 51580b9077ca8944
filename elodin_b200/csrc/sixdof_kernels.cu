// sm_100a kernels of the six_dof() hot path.
//
//   body_exact_kernel  K1/K2/K4/K5 of SURVEY §2.4 in EXACT arithmetic: one thread per
//                      body, the whole tick (clear_forces, effectors x4, calc_accel x4,
//                      stage advance x4, final combine, renormalise) in registers,
//                      n_ticks ticks per launch.
//   body_fast_kernel   the same tick restructured for the FP64 pipe / HBM roofline.
//   graph_*_kernel     K3: GraphQuery.edge_fold gravity at the 3 distinct stage
//                      positions of a tick (they depend on x0, v0 only — see below).
//   aos_to_soa / soa_to_aos   K6: host column layout <-> device planes.
//
// Why one launch per tick suffices even with body-body coupling: in the
// reference's RK4 (libs/nox-py/src/integrator/rk4.rs:85-111) every stage position
// is x0 (+) (dt*f)*v0 — it never depends on a stage acceleration — so the gravity
// at all four stages (three distinct positions, f = 0, .5, 1) is a function of
// the tick's input state alone and needs no grid-wide synchronisation.
#include <algorithm>
#include <mutex>
#include <utility>
#include <vector>
#include <cstdlib>

#include "sixdof_device.cuh"
#include "sixdof_internal.h"

namespace b200 {


// ------------------------------------------------------------------ column access
__device__ __forceinline__ double ldp(const double *base, uint64_t ld, int plane, uint64_t b)
{
    return base[(uint64_t)plane * ld + b];
}
__device__ __forceinline__ void stp(double *base, uint64_t ld, int plane, uint64_t b, double v)
{
    base[(uint64_t)plane * ld + b] = v;
}

__device__ __forceinline__ Pose load_pose(const double *p, uint64_t ld, uint64_t b)
{
    Pose o;
    o.q = Quat{ldp(p, ld, 0, b), ldp(p, ld, 1, b), ldp(p, ld, 2, b), ldp(p, ld, 3, b)};
    o.x = Vec3{ldp(p, ld, 4, b), ldp(p, ld, 5, b), ldp(p, ld, 6, b)};
    return o;
}
__device__ __forceinline__ Motion load_motion(const double *p, uint64_t ld, uint64_t b)
{
    Motion m;
    m.ang = Vec3{ldp(p, ld, 0, b), ldp(p, ld, 1, b), ldp(p, ld, 2, b)};
    m.lin = Vec3{ldp(p, ld, 3, b), ldp(p, ld, 4, b), ldp(p, ld, 5, b)};
    return m;
}
__device__ __forceinline__ Inertia load_inertia(const double *p, uint64_t ld, uint64_t b)
{
    Inertia I;
    I.diag = Vec3{ldp(p, ld, 0, b), ldp(p, ld, 1, b), ldp(p, ld, 2, b)};
    I.m = ldp(p, ld, 6, b);
    return I;
}
__device__ __forceinline__ void store_pose(double *p, uint64_t ld, uint64_t b, const Pose &o)
{
    stp(p, ld, 0, b, o.q.i); stp(p, ld, 1, b, o.q.j); stp(p, ld, 2, b, o.q.k); stp(p, ld, 3, b, o.q.w);
    stp(p, ld, 4, b, o.x.x); stp(p, ld, 5, b, o.x.y); stp(p, ld, 6, b, o.x.z);
}
__device__ __forceinline__ void store_motion(double *p, uint64_t ld, uint64_t b, const Motion &m)
{
    stp(p, ld, 0, b, m.ang.x); stp(p, ld, 1, b, m.ang.y); stp(p, ld, 2, b, m.ang.z);
    stp(p, ld, 3, b, m.lin.x); stp(p, ld, 4, b, m.lin.y); stp(p, ld, 5, b, m.lin.z);
}

// slot of the telemetry sample due after `tick_after` ticks, if any
__device__ __forceinline__ bool traj_due(const StepParams &P, uint64_t tick_after, uint64_t &slot)
{
    if (P.traj_every == 0 || (tick_after % P.traj_every) != 0) return false;
    slot = tick_after / P.traj_every - 1;
    return slot < P.traj_capacity;
}
__device__ __forceinline__ void traj_store_state(const StepParams &P, uint64_t b, uint64_t slot, const Pose &x, const Motion &v)
{
    double *t = P.traj + slot * (uint64_t)P.traj_planes * P.ld;
    store_pose(t, P.ld, b, x);
    store_motion(t + 7ull * P.ld, P.ld, b, v);
}
// B200_TRAJ_FULL: WorldAccel and Force as the tick leaves them in the ECS columns
__device__ __forceinline__ void traj_store_af(const StepParams &P, uint64_t b, uint64_t slot, const Motion &a, const Motion &f)
{
    double *t = P.traj + (slot * (uint64_t)P.traj_planes + 13ull) * P.ld;
    store_motion(t, P.ld, b, a);
    store_motion(t + 6ull * P.ld, P.ld, b, f);
}

// ================================================================== EXACT body kernel

// edge_fold gravity of one body at the three stage positions, held in registers by the kernels that
// compute it themselves (small_world_kernel) instead of reading the gforce planes
struct GravReg {
    Vec3 g0, g1, g2;
    bool has; // the body owns >= 1 out-edge
};
__device__ __forceinline__ Vec3 grav_slot(const GravReg &g, int slot) { return slot == 0 ? g.g0 : (slot == 1 ? g.g1 : g.g2); }

// clear_forces | effectors (array order) on the stage state; six_dof.rs:148-150,195
template <bool GREG>
__device__ __forceinline__ Motion effectors_exact(const StepParams &P, uint64_t b, int slot, const Pose &sx,
                                                  const ex::PoseInv &pi, const Motion &sv, const Inertia &I,
                                                  const GravReg &greg)
{
    using namespace ex;
    Motion F = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    for (uint32_t e = 0; e < P.n_eff; ++e) {
        const EffDev &E = P.eff[e];
        if (E.mask && !E.mask[b % P.n_entities]) continue; // entity does not own the effector's components (query join)
        switch (E.kind) {
        case B200_EFF_GRAVITY_CONST: { // ball/sim.py:56-58: f + SpatialForce(linear=g*m)
            F.ang = Vec3{add(F.ang.x, 0.0), add(F.ang.y, 0.0), add(F.ang.z, 0.0)};
            F.lin = Vec3{add(F.lin.x, mul(E.p[0], I.m)), add(F.lin.y, mul(E.p[1], I.m)),
                         add(F.lin.z, mul(E.p[2], I.m))};
            break;
        }
        case B200_EFF_DRAG_QUADRATIC: { // ball/sim.py:99-116; result torque is zero
            double w0 = 0.0, w1 = 0.0, w2 = 0.0;
            if (E.col) { w0 = ldp(E.col, P.ld, 0, b); w1 = ldp(E.col, P.ld, 1, b); w2 = ldp(E.col, P.ld, 2, b); }
            const Vec3 fl = {sub(w0, sv.lin.x), sub(w1, sv.lin.y), sub(w2, sv.lin.z)};
            const double speed = sqr(dot3(fl));
            const double cd_rho = E.col_width == 5 ? ldp(E.col, P.ld, 3, b) : E.p[0];
            const double area = E.col_width == 5 ? ldp(E.col, P.ld, 4, b) : E.p[1];
            const double drag = mul(0.5, mul(mul(cd_rho, mul(speed, speed)), area));
            F.ang = Vec3{0.0, 0.0, 0.0};
            F.lin = Vec3{add(F.lin.x, mul(drag, div(fl.x, speed))), add(F.lin.y, mul(drag, div(fl.y, speed))),
                         add(F.lin.z, mul(drag, div(fl.z, speed)))};
            break;
        }
        case B200_EFF_THRUST_BODY: { // rocket/main.py:429-431
            const double t = E.col ? ldp(E.col, P.ld, 0, b) : 0.0;
            const Vec3 d = qrot_with(sx.q, pi.qi, Vec3{E.p[0], E.p[1], E.p[2]});
            F.ang = Vec3{add(F.ang.x, 0.0), add(F.ang.y, 0.0), add(F.ang.z, 0.0)};
            F.lin = Vec3{add(F.lin.x, mul(d.x, t)), add(F.lin.y, mul(d.y, t)), add(F.lin.z, mul(d.z, t))};
            break;
        }
        case B200_EFF_WRENCH_BODY: { // rocket/main.py:407-413, falcon9/sim.py:659-672
            Vec3 a = {0.0, 0.0, 0.0}, c = {0.0, 0.0, 0.0};
            if (E.col) {
                a = Vec3{ldp(E.col, P.ld, 0, b), ldp(E.col, P.ld, 1, b), ldp(E.col, P.ld, 2, b)};
                c = Vec3{ldp(E.col, P.ld, 3, b), ldp(E.col, P.ld, 4, b), ldp(E.col, P.ld, 5, b)};
            }
            const bool lin_first = (E.flags & B200_EFF_FLAG_WRENCH_LINEAR_FIRST) != 0;
            const Vec3 tw = qrot_with(sx.q, pi.qi, lin_first ? c : a);
            const Vec3 fw = qrot_with(sx.q, pi.qi, lin_first ? a : c);
            F.ang = Vec3{add(F.ang.x, tw.x), add(F.ang.y, tw.y), add(F.ang.z, tw.z)};
            F.lin = Vec3{add(F.lin.x, fw.x), add(F.lin.y, fw.y), add(F.lin.z, fw.z)};
            break;
        }
        case B200_EFF_GRAVITY_FRAME: { // falcon9/sim.py:350-361, frames.py:91-109
            const double mu = E.p[0];
            const Vec3 om = {E.p[1], E.p[2], E.p[3]};
            const Vec3 r = sx.x, v = sv.lin;
            const double rn = sqr(dot3(r));
            const double rn3 = mul(mul(rn, rn), rn);
            const Vec3 g = {div(mul(-mu, r.x), rn3), div(mul(-mu, r.y), rn3), div(mul(-mu, r.z), rn3)};
            const Vec3 c = cross(om, v);
            const Vec3 c2 = cross(om, cross(om, r));
            const Vec3 acc = {add(g.x, add(mul(-2.0, c.x), -c2.x)), add(g.y, add(mul(-2.0, c.y), -c2.y)),
                              add(g.z, add(mul(-2.0, c.z), -c2.z))};
            F.ang = Vec3{add(F.ang.x, 0.0), add(F.ang.y, 0.0), add(F.ang.z, 0.0)};
            F.lin = Vec3{add(F.lin.x, mul(acc.x, I.m)), add(F.lin.y, mul(acc.y, I.m)), add(F.lin.z, mul(acc.z, I.m))};
            break;
        }
        case B200_EFF_GRAVITY_EDGES_NEWTON:
        case B200_EFF_GRAVITY_EDGES_SOFTENED: { // Force := edge_fold(init 0) for bodies that own an edge
            if (GREG) {
                if (greg.has) { F.ang = Vec3{0.0, 0.0, 0.0}; F.lin = grav_slot(greg, slot); }
            } else if (P.gforce && P.has_edge && P.has_edge[b % P.n_entities]) {
                F.ang = Vec3{0.0, 0.0, 0.0};
                F.lin = Vec3{ldp(P.gforce, P.ld, slot * 3 + 0, b), ldp(P.gforce, P.ld, slot * 3 + 1, b),
                             ldp(P.gforce, P.ld, slot * 3 + 2, b)};
            }
            break;
        }
        default: break;
        }
    }
    return F;
}

// one tick of one body in EXACT arithmetic (state in registers)
template <int INTEG, bool GREG>
__device__ __forceinline__ void exact_tick(const StepParams &P, uint64_t b, Pose &x0, Motion &v0, Motion &a_out,
                                           Motion &f_out, const Inertia &I, const GravReg &greg)
{
    using namespace ex;
    if (INTEG == B200_INTEGRATOR_RK4) {
        // rk4.rs:85-123 (see the header comment of oracle/sixdof_oracle.c for the derivation)
        Motion sa = a_out; // du.a before stage 1 is the WorldAccel column
        Motion kv, ka;
        // three distinct stage poses (f = 0, .5, 1), functions of (x0, v0) only; stages 2 and 3 share
        // the f = .5 pose and its inverses — identical inputs, identical bits — so each is built once
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const double dtf = mul(P.dt_stage, k == 0 ? 0.0 : (k == 1 ? 0.5 : 1.0));
            const Pose sx = tadd(x0, scale(dtf, v0));
            const PoseInv pi = pose_inverses(sx.q);
            const int n_stages = (k == 1) ? 2 : 1;
#pragma unroll 1
            for (int j = 0; j < n_stages; ++j) {
                const int s = (k == 0) ? 0 : (k == 1 ? 1 + j : 3);
                const Motion sv = madd(v0, scale(dtf, sa));
                f_out = effectors_exact<GREG>(P, b, k, sx, pi, sv, I, greg);
                sa = calc_accel_with(sx, pi, f_out, I);
                if (s == 0) { kv = sv; ka = sa; }
                else if (s == 3) { kv = madd(kv, sv); ka = madd(ka, sa); }
                else { kv = madd(kv, scale(2.0, sv)); ka = madd(ka, scale(2.0, sa)); }
            }
        }
        const double c = mul(P.dt_final, 1.0 / 6.0);
        x0 = tadd(x0, scale(c, kv));
        v0 = madd(v0, scale(c, ka));
        a_out = sa;
    } else {
        // semi_implicit.rs:42-62
        const PoseInv pi = pose_inverses(x0.q);
        f_out = effectors_exact<GREG>(P, b, 0, x0, pi, v0, I, greg);
        a_out = calc_accel_with(x0, pi, f_out, I);
        v0 = madd(v0, scale(P.dt_final, a_out));
        x0 = tadd(x0, scale(P.dt_final, v0));
    }
}

template <int INTEG, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) body_exact_kernel(const __grid_constant__ StepParams P)
{
    const uint64_t b = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (b >= P.n_bodies) return;

    Pose x0 = load_pose(P.pos, P.ld, b);
    Motion v0 = load_motion(P.vel, P.ld, b);
    Motion a_out = load_motion(P.acc, P.ld, b);
    Motion f_out = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    const Inertia I = load_inertia(P.ine, P.ld, b);
    const GravReg no_greg{};

    for (uint32_t t = 0; t < P.n_ticks; ++t) {
        exact_tick<INTEG, false>(P, b, x0, v0, a_out, f_out, I, no_greg);
        uint64_t slot;
        if (traj_due(P, P.tick0 + t + 1, slot)) {
            traj_store_state(P, b, slot, x0, v0);
            if (P.traj_planes == 25) traj_store_af(P, b, slot, a_out, f_out);
        }
    }
    store_pose(P.pos, P.ld, b, x0);
    store_motion(P.vel, P.ld, b, v0);
    store_motion(P.acc, P.ld, b, a_out);
    store_motion(P.frc, P.ld, b, f_out);
}

// ================================================================== FAST body kernel

// Everything the effector list contributes, folded once per launch:
//   F_lin(stage) = fw + R(q) fb + drag(v) + m*frame(x, v) + gforce[slot]
//   a_ang(stage) = R(q) u,   u = (sum of body-frame torques) / diag(I)   (R^-1 then R cancel)
struct Folded {
    Vec3 fw;      // world-frame constant force (GRAVITY_CONST: g*m)
    Vec3 fb;      // body-frame force (THRUST_BODY axis*thrust, WRENCH_BODY force part)
    Vec3 u;       // body-frame angular acceleration
    Vec3 wind;    // DRAG_QUADRATIC
    double kd;    // 0.5*Cd*rho*A
    double mu;    // GRAVITY_FRAME
    Vec3 om;
    bool drag, frame, graph;
};

template <bool GREG>
__device__ __forceinline__ Folded fold_effectors(const StepParams &P, uint64_t b, const Inertia &I, const Vec3 &invI,
                                                 const GravReg &greg)
{
    Folded f;
    f.fw = f.fb = f.u = f.wind = f.om = Vec3{0.0, 0.0, 0.0};
    f.kd = f.mu = 0.0;
    f.drag = f.frame = f.graph = false;
    Vec3 tb = {0.0, 0.0, 0.0};
    for (uint32_t e = 0; e < P.n_eff; ++e) {
        const EffDev &E = P.eff[e];
        if (E.mask && !E.mask[b % P.n_entities]) continue; // query join: not a member
        switch (E.kind) {
        case B200_EFF_GRAVITY_CONST:
            f.fw.x = fma(E.p[0], I.m, f.fw.x); f.fw.y = fma(E.p[1], I.m, f.fw.y); f.fw.z = fma(E.p[2], I.m, f.fw.z);
            break;
        case B200_EFF_DRAG_QUADRATIC:
            f.drag = true;
            f.kd = E.col_width == 5 ? 0.5 * ldp(E.col, P.ld, 3, b) * ldp(E.col, P.ld, 4, b) : 0.5 * E.p[0] * E.p[1];
            if (E.col) f.wind = Vec3{ldp(E.col, P.ld, 0, b), ldp(E.col, P.ld, 1, b), ldp(E.col, P.ld, 2, b)};
            tb = Vec3{0.0, 0.0, 0.0}; // the reference's apply_drag returns SpatialForce(linear=...): torque reset
            break;
        case B200_EFF_THRUST_BODY: {
            const double t = E.col ? ldp(E.col, P.ld, 0, b) : 0.0;
            f.fb.x = fma(E.p[0], t, f.fb.x); f.fb.y = fma(E.p[1], t, f.fb.y); f.fb.z = fma(E.p[2], t, f.fb.z);
            break;
        }
        case B200_EFF_WRENCH_BODY:
            if (E.col) {
                const int to = (E.flags & B200_EFF_FLAG_WRENCH_LINEAR_FIRST) ? 3 : 0;
                const int fo = 3 - to;
                tb.x += ldp(E.col, P.ld, to + 0, b); tb.y += ldp(E.col, P.ld, to + 1, b); tb.z += ldp(E.col, P.ld, to + 2, b);
                f.fb.x += ldp(E.col, P.ld, fo + 0, b); f.fb.y += ldp(E.col, P.ld, fo + 1, b); f.fb.z += ldp(E.col, P.ld, fo + 2, b);
            }
            break;
        case B200_EFF_GRAVITY_FRAME:
            f.frame = true;
            f.mu = E.p[0];
            f.om = Vec3{E.p[1], E.p[2], E.p[3]};
            break;
        case B200_EFF_GRAVITY_EDGES_NEWTON:
        case B200_EFF_GRAVITY_EDGES_SOFTENED: // host guarantees this is effector 0 in FAST mode
            f.graph = GREG ? greg.has : (P.gforce && P.has_edge && P.has_edge[b % P.n_entities]);
            break;
        default: break;
        }
    }
    f.u = Vec3{tb.x * invI.x, tb.y * invI.y, tb.z * invI.z};
    return f;
}

// linear acceleration of one stage: everything that depends on (q, x, v)
template <bool GREG>
__device__ __forceinline__ Vec3 lin_accel_fast(const StepParams &P, const Folded &f, uint64_t b, int slot,
                                               const Vec3 &fbw, const Vec3 &x, const Vec3 &v, double m, double inv_m,
                                               const GravReg &greg)
{
    Vec3 F = {f.fw.x + fbw.x, f.fw.y + fbw.y, f.fw.z + fbw.z};
    if (f.drag) {
        const Vec3 fl = {f.wind.x - v.x, f.wind.y - v.y, f.wind.z - v.z};
        const double s2 = fl.x * fl.x + fl.y * fl.y + fl.z * fl.z;
        // drag*dir = (kd*speed^2) * fl/speed; speed == 0 is 0/0 = NaN in the reference too
        const double k = (s2 == 0.0) ? __longlong_as_double(0x7ff8000000000000ll) : f.kd * sqrt(s2);
        F.x = fma(k, fl.x, F.x); F.y = fma(k, fl.y, F.y); F.z = fma(k, fl.z, F.z);
    }
    if (f.frame) {
        const double r2 = x.x * x.x + x.y * x.y + x.z * x.z;
        const double ir = fa::rsqrt_nr(r2);
        const double g = -f.mu * ir * ir * ir;
        const Vec3 c = fa::cross(f.om, v);
        const Vec3 c2 = fa::cross(f.om, fa::cross(f.om, x));
        F.x = fma(fma(g, x.x, -2.0 * c.x - c2.x), m, F.x);
        F.y = fma(fma(g, x.y, -2.0 * c.y - c2.y), m, F.y);
        F.z = fma(fma(g, x.z, -2.0 * c.z - c2.z), m, F.z);
    }
    if (f.graph) {
        if (GREG) {
            const Vec3 g = grav_slot(greg, slot);
            F.x += g.x; F.y += g.y; F.z += g.z;
        } else {
            F.x += ldp(P.gforce, P.ld, slot * 3 + 0, b);
            F.y += ldp(P.gforce, P.ld, slot * 3 + 1, b);
            F.z += ldp(P.gforce, P.ld, slot * 3 + 2, b);
        }
    }
    return Vec3{F.x * inv_m, F.y * inv_m, F.z * inv_m};
}

// world-frame force this stage's state produced (only materialised when Force is written back)
__device__ __forceinline__ Motion force_out_fast(const Vec3 &a_lin, const Vec3 &a_ang_body_u, const Quat &q,
                                                 const Inertia &I)
{
    // torque_world = R (I .* u)
    const Vec3 tb = {a_ang_body_u.x * I.diag.x, a_ang_body_u.y * I.diag.y, a_ang_body_u.z * I.diag.z};
    Motion F;
    F.ang = fa::rot(q, tb);
    F.lin = Vec3{a_lin.x * I.m, a_lin.y * I.m, a_lin.z * I.m};
    return F;
}

// n_ticks ticks of one body, state in registers (shared by the direct and the TMA-pipelined kernel)
// (n_ticks, tick0, want_f) are P.n_ticks, P.tick0, P.write_fa for the kernels that integrate a launch's ticks
// in one call; small_world_kernel calls it once per tick with that tick's gravity in `greg`
template <int INTEG, bool TRAJ, bool GREG = false>
__device__ __forceinline__ void fast_ticks(const StepParams &P, uint64_t b, Pose &x0, Motion &v0, const Inertia &I,
                                           Motion &a_last, Motion &f_last, uint32_t n_ticks, uint64_t tick0, bool want_f,
                                           const GravReg &greg)
{
    const Vec3 invI = {fa::rcp_nr(I.diag.x), fa::rcp_nr(I.diag.y), fa::rcp_nr(I.diag.z)};
    const double inv_m = fa::rcp_nr(I.m);
    const Folded f = fold_effectors<GREG>(P, b, I, invI, greg);

    a_last = Motion{{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    Quat q_last = x0.q;
    const double dt = P.dt_stage;
    const bool has_u = (f.u.x != 0.0) | (f.u.y != 0.0) | (f.u.z != 0.0);
    const bool has_fb = (f.fb.x != 0.0) | (f.fb.y != 0.0) | (f.fb.z != 0.0);

    for (uint32_t t = 0; t < n_ticks; ++t) {
        if (INTEG == B200_INTEGRATOR_RK4) {
            const Vec3 w0 = v0.ang, u0 = v0.lin;
            // the three distinct stage poses depend on (x0, v0) only (rk4.rs:85-111)
            // (the stage attitudes only matter to bodies that carry a body-frame force or torque)
            const double h2 = 0.25 * dt, h4 = 0.5 * dt;
            Quat q1 = x0.q, q2 = x0.q, q4 = x0.q;
            if (has_u | has_fb) {
                q1 = fa::normalize(x0.q); // x0 (+) 0*v0 still renormalises (spatial.rs:540-545)
                q2 = fa::advance(x0.q, Vec3{h2 * w0.x, h2 * w0.y, h2 * w0.z});
                q4 = fa::advance(x0.q, Vec3{h4 * w0.x, h4 * w0.y, h4 * w0.z});
            }
            const Vec3 x2 = {fma(h4, u0.x, x0.x.x), fma(h4, u0.y, x0.x.y), fma(h4, u0.z, x0.x.z)};
            const Vec3 x4 = {fma(dt, u0.x, x0.x.x), fma(dt, u0.y, x0.x.y), fma(dt, u0.z, x0.x.z)};
            // angular acceleration R(q) u and rotated body force, once per distinct attitude
            // (rotating an all-zero body torque / force is skipped: most bodies of a world carry
            // no body-frame wrench; NaNs compare unequal to zero and still take the full path)
            const Vec3 zero3 = {0.0, 0.0, 0.0};
            Vec3 aa1 = zero3, aa2 = zero3, aa4 = zero3, fb1 = zero3, fb2 = zero3, fb4 = zero3;
            if (has_u) { aa1 = fa::rot(q1, f.u); aa2 = fa::rot(q2, f.u); aa4 = fa::rot(q4, f.u); }
            if (has_fb) { fb1 = fa::rot(q1, f.fb); fb2 = fa::rot(q2, f.fb); fb4 = fa::rot(q4, f.fb); }
            // stage 1: v = v0
            const Vec3 al1 = lin_accel_fast<GREG>(P, f, b, 0, fb1, x0.x, u0, I.m, inv_m, greg);
            // stage 2: v = v0 + dt/2 a1
            const Vec3 u2 = {fma(h4, al1.x, u0.x), fma(h4, al1.y, u0.y), fma(h4, al1.z, u0.z)};
            const Vec3 al2 = lin_accel_fast<GREG>(P, f, b, 1, fb2, x2, u2, I.m, inv_m, greg);
            // stage 3: same pose as stage 2, v = v0 + dt/2 a2
            const Vec3 u3 = {fma(h4, al2.x, u0.x), fma(h4, al2.y, u0.y), fma(h4, al2.z, u0.z)};
            const Vec3 al3 = lin_accel_fast<GREG>(P, f, b, 1, fb2, x2, u3, I.m, inv_m, greg);
            // stage 4: v = v0 + dt a3
            const Vec3 u4 = {fma(dt, al3.x, u0.x), fma(dt, al3.y, u0.y), fma(dt, al3.z, u0.z)};
            const Vec3 al4 = lin_accel_fast<GREG>(P, f, b, 2, fb4, x4, u4, I.m, inv_m, greg);
            // k.v sum = 6 v0 + dt (a1 + a2 + a3);  k.a sum = a1 + 2 a2 + 2 a3 + a4   (a3.ang == a2.ang)
            const double c = P.dt_final * (1.0 / 6.0);
            const Vec3 kw = {fma(dt, aa1.x + 2.0 * aa2.x, 6.0 * w0.x), fma(dt, aa1.y + 2.0 * aa2.y, 6.0 * w0.y),
                             fma(dt, aa1.z + 2.0 * aa2.z, 6.0 * w0.z)};
            const Vec3 ku = {fma(dt, al1.x + al2.x + al3.x, 6.0 * u0.x), fma(dt, al1.y + al2.y + al3.y, 6.0 * u0.y),
                             fma(dt, al1.z + al2.z + al3.z, 6.0 * u0.z)};
            const double hc = 0.5 * c;
            x0.q = fa::advance(x0.q, Vec3{hc * kw.x, hc * kw.y, hc * kw.z});
            x0.x = Vec3{fma(c, ku.x, x0.x.x), fma(c, ku.y, x0.x.y), fma(c, ku.z, x0.x.z)};
            v0.ang = Vec3{fma(c, aa1.x + 4.0 * aa2.x + aa4.x, w0.x), fma(c, aa1.y + 4.0 * aa2.y + aa4.y, w0.y),
                          fma(c, aa1.z + 4.0 * aa2.z + aa4.z, w0.z)};
            v0.lin = Vec3{fma(c, al1.x + 2.0 * (al2.x + al3.x) + al4.x, u0.x),
                          fma(c, al1.y + 2.0 * (al2.y + al3.y) + al4.y, u0.y),
                          fma(c, al1.z + 2.0 * (al2.z + al3.z) + al4.z, u0.z)};
            a_last.ang = aa4; a_last.lin = al4; q_last = q4;
        } else {
            // semi_implicit.rs:42-62; calc_accel rotates by q/|q| whatever |q| is
            const double n2 = x0.q.i * x0.q.i + x0.q.j * x0.q.j + x0.q.k * x0.q.k + x0.q.w * x0.q.w;
            const double rn = fa::rsqrt_nr(n2);
            const Quat qn = {x0.q.i * rn, x0.q.j * rn, x0.q.k * rn, x0.q.w * rn};
            const Vec3 aa = has_u ? fa::rot(qn, f.u) : Vec3{0.0, 0.0, 0.0};
            const Vec3 fbw = has_fb ? fa::rot(qn, f.fb) : Vec3{0.0, 0.0, 0.0};
            const Vec3 al = lin_accel_fast<GREG>(P, f, b, 0, fbw, x0.x, v0.lin, I.m, inv_m, greg);
            const double d = P.dt_final;
            v0.ang = Vec3{fma(d, aa.x, v0.ang.x), fma(d, aa.y, v0.ang.y), fma(d, aa.z, v0.ang.z)};
            v0.lin = Vec3{fma(d, al.x, v0.lin.x), fma(d, al.y, v0.lin.y), fma(d, al.z, v0.lin.z)};
            const double hd = 0.5 * d;
            x0.q = fa::advance(x0.q, Vec3{hd * v0.ang.x, hd * v0.ang.y, hd * v0.ang.z});
            x0.x = Vec3{fma(d, v0.lin.x, x0.x.x), fma(d, v0.lin.y, x0.x.y), fma(d, v0.lin.z, x0.x.z)};
            a_last.ang = aa; a_last.lin = al; q_last = qn;
        }
        if (TRAJ) { // compiled out of the launches that record nothing (the roofline case)
            uint64_t slot;
            if (traj_due(P, tick0 + t + 1, slot)) {
                traj_store_state(P, b, slot, x0, v0);
                if (P.traj_planes == 25) traj_store_af(P, b, slot, a_last, force_out_fast(a_last.lin, f.u, q_last, I));
            }
        }
    }
    if (want_f) f_last = force_out_fast(a_last.lin, f.u, q_last, I);
}

// Effector columns are consumed inside the (uniform) effector switch, i.e. after the state
// loads and the reciprocal prologue; prefetching them first puts their HBM latency under the
// state loads instead of behind them.
__device__ __forceinline__ void prefetch_effector_columns(const StepParams &P, uint64_t b)
{
    for (uint32_t e = 0; e < P.n_eff; ++e) {
        const double *col = P.eff[e].col;
        if (!col) continue;
        const uint32_t w = P.eff[e].col_width;
        for (uint32_t k = 0; k < w; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(col + (uint64_t)k * P.ld + b));
    }
    if (P.gforce)
        for (uint32_t k = 0; k < 9; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.gforce + (uint64_t)k * P.ld + b));
}

template <int INTEG, int BLOCK, int MINB, bool TRAJ>
__global__ void __launch_bounds__(BLOCK, MINB) body_fast_kernel(const __grid_constant__ StepParams P)
{
    const uint64_t b = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (b >= P.n_bodies) return;
    prefetch_effector_columns(P, b);
    Pose x0 = load_pose(P.pos, P.ld, b);
    Motion v0 = load_motion(P.vel, P.ld, b);
    const Inertia I = load_inertia(P.ine, P.ld, b);
    Motion a_last, f_last;
    fast_ticks<INTEG, TRAJ>(P, b, x0, v0, I, a_last, f_last, P.n_ticks, P.tick0, P.write_fa != 0, GravReg{});
    store_pose(P.pos, P.ld, b, x0);
    store_motion(P.vel, P.ld, b, v0);
    if (P.write_fa) {
        store_motion(P.acc, P.ld, b, a_last);
        store_motion(P.frc, P.ld, b, f_last);
    }
}

// ------------------------------------------------------------------ TMA-pipelined persistent variant
//
// One CTA = kPipeTB threads = one tile of kPipeTB bodies at a time, looping over tiles
// (persistent grid sized to the SM count).  The 17 input planes of the next tiles
// (pos 7, vel 6, inertia diag 3 + mass) are fetched with cp.async.bulk into a
// kStages-deep shared-memory ring, completion signalled on mbarriers; the 13 output
// planes leave through a double-buffered shared-memory tile and cp.async.bulk stores.
// The FP64 work of tile i therefore overlaps the HBM traffic of tiles i+1.. and i-1,
// which the direct kernel (1 CTA/SM at 150+ registers) cannot do.
static constexpr int kPipeTB = 128;
static constexpr int kPipeIn = 17;
static constexpr int kPipeOut = 13;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// DIRECT_OUT: results leave with plain coalesced stores (no output tiles), which shrinks the CTA to
// STAGES x 17 KB of shared memory so that 4 CTAs/SM fit at 128 registers — the input ring then keeps
// ~17 KB per CTA in flight at all times, independent of how long the FP64 phase of a tile takes.
template <int INTEG, int STAGES, int MINB, bool DIRECT_OUT>
__global__ void __launch_bounds__(kPipeTB, MINB) body_fast_pipe_kernel(const __grid_constant__ StepParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kOutTiles = DIRECT_OUT ? 0 : 2;
    double(*sin)[kPipeIn][kPipeTB] = reinterpret_cast<double(*)[kPipeIn][kPipeTB]>(smem_raw);
    double(*sout)[kPipeOut][kPipeTB] =
        reinterpret_cast<double(*)[kPipeOut][kPipeTB]>(smem_raw + sizeof(double) * STAGES * kPipeIn * kPipeTB);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(double) * (STAGES * kPipeIn + kOutTiles * kPipeOut) * kPipeTB);

    const int tid = threadIdx.x;
    const uint64_t n_tiles = (P.n_bodies + kPipeTB - 1) / kPipeTB;
    constexpr uint32_t kPlaneBytes = kPipeTB * sizeof(double);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // source plane of input slot k for the tile starting at body b0
    auto src_plane = [&](int k, uint64_t b0) -> const double * {
        if (k < 7) return P.pos + (uint64_t)k * P.ld + b0;
        if (k < 13) return P.vel + (uint64_t)(k - 7) * P.ld + b0;
        return P.ine + (uint64_t)(k == 16 ? 6 : k - 13) * P.ld + b0;
    };
    auto issue_loads = [&](uint64_t tile, int s) { // warp 0
        if (tid == 0) mbar_expect_tx(&full[s], kPipeIn * kPlaneBytes);
        __syncwarp();
        if (tid < kPipeIn) bulk_g2s(&sin[s][tid][0], src_plane(tid, tile * kPipeTB), kPlaneBytes, &full[s]);
    };

    if (tid < 32)
        for (int s = 0; s < STAGES; ++s) {
            const uint64_t tile = blockIdx.x + (uint64_t)s * gridDim.x;
            if (tile < n_tiles) issue_loads(tile, s);
        }

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        const uint64_t b = tile * kPipeTB + tid;
        mbar_wait(&full[s], parity);
        Pose x0;
        Motion v0;
        Inertia I;
        x0.q = Quat{sin[s][0][tid], sin[s][1][tid], sin[s][2][tid], sin[s][3][tid]};
        x0.x = Vec3{sin[s][4][tid], sin[s][5][tid], sin[s][6][tid]};
        v0.ang = Vec3{sin[s][7][tid], sin[s][8][tid], sin[s][9][tid]};
        v0.lin = Vec3{sin[s][10][tid], sin[s][11][tid], sin[s][12][tid]};
        I.diag = Vec3{sin[s][13][tid], sin[s][14][tid], sin[s][15][tid]};
        I.m = sin[s][16][tid];
        // the output tile about to be written was handed to the async proxy two tiles ago
        if (!DIRECT_OUT && tid < kPipeOut) bulk_wait_read<1>();
        __syncthreads(); // everyone has drained stage s; out[it&1] is free
        if (tid < 32) {
            const uint64_t next = tile + (uint64_t)STAGES * gridDim.x;
            if (next < n_tiles) issue_loads(next, s);
        }
        Motion a_last, f_last;
        const bool live = b < P.n_bodies;
        if (live) fast_ticks<INTEG, true>(P, b, x0, v0, I, a_last, f_last, P.n_ticks, P.tick0, P.write_fa != 0, GravReg{});
        if (DIRECT_OUT) {
            if (live) {
                store_pose(P.pos, P.ld, b, x0);
                store_motion(P.vel, P.ld, b, v0);
                if (P.write_fa) {
                    store_motion(P.acc, P.ld, b, a_last);
                    store_motion(P.frc, P.ld, b, f_last);
                }
            }
            continue;
        }
        const int ob = it & 1;
        sout[ob][0][tid] = x0.q.i; sout[ob][1][tid] = x0.q.j; sout[ob][2][tid] = x0.q.k; sout[ob][3][tid] = x0.q.w;
        sout[ob][4][tid] = x0.x.x; sout[ob][5][tid] = x0.x.y; sout[ob][6][tid] = x0.x.z;
        sout[ob][7][tid] = v0.ang.x; sout[ob][8][tid] = v0.ang.y; sout[ob][9][tid] = v0.ang.z;
        sout[ob][10][tid] = v0.lin.x; sout[ob][11][tid] = v0.lin.y; sout[ob][12][tid] = v0.lin.z;
        if (live && P.write_fa) {
            store_motion(P.acc, P.ld, b, a_last);
            store_motion(P.frc, P.ld, b, f_last);
        }
        fence_async_smem();
        __syncthreads();
        if (tid < kPipeOut) {
            double *dst = (tid < 7 ? P.pos + (uint64_t)tid * P.ld : P.vel + (uint64_t)(tid - 7) * P.ld) + tile * kPipeTB;
            bulk_s2g(dst, &sout[ob][tid][0], kPlaneBytes);
            bulk_commit();
        }
    }
    if (!DIRECT_OUT && tid < kPipeOut) bulk_wait_all();
}

// ================================================================== edge_fold gravity

// stage position of a body for slot 0/1/2 (f = 0, .5, 1): x (+) (dt*f)*v, linear part
template <bool EXACT>
__device__ __forceinline__ Vec3 stage_pos(const Vec3 &x, const Vec3 &v, double dtf)
{
    if (EXACT) return Vec3{ex::add(x.x, ex::mul(dtf, v.x)), ex::add(x.y, ex::mul(dtf, v.y)), ex::add(x.z, ex::mul(dtf, v.z))};
    return Vec3{fma(dtf, v.x, x.x), fma(dtf, v.y, x.y), fma(dtf, v.z, x.z)};
}

// Dense all-pairs (every body's out-edges are all other bodies, ascending): block =
// kBlockG source bodies of one world, targets streamed through shared memory in
// tiles; each thread folds its targets sequentially in ascending order, which is
// the reference's fold order (graph.rs:177-236) — so EXACT stays bit-exact.
static constexpr int kBlockG = 64;

// blockDim = (kBlockG, NS): thread (x, y) folds source x over all targets for stage slot y,
// so the three stage positions of a tick proceed in parallel while every fold stays sequential.
template <bool EXACT, bool RK4>
__global__ void __launch_bounds__(kBlockG * 3) graph_dense_kernel(const __grid_constant__ GraphParams G)
{
    constexpr int NS = RK4 ? 3 : 1;
    __shared__ double sx[NS][3][kBlockG];
    __shared__ double sm[kBlockG];

    const uint32_t N = G.n_entities;
    const uint32_t tiles = (N + kBlockG - 1) / kBlockG;
    const uint32_t world = blockIdx.x / tiles;
    const uint32_t tile = blockIdx.x % tiles;
    const uint32_t tx = threadIdx.x, sl = threadIdx.y; // sl = stage slot
    const uint32_t i = tile * kBlockG + tx;
    const uint64_t wbase = (uint64_t)world * N;
    const bool active = i < N;
    const bool newton = G.kind == B200_EFF_GRAVITY_EDGES_NEWTON;

    const double fac = sl == 0 ? 0.0 : (sl == 1 ? 0.5 : 1.0);
    const double dtf = EXACT ? ex::mul(G.dt_stage, fac) : fac * G.dt_stage;

    Vec3 xi = {0, 0, 0}, acc = {0, 0, 0};
    double mi = 0.0;
    if (active) {
        const uint64_t b = wbase + i;
        const Vec3 x = {ldp(G.pos, G.ld, 4, b), ldp(G.pos, G.ld, 5, b), ldp(G.pos, G.ld, 6, b)};
        const Vec3 v = {ldp(G.vel, G.ld, 3, b), ldp(G.vel, G.ld, 4, b), ldp(G.vel, G.ld, 5, b)};
        mi = ldp(G.ine, G.ld, 6, b);
        xi = RK4 ? stage_pos<EXACT>(x, v, dtf) : x;
    }

    for (uint32_t j0 = 0; j0 < N; j0 += kBlockG) {
        const uint32_t j = j0 + tx;
        __syncthreads();
        if (j < N) {
            const uint64_t b = wbase + j;
            const Vec3 x = {ldp(G.pos, G.ld, 4, b), ldp(G.pos, G.ld, 5, b), ldp(G.pos, G.ld, 6, b)};
            const Vec3 v = {ldp(G.vel, G.ld, 3, b), ldp(G.vel, G.ld, 4, b), ldp(G.vel, G.ld, 5, b)};
            const Vec3 p = RK4 ? stage_pos<EXACT>(x, v, dtf) : x;
            sx[sl][0][tx] = p.x; sx[sl][1][tx] = p.y; sx[sl][2][tx] = p.z;
            if (sl == 0) sm[tx] = ldp(G.ine, G.ld, 6, b);
        }
        __syncthreads();
        const uint32_t jn = min((uint32_t)kBlockG, N - j0);
        if (active) {
            for (uint32_t jj = 0; jj < jn; ++jj) {
                if (j0 + jj == i) continue;
                const double mj = sm[jj];
                const Vec3 xj = {sx[sl][0][jj], sx[sl][1][jj], sx[sl][2][jj]};
                if (EXACT) {
                    if (newton) ex::fold_newton(G.p0, xi, mi, xj, mj, acc);
                    else ex::fold_softened(G.p0, G.p1, xi, mi, xj, mj, acc);
                } else {
                    // common factor (G|K^2)*m_i applied after the loop
                    const Vec3 r = {xj.x - xi.x, xj.y - xi.y, xj.z - xi.z};
                    const double d2 = r.x * r.x + r.y * r.y + r.z * r.z + (newton ? 0.0 : G.p1);
                    const double inv = fa::rsqrt_nr(d2);
                    const double w = mj * inv * inv * inv;
                    acc.x = fma(w, r.x, acc.x); acc.y = fma(w, r.y, acc.y); acc.z = fma(w, r.z, acc.z);
                }
            }
        }
    }
    if (active) {
        const uint64_t b = wbase + i;
        const double k = EXACT ? 1.0 : G.p0 * mi;
        stp(G.gforce, G.ld, sl * 3 + 0, b, EXACT ? acc.x : k * acc.x);
        stp(G.gforce, G.ld, sl * 3 + 1, b, EXACT ? acc.y : k * acc.y);
        stp(G.gforce, G.ld, sl * 3 + 2, b, EXACT ? acc.z : k * acc.z);
    }
}

// FAST all-pairs: one warp per (source body, stage slot); lanes stride over the targets of a
// tile and keep private partial sums, a fixed xor-butterfly of warp shuffles combines them
// (summation order differs from the reference's sequential fold -> tolerance, not bit
// parity; EXACT uses graph_dense_kernel).  blockDim = (32, kFastSrc, NS): the 8 sources x 3
// slots of a CTA share one shared-memory tile of stage positions, and N = 1024, M = 1 still
// spreads over 128 CTAs x 24 warps.
static constexpr int kFastSrc = 8;

// SPLIT = true : blockDim (32, kFastSrc, NS), one warp per (source, slot)  — small batches
// SPLIT = false: blockDim (32, kFastSrc, 1),  one warp per source, NS slots — large batches
// TJ = targets per shared-memory tile: 256 (3 CTAs/SM) for big grids; 1024 for small grids, where
// the whole world of an N <= 1024 system is staged in ONE load phase instead of four dependent ones.
template <bool RK4, bool SPLIT, int TJ>
__global__ void __launch_bounds__(32 * kFastSrc * ((RK4 && SPLIT) ? 3 : 1)) graph_dense_fast_kernel(const __grid_constant__ GraphParams G)
{
    constexpr int NS = RK4 ? 3 : 1;       // stage slots of a tick
    constexpr int NW = SPLIT ? 1 : NS;    // slots folded by one warp
    constexpr int NT = 32 * kFastSrc * (SPLIT ? NS : 1);
    constexpr int kFastTJ = TJ;
    extern __shared__ double dsm[];
    double(*sx)[3][TJ] = reinterpret_cast<double(*)[3][TJ]>(dsm);
    double *sm = dsm + NS * 3 * TJ;

    const uint32_t N = G.n_entities;
    const uint32_t groups = (N + kFastSrc - 1) / kFastSrc;
    const uint32_t world = blockIdx.x / groups;
    const uint32_t grp = blockIdx.x % groups;
    const uint32_t lane = threadIdx.x, src = threadIdx.y, sl0 = SPLIT ? threadIdx.z : 0;
    const uint32_t flat = (threadIdx.z * kFastSrc + src) * 32 + lane;
    const uint32_t i = grp * kFastSrc + src;
    const uint64_t wbase = (uint64_t)world * N;
    const bool active = i < N;
    const bool newton = G.kind == B200_EFF_GRAVITY_EDGES_NEWTON;
    const double soft = newton ? 0.0 : G.p1;
    auto dtf_of = [&](uint32_t sl) { return sl == 0 ? 0.0 : (sl == 1 ? 0.5 * G.dt_stage : G.dt_stage); };

    Vec3 xi[NW], acc[NW];
    double mi = 0.0;
#pragma unroll
    for (int s = 0; s < NW; ++s) { xi[s] = Vec3{0, 0, 0}; acc[s] = Vec3{0, 0, 0}; }
    if (active) {
        const uint64_t b = wbase + i;
        const Vec3 x = {ldp(G.pos, G.ld, 4, b), ldp(G.pos, G.ld, 5, b), ldp(G.pos, G.ld, 6, b)};
        const Vec3 v = {ldp(G.vel, G.ld, 3, b), ldp(G.vel, G.ld, 4, b), ldp(G.vel, G.ld, 5, b)};
        mi = ldp(G.ine, G.ld, 6, b);
#pragma unroll
        for (int s = 0; s < NW; ++s) xi[s] = RK4 ? stage_pos<false>(x, v, dtf_of(sl0 + s)) : x;
    }
    for (uint32_t j0 = 0; j0 < N; j0 += kFastTJ) {
        __syncthreads();
        // NT threads fill the (kFastTJ x NS) tile: element t -> target t % TJ, slot t / TJ
        for (uint32_t t = flat; t < kFastTJ * NS; t += NT) {
            const uint32_t jt = t % kFastTJ, st = t / kFastTJ;
            const uint32_t j = j0 + jt;
            if (j < N) {
                const uint64_t b = wbase + j;
                const Vec3 x = {ldp(G.pos, G.ld, 4, b), ldp(G.pos, G.ld, 5, b), ldp(G.pos, G.ld, 6, b)};
                Vec3 pnt = x;
                if (RK4) {
                    const Vec3 v = {ldp(G.vel, G.ld, 3, b), ldp(G.vel, G.ld, 4, b), ldp(G.vel, G.ld, 5, b)};
                    pnt = stage_pos<false>(x, v, dtf_of(st));
                }
                sx[st][0][jt] = pnt.x; sx[st][1][jt] = pnt.y; sx[st][2][jt] = pnt.z;
                if (st == 0) sm[jt] = ldp(G.ine, G.ld, 6, b);
            }
        }
        __syncthreads();
        const uint32_t jn = min((uint32_t)kFastTJ, N - j0);
        if (active) {
#pragma unroll 4
            for (uint32_t jj = lane; jj < jn; jj += 32) {
                // the self pair contributes exactly nothing (and would be 0 * inf for Newton)
                const double mj = (j0 + jj == i) ? 0.0 : sm[jj];
                const bool self = j0 + jj == i;
#pragma unroll
                for (int s = 0; s < NW; ++s) {
                    const Vec3 r = {sx[sl0 + s][0][jj] - xi[s].x, sx[sl0 + s][1][jj] - xi[s].y, sx[sl0 + s][2][jj] - xi[s].z};
                    const double d2 = fma(r.x, r.x, fma(r.y, r.y, fma(r.z, r.z, soft)));
                    const double inv = fa::rsqrt_nr(d2);
                    const double w = self ? 0.0 : mj * inv * inv * inv;
                    acc[s].x = fma(w, r.x, acc[s].x); acc[s].y = fma(w, r.y, acc[s].y); acc[s].z = fma(w, r.z, acc[s].z);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NW; ++s) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc[s].x += __shfl_xor_sync(0xffffffffu, acc[s].x, off);
            acc[s].y += __shfl_xor_sync(0xffffffffu, acc[s].y, off);
            acc[s].z += __shfl_xor_sync(0xffffffffu, acc[s].z, off);
        }
    }
    if (active && lane == 0) {
        const uint64_t b = wbase + i;
        const double k = G.p0 * mi;
#pragma unroll
        for (int s = 0; s < NW; ++s) {
            stp(G.gforce, G.ld, (sl0 + s) * 3 + 0, b, k * acc[s].x);
            stp(G.gforce, G.ld, (sl0 + s) * 3 + 1, b, k * acc[s].y);
            stp(G.gforce, G.ld, (sl0 + s) * 3 + 2, b, k * acc[s].z);
        }
    }
}

// Small grids (one or a few worlds): gravity AND the body tick in ONE launch.  The CTA computes the
// three stage-slot forces of its 8 sources exactly like graph_dense_fast_kernel<RK4, SPLIT>, then 8 of
// its threads integrate those sources.  Other CTAs are still reading this tick's positions, so the
// new pose / velocity go to a second set of planes (ping-pong, swapped by the host after the launch).
// Saves the dependent second launch (~7 us of pure latency per tick at N = 1024, M = 1).
template <int TJ>
__global__ void __launch_bounds__(32 * kFastSrc * 3) nbody_tick_fused_kernel(const __grid_constant__ GraphParams G,
                                                                              const __grid_constant__ StepParams P,
                                                                              double *__restrict__ pos_out,
                                                                              double *__restrict__ vel_out)
{
    constexpr int NS = 3;
    constexpr int NT = 32 * kFastSrc * NS;
    extern __shared__ double dsm[];
    double(*sx)[3][TJ] = reinterpret_cast<double(*)[3][TJ]>(dsm);
    double *sm = dsm + NS * 3 * TJ;

    const uint32_t N = G.n_entities;
    const uint32_t groups = (N + kFastSrc - 1) / kFastSrc;
    const uint32_t world = blockIdx.x / groups;
    const uint32_t grp = blockIdx.x % groups;
    const uint32_t lane = threadIdx.x, src = threadIdx.y, sl = threadIdx.z;
    const uint32_t flat = (sl * kFastSrc + src) * 32 + lane;
    const uint32_t i = grp * kFastSrc + src;
    const uint64_t wbase = (uint64_t)world * N;
    const bool active = i < N;
    const bool newton = G.kind == B200_EFF_GRAVITY_EDGES_NEWTON;
    const double soft = newton ? 0.0 : G.p1;
    auto dtf_of = [&](uint32_t k) { return k == 0 ? 0.0 : (k == 1 ? 0.5 * G.dt_stage : G.dt_stage); };

    Vec3 xi = {0, 0, 0}, acc = {0, 0, 0};
    double mi = 0.0;
    if (active) {
        const uint64_t b = wbase + i;
        const Vec3 x = {ldp(G.pos, G.ld, 4, b), ldp(G.pos, G.ld, 5, b), ldp(G.pos, G.ld, 6, b)};
        const Vec3 v = {ldp(G.vel, G.ld, 3, b), ldp(G.vel, G.ld, 4, b), ldp(G.vel, G.ld, 5, b)};
        mi = ldp(G.ine, G.ld, 6, b);
        xi = stage_pos<false>(x, v, dtf_of(sl));
    }
    for (uint32_t j0 = 0; j0 < N; j0 += TJ) {
        __syncthreads();
        for (uint32_t t = flat; t < TJ * NS; t += NT) {
            const uint32_t jt = t % TJ, st = t / TJ;
            const uint32_t j = j0 + jt;
            if (j < N) {
                const uint64_t b = wbase + j;
                const Vec3 x = {ldp(G.pos, G.ld, 4, b), ldp(G.pos, G.ld, 5, b), ldp(G.pos, G.ld, 6, b)};
                const Vec3 v = {ldp(G.vel, G.ld, 3, b), ldp(G.vel, G.ld, 4, b), ldp(G.vel, G.ld, 5, b)};
                const Vec3 pnt = stage_pos<false>(x, v, dtf_of(st));
                sx[st][0][jt] = pnt.x; sx[st][1][jt] = pnt.y; sx[st][2][jt] = pnt.z;
                if (st == 0) sm[jt] = ldp(G.ine, G.ld, 6, b);
            }
        }
        __syncthreads();
        const uint32_t jn = min((uint32_t)TJ, N - j0);
        if (active) {
#pragma unroll 4
            for (uint32_t jj = lane; jj < jn; jj += 32) {
                const bool self = j0 + jj == i;
                const Vec3 r = {sx[sl][0][jj] - xi.x, sx[sl][1][jj] - xi.y, sx[sl][2][jj] - xi.z};
                const double d2 = fma(r.x, r.x, fma(r.y, r.y, fma(r.z, r.z, soft)));
                const double inv = fa::rsqrt_nr(d2);
                const double w = self ? 0.0 : sm[jj] * inv * inv * inv;
                acc.x = fma(w, r.x, acc.x); acc.y = fma(w, r.y, acc.y); acc.z = fma(w, r.z, acc.z);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    }
    if (active && lane == 0) {
        const uint64_t b = wbase + i;
        const double k = G.p0 * mi;
        stp(G.gforce, G.ld, sl * 3 + 0, b, k * acc.x);
        stp(G.gforce, G.ld, sl * 3 + 1, b, k * acc.y);
        stp(G.gforce, G.ld, sl * 3 + 2, b, k * acc.z);
    }
    __syncthreads(); // the CTA's gforce entries are visible to its integrating threads
    if (active && lane == 0 && sl == 0) {
        const uint64_t b = wbase + i;
        Pose x0 = load_pose(P.pos, P.ld, b);
        Motion v0 = load_motion(P.vel, P.ld, b);
        const Inertia I = load_inertia(P.ine, P.ld, b);
        Motion a_last, f_last;
        fast_ticks<B200_INTEGRATOR_RK4, true>(P, b, x0, v0, I, a_last, f_last, P.n_ticks, P.tick0, P.write_fa != 0, GravReg{});
        store_pose(pos_out, P.ld, b, x0);
        store_motion(vel_out, P.ld, b, v0);
        if (P.write_fa) {
            store_motion(P.acc, P.ld, b, a_last);
            store_motion(P.frc, P.ld, b, f_last);
        }
    }
}

// General edge list (CSR by source, spawn order inside a row): one thread per
// (world, source) gathers its targets.  Used for sparse / irregular graphs.
template <bool EXACT, bool RK4>
__global__ void __launch_bounds__(kBlockG) graph_csr_kernel(const __grid_constant__ GraphParams G)
{
    constexpr int NS = RK4 ? 3 : 1;
    const uint64_t t = (uint64_t)blockIdx.x * kBlockG + threadIdx.x;
    const uint64_t total = (uint64_t)G.n_entities * G.n_worlds;
    if (t >= total) return;
    const uint32_t i = (uint32_t)(t % G.n_entities);
    const uint64_t wbase = t - i;
    const bool newton = G.kind == B200_EFF_GRAVITY_EDGES_NEWTON;
    double dtf[3];
    dtf[0] = EXACT ? ex::mul(G.dt_stage, 0.0) : 0.0;
    dtf[1] = EXACT ? ex::mul(G.dt_stage, 0.5) : 0.5 * G.dt_stage;
    dtf[2] = EXACT ? ex::mul(G.dt_stage, 1.0) : G.dt_stage;

    const Vec3 x = {ldp(G.pos, G.ld, 4, t), ldp(G.pos, G.ld, 5, t), ldp(G.pos, G.ld, 6, t)};
    const Vec3 v = {ldp(G.vel, G.ld, 3, t), ldp(G.vel, G.ld, 4, t), ldp(G.vel, G.ld, 5, t)};
    const double mi = ldp(G.ine, G.ld, 6, t);
    Vec3 xi[NS], acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { xi[s] = RK4 ? stage_pos<EXACT>(x, v, dtf[s]) : x; acc[s] = Vec3{0, 0, 0}; }

    for (uint32_t e = G.row_ptr[i]; e < G.row_ptr[i + 1]; ++e) {
        const uint64_t bj = wbase + G.col_idx[e];
        const Vec3 xj0 = {ldp(G.pos, G.ld, 4, bj), ldp(G.pos, G.ld, 5, bj), ldp(G.pos, G.ld, 6, bj)};
        const Vec3 vj = {ldp(G.vel, G.ld, 3, bj), ldp(G.vel, G.ld, 4, bj), ldp(G.vel, G.ld, 5, bj)};
        const double mj = ldp(G.ine, G.ld, 6, bj);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const Vec3 xj = RK4 ? stage_pos<EXACT>(xj0, vj, dtf[s]) : xj0;
            if (EXACT) {
                if (newton) ex::fold_newton(G.p0, xi[s], mi, xj, mj, acc[s]);
                else ex::fold_softened(G.p0, G.p1, xi[s], mi, xj, mj, acc[s]);
            } else {
                const Vec3 r = {xj.x - xi[s].x, xj.y - xi[s].y, xj.z - xi[s].z};
                const double d2 = r.x * r.x + r.y * r.y + r.z * r.z + (newton ? 0.0 : G.p1);
                const double inv = fa::rsqrt_nr(d2);
                const double w = mj * inv * inv * inv;
                acc[s].x = fma(w, r.x, acc[s].x); acc[s].y = fma(w, r.y, acc[s].y); acc[s].z = fma(w, r.z, acc[s].z);
            }
        }
    }
    const double k = EXACT ? 1.0 : G.p0 * mi;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        stp(G.gforce, G.ld, s * 3 + 0, t, EXACT ? acc[s].x : k * acc[s].x);
        stp(G.gforce, G.ld, s * 3 + 1, t, EXACT ? acc[s].y : k * acc[s].y);
        stp(G.gforce, G.ld, s * 3 + 2, t, EXACT ? acc[s].z : k * acc[s].z);
    }
}

// ================================================================== small graph worlds: whole ticks in one warp
//
// A world of N <= 32 bodies fits in a warp: lane = body, floor(32/N) whole worlds per warp.  The edge_fold
// gravity of a tick needs the other bodies' three stage positions — functions of (x0, v0) only — which the
// lanes exchange with warp shuffles, so the state never leaves registers between ticks: one launch integrates
// n_ticks ticks (the generic route is two launches and a round trip of the 9 gravity planes through HBM per
// tick).  Every lane folds its out-edges sequentially in CSR (= spawn) order with the same ex:: functions as
// graph_dense_kernel / graph_csr_kernel, then runs the same tick function as body_exact_kernel — EXACT stays
// bit-identical to the oracle.  FAST folds sequentially too (no tree), with the FAST kernels' arithmetic.
template <bool EXACT, int INTEG, int MINB>
__global__ void __launch_bounds__(128, MINB) small_world_kernel(const __grid_constant__ GraphParams G,
                                                          const __grid_constant__ StepParams P)
{
    constexpr bool RK4 = INTEG == B200_INTEGRATOR_RK4;
    constexpr int NS = RK4 ? 3 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t N = G.n_entities;
    const uint32_t wpw = 32u / N; // worlds per warp
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t wl = lane / N, ent = lane - wl * N;
    const uint64_t world = warp * wpw + wl;
    const bool live = wl < wpw && world < G.n_worlds;
    const uint64_t b = live ? world * N + ent : 0;
    const uint32_t lane0 = lane - ent; // first lane of this lane's world
    const bool newton = G.kind == B200_EFF_GRAVITY_EDGES_NEWTON;

    Pose x0 = {{0.0, 0.0, 0.0, 1.0}, {0.0, 0.0, 0.0}};
    Motion v0 = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}, a_out = v0, f_out = v0;
    Inertia I = {{1.0, 1.0, 1.0}, 1.0};
    uint32_t e0 = 0, deg = 0;
    if (live) {
        x0 = load_pose(P.pos, P.ld, b);
        v0 = load_motion(P.vel, P.ld, b);
        if (EXACT) a_out = load_motion(P.acc, P.ld, b);
        I = load_inertia(P.ine, P.ld, b);
        e0 = G.row_ptr[ent];
        deg = G.row_ptr[ent + 1] - e0;
    }
    GravReg g;
    g.g0 = g.g1 = g.g2 = Vec3{0.0, 0.0, 0.0};
    g.has = deg != 0;

    for (uint32_t t = 0; t < P.n_ticks; ++t) {
        // stage positions of this body and the running folds, one per distinct stage position
        Vec3 p[NS], acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double fac = s == 0 ? 0.0 : (s == 1 ? 0.5 : 1.0);
            const double dtf = EXACT ? ex::mul(G.dt_stage, fac) : fac * G.dt_stage;
            p[s] = RK4 ? stage_pos<EXACT>(x0.x, v0.lin, dtf) : x0.x;
            acc[s] = Vec3{0.0, 0.0, 0.0};
        }
        for (uint32_t k = 0; k < G.max_deg; ++k) { // warp-uniform trip count: every lane takes part in the shuffles
            const bool on = k < deg;
            const uint32_t src = lane0 + (on ? G.col_idx[e0 + k] : ent);
            const double mj = __shfl_sync(FULL, I.m, src);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const Vec3 xj = {__shfl_sync(FULL, p[s].x, src), __shfl_sync(FULL, p[s].y, src), __shfl_sync(FULL, p[s].z, src)};
                if (!on) continue;
                if (EXACT) {
                    if (newton) ex::fold_newton(G.p0, p[s], I.m, xj, mj, acc[s]);
                    else ex::fold_softened(G.p0, G.p1, p[s], I.m, xj, mj, acc[s]);
                } else {
                    const Vec3 r = {xj.x - p[s].x, xj.y - p[s].y, xj.z - p[s].z};
                    const double d2 = fma(r.x, r.x, fma(r.y, r.y, fma(r.z, r.z, newton ? 0.0 : G.p1)));
                    const double inv = fa::rsqrt_nr(d2);
                    const double w = mj * inv * inv * inv;
                    acc[s].x = fma(w, r.x, acc[s].x); acc[s].y = fma(w, r.y, acc[s].y); acc[s].z = fma(w, r.z, acc[s].z);
                }
            }
        }
        const double kf = EXACT ? 1.0 : G.p0 * I.m; // FAST: common factor (G | K^2) * m_i applied once
        g.g0 = EXACT ? acc[0] : Vec3{kf * acc[0].x, kf * acc[0].y, kf * acc[0].z};
        if (RK4) {
            g.g1 = EXACT ? acc[NS - 2] : Vec3{kf * acc[NS - 2].x, kf * acc[NS - 2].y, kf * acc[NS - 2].z};
            g.g2 = EXACT ? acc[NS - 1] : Vec3{kf * acc[NS - 1].x, kf * acc[NS - 1].y, kf * acc[NS - 1].z};
        }
        if (!live) continue;
        if (EXACT) {
            exact_tick<INTEG, true>(P, b, x0, v0, a_out, f_out, I, g);
            uint64_t slot;
            if (traj_due(P, P.tick0 + t + 1, slot)) {
                traj_store_state(P, b, slot, x0, v0);
                if (P.traj_planes == 25) traj_store_af(P, b, slot, a_out, f_out);
            }
        } else {
            fast_ticks<INTEG, true, true>(P, b, x0, v0, I, a_out, f_out, 1u, P.tick0 + t, P.write_fa && t + 1 == P.n_ticks, g);
        }
    }
    if (!live) return;
    store_pose(P.pos, P.ld, b, x0);
    store_motion(P.vel, P.ld, b, v0);
    if (P.write_fa) {
        store_motion(P.acc, P.ld, b, a_out);
        store_motion(P.frc, P.ld, b, f_out);
    }
}

// ================================================================== layout kernels (K6)

static constexpr int kTile = 256;

__global__ void __launch_bounds__(kTile) aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa,
                                                           uint64_t n_bodies, uint32_t width, uint64_t ld)
{
    extern __shared__ double tile[]; // kTile * (width | 1)
    const uint32_t pitch = width | 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    const uint32_t nb = (uint32_t)min((uint64_t)kTile, n_bodies - base);
    const double *src = aos + base * width;
    for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) tile[(i / width) * pitch + (i % width)] = src[i];
    __syncthreads();
    if (threadIdx.x < nb)
        for (uint32_t k = 0; k < width; ++k) soa[(uint64_t)k * ld + base + threadIdx.x] = tile[threadIdx.x * pitch + k];
}

// gridDim.y = samples; a sample's planes start at soa + y*width*ld, its rows at aos + y*n_bodies*width
__global__ void __launch_bounds__(kTile) soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos,
                                                           uint64_t n_bodies, uint32_t width, uint64_t ld)
{
    extern __shared__ double tile[];
    const uint32_t pitch = width | 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    const uint32_t nb = (uint32_t)min((uint64_t)kTile, n_bodies - base);
    const double *s = soa + (uint64_t)blockIdx.y * width * ld;
    double *dst = aos + (uint64_t)blockIdx.y * n_bodies * width + base * width;
    if (threadIdx.x < nb)
        for (uint32_t k = 0; k < width; ++k) tile[threadIdx.x * pitch + k] = s[(uint64_t)k * ld + base + threadIdx.x];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) dst[i] = tile[(i / width) * pitch + (i % width)];
}

// All columns of a small batch in one launch: blockIdx.y selects the column, the AoS side of every
// column lives in one packed staging buffer (one PCIe transfer per direction per invoke_batch).
__global__ void __launch_bounds__(kTile) multi_transpose_kernel(const __grid_constant__ MultiColumns mc, uint64_t n_bodies, uint64_t ld,
                                                                int to_soa)
{
    extern __shared__ double tile[];
    const MultiColumns::Col c = mc.col[blockIdx.y];
    const uint32_t width = c.width, pitch = width | 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    if (base >= n_bodies) return;
    const uint32_t nb = (uint32_t)min((uint64_t)kTile, n_bodies - base);
    double *aos = mc.packed + c.aos_offset + base * width;
    if (to_soa) {
        for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) tile[(i / width) * pitch + (i % width)] = aos[i];
        __syncthreads();
        if (threadIdx.x < nb)
            for (uint32_t k = 0; k < width; ++k) c.soa[(uint64_t)k * ld + base + threadIdx.x] = tile[threadIdx.x * pitch + k];
    } else {
        if (threadIdx.x < nb)
            for (uint32_t k = 0; k < width; ++k) tile[threadIdx.x * pitch + k] = c.soa[(uint64_t)k * ld + base + threadIdx.x];
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) aos[i] = tile[(i / width) * pitch + (i % width)];
    }
}

// FP64 FMA throughput probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) probe_fp64_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(uint64_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ================================================================== launchers

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): remember the size each pair has been
// raised to (a process may hold handles on several GPUs; several kernels share a function-pointer type;
// one kernel may be launched with several tile sizes).
template <typename K>
static cudaError_t ensure_dynamic_smem(K kernel, size_t bytes)
{
    struct Entry { const void *kernel; int device; size_t bytes; };
    static std::mutex mu;
    static std::vector<Entry> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const void *key = reinterpret_cast<const void *>(kernel);
    std::lock_guard<std::mutex> lock(mu);
    Entry *hit = nullptr;
    for (auto &d : done) if (d.kernel == key && d.device == dev) hit = &d;
    if (hit && hit->bytes >= bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (hit) hit->bytes = bytes; else done.push_back(Entry{key, dev, bytes});
    return cudaSuccess;
}

// the instantiation without trajectory code for launches that record nothing, the generic one otherwise
#define BODY_FAST(INTEG, BLOCK, MINB)                                                        \
    do {                                                                                     \
        if (P.traj_every) body_fast_kernel<INTEG, BLOCK, MINB, true><<<g(BLOCK), BLOCK, 0, s>>>(P);  \
        else body_fast_kernel<INTEG, BLOCK, MINB, false><<<g(BLOCK), BLOCK, 0, s>>>(P);      \
    } while (0)

cudaError_t launch_body_step(const StepParams &P, int integrator, int math_mode, cudaStream_t s)
{
    if (P.n_bodies == 0) return cudaSuccess;
    const bool rk4 = integrator == B200_INTEGRATOR_RK4;
    if (math_mode == B200_MATH_EXACT) {
        static const int xcfg = [] { const char *e = getenv("B200_EXACT_CFG"); return e ? atoi(e) : 3; }();
        auto g = [&](int blk) { return (unsigned)((P.n_bodies + blk - 1) / blk); };
        if (!rk4) body_exact_kernel<B200_INTEGRATOR_SEMI_IMPLICIT, 256, 1><<<g(256), 256, 0, s>>>(P);
        else switch (xcfg) {
        case 1: body_exact_kernel<B200_INTEGRATOR_RK4, 256, 2><<<g(256), 256, 0, s>>>(P); break;
        case 2: body_exact_kernel<B200_INTEGRATOR_RK4, 128, 3><<<g(128), 128, 0, s>>>(P); break;
        case 0: body_exact_kernel<B200_INTEGRATOR_RK4, 256, 1><<<g(256), 256, 0, s>>>(P); break;
        default: body_exact_kernel<B200_INTEGRATOR_RK4, 128, 4><<<g(128), 128, 0, s>>>(P); break; // 4.3e9 vs 2.7e9 (256x1)
        }
    } else {
        static const int cfg = [] { const char *e = getenv("B200_BODY_CFG"); return e ? atoi(e) : 3; }();
        auto g = [&](int blk) { return (unsigned)((P.n_bodies + blk - 1) / blk); };
        if (!rk4) BODY_FAST(B200_INTEGRATOR_SEMI_IMPLICIT, 128, 4);
        else switch (cfg) {
        case 1: BODY_FAST(B200_INTEGRATOR_RK4, 256, 2); break;
        case 2: BODY_FAST(B200_INTEGRATOR_RK4, 128, 3); break;
        case 0: BODY_FAST(B200_INTEGRATOR_RK4, 256, 1); break;
        case 4: BODY_FAST(B200_INTEGRATOR_RK4, 64, 8); break;
        case 5: BODY_FAST(B200_INTEGRATOR_RK4, 128, 5); break;
        case 10: case 11: case 12: case 13: case 14: {
            // persistent TMA-pipelined kernel; needs plane stride % 128 == 0 (whole tiles inside a plane)
            const bool direct = cfg >= 13;
            const int stages = cfg == 10 ? 2 : (cfg == 11 ? 3 : (cfg == 12 ? 4 : (cfg == 13 ? 2 : 3)));
            const size_t smem = sizeof(double) * (stages * kPipeIn + (direct ? 0 : 2) * kPipeOut) * kPipeTB + 8 * stages;
            auto kern = cfg == 10 ? body_fast_pipe_kernel<B200_INTEGRATOR_RK4, 2, 3, false>
                      : cfg == 11 ? body_fast_pipe_kernel<B200_INTEGRATOR_RK4, 3, 2, false>
                      : cfg == 12 ? body_fast_pipe_kernel<B200_INTEGRATOR_RK4, 4, 2, false>
                      : cfg == 13 ? body_fast_pipe_kernel<B200_INTEGRATOR_RK4, 2, 4, true>
                                  : body_fast_pipe_kernel<B200_INTEGRATOR_RK4, 3, 4, true>;
            // opt-in tuning variant: attributes are (re)set on every launch, cheap next to a >100 us kernel
            int dev = 0, sm_count = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPipeTB, smem);
            if (per_sm < 1) per_sm = 1;
            const uint64_t n_tiles = (P.n_bodies + kPipeTB - 1) / kPipeTB;
            const unsigned grid_p = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sm_count * per_sm);
            kern<<<grid_p, kPipeTB, smem, s>>>(P);
            break;
        }
        // default (cfg 3): 128 threads x 4 CTAs/SM = 16 warps/SM at <= 128 registers — measured 95% of the
        // HBM copy peak on B200 vs 57% for 256x1 (profiles/r01_tuning.md)
        default: BODY_FAST(B200_INTEGRATOR_RK4, 128, 4); break;
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_graph_force(const GraphParams &G, int math_mode, bool dense, cudaStream_t s)
{
    const bool exact = math_mode == B200_MATH_EXACT;
    const bool rk4 = G.integrator == B200_INTEGRATOR_RK4;
    if (G.n_entities == 0 || G.n_worlds == 0) return cudaSuccess;
    if (dense) {
        const unsigned tiles = (G.n_entities + kBlockG - 1) / kBlockG;
        const unsigned grid = tiles * G.n_worlds;
        static const int gcfg = [] { const char *e = getenv("B200_GRAPH_CFG"); return e ? atoi(e) : 1; }();
        const dim3 blk3(kBlockG, 3), blk1(kBlockG, 1);
        if (exact) { if (rk4) graph_dense_kernel<true, true><<<grid, blk3, 0, s>>>(G); else graph_dense_kernel<true, false><<<grid, blk1, 0, s>>>(G); }
        else if (gcfg == 0) { if (rk4) graph_dense_kernel<false, true><<<grid, blk3, 0, s>>>(G); else graph_dense_kernel<false, false><<<grid, blk1, 0, s>>>(G); }
        else {
            const unsigned gridf = ((G.n_entities + kFastSrc - 1) / kFastSrc) * G.n_worlds;
            // few CTAs: split the stage slots over warps to fill the machine; many CTAs: keep
            // 3 slots per warp (more ILP per lane, 3 CTAs/SM) — measured on N = 1024, M = 1 / 8
            const bool split = gcfg == 2 || (gcfg == 1 && gridf < 3u * 148u);
            constexpr size_t smem256 = (3 * 3 + 1) * 256 * sizeof(double), smem1024 = (3 * 3 + 1) * 1024 * sizeof(double);
            if (!rk4) graph_dense_fast_kernel<false, false, 256><<<gridf, dim3(32, kFastSrc, 1), smem256, s>>>(G);
            else if (split) {
                const cudaError_t e = ensure_dynamic_smem(graph_dense_fast_kernel<true, true, 1024>, smem1024);
                if (e != cudaSuccess) return e;
                graph_dense_fast_kernel<true, true, 1024><<<gridf, dim3(32, kFastSrc, 3), smem1024, s>>>(G);
            } else graph_dense_fast_kernel<true, false, 256><<<gridf, dim3(32, kFastSrc, 1), smem256, s>>>(G);
        }
    } else {
        const uint64_t total = (uint64_t)G.n_entities * G.n_worlds;
        const unsigned grid = (unsigned)((total + kBlockG - 1) / kBlockG);
        if (exact) { if (rk4) graph_csr_kernel<true, true><<<grid, kBlockG, 0, s>>>(G); else graph_csr_kernel<true, false><<<grid, kBlockG, 0, s>>>(G); }
        else { if (rk4) graph_csr_kernel<false, true><<<grid, kBlockG, 0, s>>>(G); else graph_csr_kernel<false, false><<<grid, kBlockG, 0, s>>>(G); }
    }
    return cudaGetLastError();
}

bool nbody_fused_applicable(const GraphParams &G, int math_mode, bool dense)
{
    if (math_mode != B200_MATH_FAST || !dense || G.integrator != B200_INTEGRATOR_RK4) return false;
    static const int fcfg = [] { const char *e = getenv("B200_NBODY_FUSED"); return e ? atoi(e) : 1; }();
    const unsigned gridf = ((G.n_entities + kFastSrc - 1) / kFastSrc) * G.n_worlds;
    return fcfg != 0 && gridf < 3u * 148u; // the same "small grid" rule as the split gravity kernel
}

cudaError_t launch_nbody_tick_fused(const GraphParams &G, const StepParams &P, double *pos_out, double *vel_out, cudaStream_t s)
{
    constexpr size_t smem = (3 * 3 + 1) * 1024 * sizeof(double);
    const cudaError_t e = ensure_dynamic_smem(nbody_tick_fused_kernel<1024>, smem);
    if (e != cudaSuccess) return e;
    const unsigned gridf = ((G.n_entities + kFastSrc - 1) / kFastSrc) * G.n_worlds;
    nbody_tick_fused_kernel<1024><<<gridf, dim3(32, kFastSrc, 3), smem, s>>>(G, P, pos_out, vel_out);
    return cudaGetLastError();
}

bool small_world_applicable(const GraphParams &G, int math_mode)
{
    // measured (profiles/r01_small_world.md): faster than the two-launch route over the whole range a warp can
    // hold, in both arithmetic modes (N = 3: 53x FAST / 4.4x EXACT; N = 32: 3.6x / 1.35x)
    (void)math_mode;
    static const int cfg = [] { const char *e = getenv("B200_SMALL_WORLD"); return e ? atoi(e) : 1; }();
    return cfg != 0 && G.n_entities >= 1 && G.n_entities <= 32;
}

template <int MINB>
static void launch_small_world_cfg(const GraphParams &G, const StepParams &P, int math_mode, unsigned grid, cudaStream_t s)
{
    const bool rk4 = G.integrator == B200_INTEGRATOR_RK4;
    if (math_mode == B200_MATH_EXACT) {
        if (rk4) small_world_kernel<true, B200_INTEGRATOR_RK4, MINB><<<grid, 128, 0, s>>>(G, P);
        else small_world_kernel<true, B200_INTEGRATOR_SEMI_IMPLICIT, MINB><<<grid, 128, 0, s>>>(G, P);
    } else {
        if (rk4) small_world_kernel<false, B200_INTEGRATOR_RK4, MINB><<<grid, 128, 0, s>>>(G, P);
        else small_world_kernel<false, B200_INTEGRATOR_SEMI_IMPLICIT, MINB><<<grid, 128, 0, s>>>(G, P);
    }
}

cudaError_t launch_small_world(const GraphParams &G, const StepParams &P, int math_mode, cudaStream_t s)
{
    if (G.n_entities == 0 || G.n_worlds == 0) return cudaSuccess;
    const uint32_t wpw = 32u / G.n_entities;
    const uint64_t warps = ((uint64_t)G.n_worlds + wpw - 1) / wpw;
    const unsigned grid = (unsigned)((warps + 3) / 4); // 4 warps per CTA
    // resident CTAs per SM the register allocation is bounded for (B200_SMALL_WORLD_CFG = 2 | 3 | 4): the kernel
    // is latency-bound, 16 warps/SM at 128 registers (a few spilled doubles) beat 8 warps at 196 by 1.3-1.45x
    static const int cfg = [] { const char *e = getenv("B200_SMALL_WORLD_CFG"); return e ? atoi(e) : 4; }();
    switch (cfg) {
    case 2: launch_small_world_cfg<2>(G, P, math_mode, grid, s); break;
    case 3: launch_small_world_cfg<3>(G, P, math_mode, grid, s); break;
    default: launch_small_world_cfg<4>(G, P, math_mode, grid, s); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_aos_to_soa(const double *aos, double *soa, uint64_t n_bodies, uint32_t width, uint64_t ld, cudaStream_t s)
{
    if (n_bodies == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_bodies + kTile - 1) / kTile);
    aos_to_soa_kernel<<<grid, kTile, kTile * (width | 1u) * sizeof(double), s>>>(aos, soa, n_bodies, width, ld);
    return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(const double *soa, double *aos, uint64_t n_bodies, uint32_t width, uint64_t ld, cudaStream_t s)
{
    if (n_bodies == 0) return cudaSuccess;
    const dim3 grid((unsigned)((n_bodies + kTile - 1) / kTile), 1);
    soa_to_aos_kernel<<<grid, kTile, kTile * (width | 1u) * sizeof(double), s>>>(soa, aos, n_bodies, width, ld);
    return cudaGetLastError();
}

cudaError_t launch_traj_to_aos(const double *traj, double *aos, uint64_t n_samples, uint64_t n_bodies, uint64_t ld,
                               uint32_t width, cudaStream_t s)
{
    if (n_bodies == 0 || n_samples == 0) return cudaSuccess;
    const size_t smem = (size_t)kTile * (width | 1u) * sizeof(double); // 25 planes: 52 KB, above the 48 KB default
    cudaError_t e = ensure_dynamic_smem(soa_to_aos_kernel, smem);
    if (e != cudaSuccess) return e;
    for (uint64_t s0 = 0; s0 < n_samples; s0 += 32768) {
        const unsigned ny = (unsigned)min((uint64_t)32768, n_samples - s0);
        const dim3 grid((unsigned)((n_bodies + kTile - 1) / kTile), ny);
        soa_to_aos_kernel<<<grid, kTile, smem, s>>>(traj + s0 * width * ld, aos + s0 * n_bodies * width, n_bodies, width, ld);
    }
    return cudaGetLastError();
}

cudaError_t launch_multi_transpose(const MultiColumns &mc, uint64_t n_bodies, uint64_t ld, bool to_soa, cudaStream_t s)
{
    if (n_bodies == 0 || mc.n == 0) return cudaSuccess;
    uint32_t wmax = 1;
    for (uint32_t i = 0; i < mc.n; ++i) wmax = std::max(wmax, mc.col[i].width);
    const dim3 grid((unsigned)((n_bodies + kTile - 1) / kTile), mc.n);
    multi_transpose_kernel<<<grid, kTile, kTile * (wmax | 1u) * sizeof(double), s>>>(mc, n_bodies, ld, to_soa ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_probe_fp64(double *out, int iters, int blocks, cudaStream_t s)
{
    probe_fp64_kernel<<<blocks, 256, 0, s>>>(out, iters);
    return cudaGetLastError();
}

} // namespace b200
