/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, f64, no FMA contraction) of the arithmetic of
 * elodin-sys/elodin's six_dof() path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.  The
 * product (elodin_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED against the reference's own golden telemetry
 * (scripts/ci/baseline/{three-body,rocket,ball}-csv for RK4 and the `earth`
 * entity of cube-sat-csv for SemiImplicit, repacked under tests/golden/ by
 * tests/golden/make_golden.py) and its known-answer tests
 * (libs/nox/src/spatial.rs:630-676, libs/nox/src/quaternion.rs:352-388,
 * libs/nox-py/python/tests/test_all.py:67-83,228-291,342-366).
 * UNPINNED (no golden in the reference): softened n-body gravity at N>3 and the
 * falcon9 frame-force effector — oracle-vs-kernel only.
 *
 * The reference implementation itself (Rust + JAX + Cranelift) cannot be built
 * or imported in this image (no cargo/rustc/jax), so there is no oracle/_ref.
 */
#ifndef SIXDOF_ORACLE_H
#define SIXDOF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* effector kinds: same numbering as include/b200_sixdof.h */
enum {
    ORC_EFF_GRAVITY_CONST = 1,
    ORC_EFF_DRAG_QUADRATIC = 2,
    ORC_EFF_THRUST_BODY = 3,
    ORC_EFF_WRENCH_BODY = 4,
    ORC_EFF_GRAVITY_FRAME = 5,
    ORC_EFF_GRAVITY_EDGES_NEWTON = 6,
    ORC_EFF_GRAVITY_EDGES_SOFTENED = 7,
    ORC_EFF_WRENCH_WORLD = 8,
    ORC_EFF_TORQUE_BODY_FOLD = 9,
    ORC_EFF_GRAVITY_J2 = 10,
    ORC_EFF_GRAVITY_EGM08 = 11
};
#define ORC_FLAG_WRENCH_LINEAR_FIRST 1u

typedef struct orc_effector {
    uint32_t kind;
    uint32_t flags;
    double p[8];
    const double *column;  /* [n_worlds][n][width] AoS, may be NULL */
    uint32_t column_width;
    uint32_t reserved;
    uint64_t n_edges;
    const uint32_t *edge_from;
    const uint32_t *edge_to;
    const uint8_t *entity_mask; /* [n] or NULL: query-join membership (query.rs:672-710) */
    const double *table0;   /* GRAVITY_EGM08: normalised C coefficients, [(L+1)][(L+1)] row-major (row = degree l) */
    const double *table1;   /* GRAVITY_EGM08: normalised S coefficients, same shape */
    uint64_t table_len;     /* (L+1)^2 */
} orc_effector;

/* Derived tables of the EGM08 recursion (python/elodin/egm08.py:84-141): the same arithmetic in the oracle and in
 * libb200_sixdof's host code, so both sides hold bit-identical tables.  out = [n1 | n2 | nq1 | nq2] each (L+1)^2 row-major
 * [l][m], then diag[L+1], then offc[L+1] (sub-diagonal factor without u); out must hold 4 (L+1)^2 + 2 (L+1) doubles. */
void orc_egm08_tables(int L, double *out);

/* AoS columns [n_worlds][n][width], exactly the host layout of the C ABI */
typedef struct orc_world {
    uint64_t n;        /* bodies per world */
    uint64_t n_worlds;
    double *pos;       /* [..,7]  q(i,j,k,w), x,y,z */
    double *vel;       /* [..,6]  omega, v          */
    double *accel;     /* [..,6]                    */
    double *force;     /* [..,6]  tau, f            */
    const double *inertia; /* [..,7] diag(3), momentum(3), mass */
} orc_world;

/* dot evaluation mode, see sixdof_oracle.c: 0 = plain IEEE (canonical),
 * 1 = reproduce the FMA-contracted `dot` of the host JIT that generated the
 * reference's golden rocket telemetry (test pinning only) */
#define ORC_DOT_PLAIN 0
#define ORC_DOT_GOLDEN_HOST 1
void orc_set_dot_mode(int mode);
int orc_get_dot_mode(void);

/* primitives (exposed for the known-answer tests) */
void orc_qmul(const double l[4], const double r[4], double out[4]);
void orc_qinv(const double q[4], double out[4]);
void orc_qrot(const double q[4], const double v[3], double out[3]);
void orc_qnormalize(const double q[4], double out[4]);
void orc_transform_add_motion(const double pos[7], const double motion[6], double out[7]);
void orc_calc_accel(const double pos[7], const double force[6], const double inertia[7], double accel[6]);

/* one tick of six_dof(Rk4) / six_dof(SemiImplicit) over every world.
 * dt_stage = SimulationTimeStep, dt_final = six_dof(time_step=) or dt_stage.
 * n_threads > 1 parallelises over worlds with pthreads (bench baseline only). */
void orc_rk4_ticks(orc_world *w, uint32_t n_eff, const orc_effector *effs,
                   double dt_stage, double dt_final, uint64_t n_ticks, int n_threads);
void orc_semi_implicit_ticks(orc_world *w, uint32_t n_eff, const orc_effector *effs,
                             double dt, uint64_t n_ticks, int n_threads);

/* evaluate only the effector pipe + calc_accel on the given state (one "stage") */
void orc_eval_stage(const orc_world *w, uint64_t world, uint32_t n_eff, const orc_effector *effs,
                    const double *pos, const double *vel, double *force, double *accel);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
