// C++ view of the drop-in seam (include/b200_world.hpp): the reference's own checks, restated.
//   * test_six_dof            libs/nox-py/python/tests/test_all.py:67-83   x = dt after one tick of six_dof(1/60)
//   * three-body tick 1       scripts/ci/baseline/three-body-csv/a.world_pos.csv row 2, bit for bit
//   * ValueSizeMismatch / ComponentNotFound error mapping (error.rs:7-58)
//   * parse_backend_config (world_builder.rs:245-260) and the system_names -> built-in effector matcher (system.rs:213-222)
// Without a GPU the only legal outcome is a loud B200_ERR_NO_DEVICE.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "b200_world.hpp"

#define REQUIRE(c)                                                                 \
    do {                                                                           \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main()
{
    using namespace b200;
    REQUIRE(component_id("world_pos") == B200_ID_WORLD_POS);
    { // host-side logic that needs no device: backend strings and the effector matcher
        unsetenv("ELODIN_BACKEND");
        REQUIRE(parse_backend_config("b200").math_mode == B200_MATH_FAST);
        REQUIRE(parse_backend_config("  B200-Exact ").math_mode == B200_MATH_EXACT);
        REQUIRE(parse_backend_config("b200-fast").math_mode == B200_MATH_FAST);
        for (const char *other : {"cranelift", "jax-cpu", "jax-gpu", "cpu", ""}) {
            bool threw = false;
            try { parse_backend_config(other); } catch (const Error &e) { threw = std::string(e.what()).find("unknown backend") != std::string::npos; }
            REQUIRE(threw);
        }
        setenv("ELODIN_BACKEND", "b200-exact", 1); // the environment wins, as in the reference
        REQUIRE(parse_backend_config("b200").math_mode == B200_MATH_EXACT);
        unsetenv("ELODIN_BACKEND");
        // the rocket example's pipeline as CompiledSystem.system_names would list it
        const std::vector<std::string> rocket = {"<system>", "<function clear_forces at 0x7f00aa>", "<function gravity at 0x7f00bb>",
                                                 "<function apply_thrust at 0x7f00cc>", "<function apply_aero_forces at 0x7f00dd>",
                                                 "<function calc_accel at 0x7f00ee>"};
        const std::vector<b200_effector> effs = match_effectors(rocket);
        REQUIRE(effs.size() == 3);
        REQUIRE(effs[0].kind == B200_EFF_GRAVITY_CONST && effs[0].p[2] == -9.81 && effs[0].column_id == 0);
        REQUIRE(effs[1].kind == B200_EFF_THRUST_BODY && effs[1].column_id == component_id("thrust") && effs[1].column_width == 1);
        REQUIRE(effs[2].kind == B200_EFF_WRENCH_BODY && effs[2].column_id == component_id("aero_force") && effs[2].flags == 0);
        const std::vector<b200_effector> f9 = match_effectors({"<function gravity_and_frame_forces at 0x1>", "<function apply_body_wrenches at 0x2>"});
        REQUIRE(f9.size() == 2 && f9[0].kind == B200_EFF_GRAVITY_FRAME && f9[1].flags == B200_EFF_FLAG_WRENCH_LINEAR_FIRST);
        // an arbitrary user system is a hard error that names it; an @el.map wrapper says why it cannot be matched
        for (const char *bad : {"<function kalman_filter at 0x3>", "<function map.<locals>.inner at 0x4>"}) {
            bool ok = false;
            try { match_effectors({"<function gravity at 0x1>", bad}); }
            catch (const Error &e) { ok = e.code == B200_ERR_UNSUPPORTED && std::string(e.what()).find(system_function_name(bad)) != std::string::npos; }
            REQUIRE(ok);
        }
        // a host registers its own names
        auto reg = default_effector_registry();
        reg["wind_drag"] = make_spec(B200_EFF_DRAG_QUADRATIC, {0.6, 0.01}, "wind", 3);
        REQUIRE(match_effectors({"<function wind_drag at 0x5>"}, reg)[0].kind == B200_EFF_DRAG_QUADRATIC);
    }
    if (b200_device_count() <= 0) {
        World w;
        w.spawn(Body{});
        try {
            WorldExec ex(w, {});
            std::printf("FAILED: create succeeded without a GPU\n");
            return 1;
        } catch (const Error &e) {
            REQUIRE(e.code == B200_ERR_NO_DEVICE);
            std::printf("no GPU: failed loudly as designed (%s)\n", e.what());
            return 0;
        }
    }
    { // test_six_dof
        World w;
        Body b;
        b.world_vel[3] = 1.0;
        w.spawn(b);
        w.sim_time_step = 0.008333333;
        WorldExec ex(w, {}, B200_INTEGRATOR_RK4, B200_MATH_EXACT, 1.0 / 60.0);
        ex.run();
        const double *p = ex.world.row(B200_ID_WORLD_POS, 0);
        REQUIRE(p[3] == 1.0 && std::fabs(p[4] - 0.01666667) < 1e-8 && p[5] == 0.0);
        REQUIRE(ex.world.tick == 1);
    }
    { // three-body, first recorded tick of the reference's golden telemetry
        const double G = 6.6743e-11;
        World w;
        const double x[3] = {0.8920281421, -0.6628498947, -0.2291782474}, vy[3] = {0.9957939373, -1.6191613336, 0.6233673964};
        for (int i = 0; i < 3; ++i) {
            Body b;
            b.world_pos[4] = x[i];
            b.world_vel[4] = vy[i];
            for (int k : {0, 1, 2, 6}) b.inertia[k] = 1.0 / G;
            w.spawn(b);
        }
        w.sim_time_step = 0.008333333;
        w.ticks_per_telemetry = 1;
        const uint32_t from[6] = {0, 1, 0, 1, 2, 2}, to[6] = {1, 0, 2, 2, 0, 1}; // spawn order, main.py:82-89
        b200_effector e;
        std::memset(&e, 0, sizeof e);
        e.kind = B200_EFF_GRAVITY_EDGES_NEWTON;
        e.p[0] = G;
        e.n_edges = 6;
        e.edge_from = from;
        e.edge_to = to;
        WorldExec ex(w, {e});
        ex.run();
        const double *a = ex.world.row(B200_ID_WORLD_POS, 0);
        REQUIRE(a[4] == 0.8919861600553762 && a[5] == 0.00829818990519979 && a[6] == 0.0 && a[3] == 1.0);
        for (int t = 1; t < 100; ++t) ex.run();
        REQUIRE(ex.world.tick == 100);
        // a batch of 10 ticks per telemetry cycle lands on the same bits
        World w2 = w;
        w2.ticks_per_telemetry = 10;
        WorldExec ex2(w2, {e});
        for (int c = 0; c < 10; ++c) ex2.run();
        for (int k = 0; k < 7; ++k) REQUIRE(ex2.world.row(B200_ID_WORLD_POS, 1)[k] == ex.world.row(B200_ID_WORLD_POS, 1)[k]);
    }
    { // error mapping
        World w;
        w.spawn(Body{});
        WorldExec ex(w, {});
        ex.world.host[B200_ID_WORLD_VEL].buffer.resize(8);
        try { ex.run(); REQUIRE(!"expected ValueSizeMismatch"); } catch (const ValueSizeMismatch &) {}
        ex.world.host.erase(B200_ID_INERTIA);
        try { ex.run(); REQUIRE(!"expected ComponentNotFound"); } catch (const ComponentNotFound &) {}
    }
    std::printf("C++ host mirror ok\n");
    return 0;
}
