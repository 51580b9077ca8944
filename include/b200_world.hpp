// b200_world.hpp — header-only C++17 mirror of the nox-py executor seam, on top of the C ABI.
//
// The reference host is Rust; Rust is not in this image, so this is the compiled-language view of the
// same three types a maintainer would touch when adding a `WorldExec::B200` variant:
//
//   b200::Column / b200::World   <- `Column{buffer, entity_ids}` / `World{host: BTreeMap<ComponentId, Column>, ..}`
//                                   (libs/nox-py/src/world.rs:25-60,174-200)
//   b200::Exec                   <- `CraneliftExec` (libs/nox-py/src/cranelift_exec.rs:13-195): input/output id
//                                   tables, executor-owned output buffers, invoke_batch(world, n)
//   b200::WorldExec              <- `CraneliftWorldExec::run` (cranelift_exec.rs:284-303): one invoke_batch of
//                                   ticks_per_telemetry ticks, then world.advance_tick() x n
//
// Errors follow libs/nox-py/src/error.rs: ComponentNotFound / ValueSizeMismatch / backend(String).
// Nothing here computes: every tick runs in libb200_sixdof.so (include/b200_sixdof.h).
#pragma once

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200_sixdof.h"

namespace b200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
struct ComponentNotFound : Error { ComponentNotFound() : Error(B200_ERR_COMPONENT_NOT_FOUND, "component not found") {} };
struct ValueSizeMismatch : Error { ValueSizeMismatch() : Error(B200_ERR_VALUE_SIZE_MISMATCH, "component value had wrong size") {} };

inline void check(int rc)
{
    if (rc == B200_OK) return;
    if (rc == B200_ERR_COMPONENT_NOT_FOUND) throw ComponentNotFound();
    if (rc == B200_ERR_VALUE_SIZE_MISMATCH) throw ValueSizeMismatch();
    throw Error(rc, b200_last_error());
}

using ComponentId = uint64_t;
inline ComponentId component_id(const char *name) { return b200_component_id(name); }

// world.rs:25-29
struct Column {
    std::vector<uint8_t> buffer;      // [n_worlds][rows][width] f64 LE (u64 for tick)
    std::vector<uint64_t> entity_ids; // row i = i-th spawned entity owning the component
    uint32_t width = 0;
};

// The Body archetype (six_dof.rs:153-159): pos, vel, accel, force, inertia
struct Body {
    double world_pos[7] = {0, 0, 0, 1, 0, 0, 0};
    double world_vel[6] = {0, 0, 0, 0, 0, 0};
    double world_accel[6] = {0, 0, 0, 0, 0, 0};
    double force[6] = {0, 0, 0, 0, 0, 0};
    double inertia[7] = {1, 1, 1, 0, 0, 0, 1};
};

class World {
  public:
    std::map<ComponentId, Column> host; // BTreeMap order = ComponentId order
    uint64_t tick = 0;
    double sim_time_step = 1.0 / 120.0; // TimeStep (world.rs:31-39), already Duration-quantised by the caller
    uint64_t ticks_per_telemetry = 1;
    uint64_t entity_len = 1;            // entity 0 = Globals (world.rs:174-183)
    uint64_t n_worlds = 1;

    World()
    {
        put_scalar(B200_ID_TICK);
        put_scalar(B200_ID_SIMULATION_TIME_STEP);
    }

    // spawn a Body on every world of the batch; returns the EntityId
    uint64_t spawn(const Body &b)
    {
        const uint64_t id = entity_len++;
        append(B200_ID_WORLD_POS, id, b.world_pos, 7);
        append(B200_ID_WORLD_VEL, id, b.world_vel, 6);
        append(B200_ID_WORLD_ACCEL, id, b.world_accel, 6);
        append(B200_ID_FORCE, id, b.force, 6);
        append(B200_ID_INERTIA, id, b.inertia, 7);
        return id;
    }
    // attach an extra f64 component (effector input) to the most recently spawned entity
    void insert(ComponentId cid, const double *v, uint32_t width) { append(cid, entity_len - 1, v, width); }

    Column *column_by_id(ComponentId id)
    {
        auto it = host.find(id);
        return it == host.end() ? nullptr : &it->second;
    }
    uint64_t body_rows() const
    {
        auto it = host.find(B200_ID_WORLD_POS);
        return it == host.end() ? 0 : it->second.entity_ids.size();
    }
    void advance_tick() { ++tick; }
    void set_globals() // World::set_globals, world.rs:185-191
    {
        std::memcpy(host[B200_ID_SIMULATION_TIME_STEP].buffer.data(), &sim_time_step, 8);
        std::memcpy(host[B200_ID_TICK].buffer.data(), &tick, 8);
    }
    const double *row(ComponentId id, uint64_t r, uint64_t world = 0) const
    {
        const Column &c = host.at(id);
        return reinterpret_cast<const double *>(c.buffer.data()) + (world * c.entity_ids.size() + r) * c.width;
    }

  private:
    void put_scalar(ComponentId id)
    {
        Column c;
        c.buffer.assign(8, 0);
        c.entity_ids = {0};
        c.width = 1;
        host[id] = c;
    }
    void append(ComponentId cid, uint64_t entity, const double *v, uint32_t width)
    {
        if (n_worlds != 1) throw Error(B200_ERR_INVALID_ARGUMENT, "spawn before widening the world axis");
        Column &c = host[cid];
        if (c.width == 0) c.width = width;
        if (c.width != width) throw ValueSizeMismatch();
        const uint8_t *p = reinterpret_cast<const uint8_t *>(v);
        c.buffer.insert(c.buffer.end(), p, p + width * 8);
        c.entity_ids.push_back(entity);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Backend selection and effector matching: what the third `WorldExec` arm needs on the host side.
//
// parse_backend_config mirrors libs/nox-py/src/world_builder.rs:245-260: ELODIN_BACKEND overrides the requested
// string, the match is case-insensitive and trimmed; the B200 arm adds "b200" (= FAST math), "b200-fast" and
// "b200-exact" (bit-identical to the CPU arithmetic).  Every other name — including the reference's own
// "cranelift" / "jax-cpu" / "jax-gpu", which this library does not provide — is the reference's UnknownCommand error.
struct BackendConfig {
    uint32_t math_mode; // B200_MATH_*
    const char *device; // "gpu": there is no CPU arm behind this library
};

inline BackendConfig parse_backend_config(const std::string &requested_backend)
{
    const char *env = std::getenv("ELODIN_BACKEND");
    std::string s = env ? env : requested_backend;
    auto not_space = [](unsigned char c) { return !std::isspace(c); };
    s.erase(s.begin(), std::find_if(s.begin(), s.end(), not_space));
    s.erase(std::find_if(s.rbegin(), s.rend(), not_space).base(), s.end());
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (s == "b200" || s == "b200-fast") return BackendConfig{B200_MATH_FAST, "gpu"};
    if (s == "b200-exact") return BackendConfig{B200_MATH_EXACT, "gpu"};
    throw Error(B200_ERR_INVALID_ARGUMENT, "unknown backend '" + s + "': expected one of 'b200', 'b200-fast', 'b200-exact'");
}

// One entry of the effector registry: the Python-level system name -> the built-in effector it lowers to.
struct EffectorSpec {
    b200_effector effector; // kind / flags / constants / column width (column_id filled from `column`)
    std::string column;     // component name of the per-body input column, "" = none
};

inline EffectorSpec make_spec(uint32_t kind, std::initializer_list<double> p, const char *column, uint32_t width, uint32_t flags = 0)
{
    EffectorSpec s;
    std::memset(&s.effector, 0, sizeof s.effector);
    s.effector.kind = kind;
    s.effector.flags = flags;
    size_t i = 0;
    for (double v : p) if (i < 8) s.effector.p[i++] = v;
    s.column = column ? column : "";
    s.effector.column_width = width;
    return s;
}

// The effector systems of the reference's own examples, by function name (what `CompiledSystem.system_names`
// carries, libs/nox-py/src/system.rs:213-222,908: the repr of the Python function, "<function NAME at 0x..>").
// A host extends / overrides the table for its own simulation; names are matched exactly.
inline std::map<std::string, EffectorSpec> default_effector_registry()
{
    std::map<std::string, EffectorSpec> r;
    r["gravity"] = make_spec(B200_EFF_GRAVITY_CONST, {0.0, 0.0, -9.81}, nullptr, 0);               // ball/sim.py:56-58, rocket/main.py:292-294, drone/sim.py:106-108
    r["apply_drag"] = make_spec(B200_EFF_DRAG_QUADRATIC, {0.5 * 1.225, 2 * 3.1415 * 0.2 * 0.2}, "wind", 3); // ball/sim.py:99-116
    r["apply_thrust"] = make_spec(B200_EFF_THRUST_BODY, {-1.0, 0.0, 0.0}, "thrust", 1);             // rocket/main.py:429-431
    r["apply_aero_forces"] = make_spec(B200_EFF_WRENCH_BODY, {}, "aero_force", 6);                   // rocket/main.py:407-413
    r["apply_body_wrenches"] = make_spec(B200_EFF_WRENCH_BODY, {}, "body_wrench", 6, B200_EFF_FLAG_WRENCH_LINEAR_FIRST); // falcon9/sim.py:659-672
    r["gravity_and_frame_forces"] = make_spec(B200_EFF_GRAVITY_FRAME, {3.986004418e14, 0.0, 0.0, 7.292115e-5}, nullptr, 0); // falcon9/sim.py:350-361
    r["rw_effector"] = make_spec(B200_EFF_TORQUE_BODY_FOLD, {}, "wheel_torques", 9);                 // cube-sat/main.py:492-505 (3 wheels)
    r["j2_gravity"] = make_spec(B200_EFF_GRAVITY_J2, {3.986004418e14, 1.08262668e-3, 6.378e6}, nullptr, 0); // python/elodin/j2.py:5-29
    return r;
}

// "<function apply_drag at 0x7f..>" -> "apply_drag"; "<function map.<locals>.inner at 0x..>" -> "map.<locals>.inner";
// anything else is returned unchanged
inline std::string system_function_name(const std::string &system_name)
{
    const std::string pre = "<function ";
    if (system_name.compare(0, pre.size(), pre) != 0) return system_name;
    const size_t at = system_name.rfind(" at 0x");
    return system_name.substr(pre.size(), (at == std::string::npos ? system_name.size() - 1 : at) - pre.size());
}

// CompiledSystem.system_names -> built-in effectors, in pipeline order.  Structural entries of the six_dof()
// pipeline ("<system>", "<compiled>", "<empty>", clear_forces, calc_accel, the integrator wrappers) are skipped; every
// other system must be in the registry — the B200 arm has no tracing compiler and no CPU fallback, so an unknown
// system is a hard B200_ERR_UNSUPPORTED error that names it.  Edge-fold gravity is declared through
// `graph_effector` (its edge list comes from the world, not from a name).
inline std::vector<b200_effector> match_effectors(const std::vector<std::string> &system_names,
                                                  const std::map<std::string, EffectorSpec> &registry = default_effector_registry())
{
    static const char *structural[] = {"<system>", "<compiled>", "<empty>", "clear_forces", "calc_accel", "six_dof", "rk4",
                                        "semi_implicit_euler", "increment_sim_tick", "advance_time"};
    std::vector<b200_effector> out;
    for (const std::string &raw : system_names) {
        const std::string name = system_function_name(raw);
        if (std::any_of(std::begin(structural), std::end(structural), [&](const char *s) { return name == s; })) continue;
        auto it = registry.find(name);
        if (it == registry.end()) {
            std::string why = "system '" + name + "' is not a built-in B200 effector";
            if (name.find("<locals>") != std::string::npos)
                why += " (an @el.map wrapper: the wrapped function's name is not visible in system_names; register it under the name "
                       "you give it, or declare the effector list explicitly)";
            throw Error(B200_ERR_UNSUPPORTED, why + "; the B200 backend cannot trace arbitrary systems and has no CPU fallback");
        }
        b200_effector e = it->second.effector;
        e.column_id = it->second.column.empty() ? 0 : component_id(it->second.column.c_str());
        out.push_back(e);
        if (out.size() > B200_MAX_EFFECTORS)
            throw Error(B200_ERR_UNSUPPORTED, "too many effectors (" + std::to_string(out.size()) + " > " + std::to_string(B200_MAX_EFFECTORS) + ")");
    }
    return out;
}

// cranelift_exec.rs:13-195
class Exec {
  public:
    Exec(const World &w, std::vector<b200_effector> effectors, uint32_t integrator = B200_INTEGRATOR_RK4,
         uint32_t math_mode = B200_MATH_EXACT, double time_step = NAN, int device = -1)
    {
        b200_sixdof_desc d;
        std::memset(&d, 0, sizeof d);
        d.abi_version = B200_SIXDOF_ABI_VERSION;
        d.integrator = integrator;
        d.math_mode = math_mode;
        d.n_effectors = (uint32_t)effectors.size();
        d.n_entities = w.body_rows();
        d.n_worlds = w.n_worlds;
        d.sim_time_step = w.sim_time_step;
        d.time_step = time_step;
        d.effectors = effectors.data();
        d.device = device;
        d.max_fused_ticks = 32;
        check(b200_sixdof_create(&d, &h_));
        uint64_t ids[32];
        uint32_t n = b200_sixdof_input_ids(h_, ids, 32);
        input_ids.assign(ids, ids + n);
        n = b200_sixdof_output_ids(h_, ids, 32);
        output_ids.assign(ids, ids + n);
        for (ComponentId id : output_ids) { // executor-owned output buffers (cranelift_exec.rs:101-107)
            auto it = w.host.find(id);
            if (it == w.host.end()) { b200_sixdof_destroy(h_); h_ = nullptr; throw ComponentNotFound(); }
            if (b200_sixdof_column_bytes(h_, id) != it->second.buffer.size()) { b200_sixdof_destroy(h_); h_ = nullptr; throw ValueSizeMismatch(); }
            output_buffers.emplace_back(it->second.buffer.size());
        }
    }
    Exec(const Exec &) = delete;
    Exec &operator=(const Exec &) = delete;
    ~Exec() { if (h_) b200_sixdof_destroy(h_); }

    // invoke_batch (cranelift_exec.rs:129-195): borrowed inputs, n ticks, outputs copied back into the world
    void invoke_batch(World &world, uint64_t n)
    {
        world.set_globals();
        std::vector<const uint8_t *> ins;
        for (ComponentId id : input_ids) {
            Column *c = world.column_by_id(id);
            if (!c) throw ComponentNotFound();
            ins.push_back(c->buffer.data());
        }
        std::vector<uint8_t *> outs;
        for (auto &b : output_buffers) outs.push_back(b.data());
        check(b200_sixdof_invoke_batch(h_, ins.data(), outs.data(), n));
        for (size_t i = 0; i < output_ids.size(); ++i) {
            if (output_ids[i] == B200_ID_TICK || output_ids[i] == B200_ID_SIMULATION_TIME_STEP) continue; // advance_tick() owns the counter
            Column *host = world.column_by_id(output_ids[i]);
            if (!host) throw ComponentNotFound();
            if (host->buffer.size() != output_buffers[i].size()) throw ValueSizeMismatch();
            host->buffer = output_buffers[i];
        }
    }
    b200_sixdof *handle() { return h_; }

    std::vector<ComponentId> input_ids, output_ids;
    std::vector<std::vector<uint8_t>> output_buffers;

  private:
    b200_sixdof *h_ = nullptr;
};

// cranelift_exec.rs:270-303
class WorldExec {
  public:
    World world;
    Exec tick_exec;
    WorldExec(World w, std::vector<b200_effector> effectors, uint32_t integrator = B200_INTEGRATOR_RK4,
              uint32_t math_mode = B200_MATH_EXACT, double time_step = NAN, int device = -1)
        : world(std::move(w)), tick_exec(world, std::move(effectors), integrator, math_mode, time_step, device)
    {
    }
    void run()
    {
        const uint64_t n = world.ticks_per_telemetry;
        tick_exec.invoke_batch(world, n);
        for (uint64_t i = 0; i < n; ++i) world.advance_tick();
    }
};

} // namespace b200
