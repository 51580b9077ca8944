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

#include <cmath>
#include <cstdint>
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
