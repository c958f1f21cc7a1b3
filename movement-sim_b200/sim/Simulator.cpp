// Simulator.cpp — host driver over the C ABI.  Behaviour mirrored from
// /root/reference/src/sim/Simulator.cpp (cited per function); every Kompute call is replaced by the
// msim_* entry point named in include/msim.h.
#include "Simulator.hpp"

#include <cassert>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <stdexcept>

namespace sim {
namespace {
const char* env_or_null(const char* name) {
    const char* v = std::getenv(name);
    return (v && *v) ? v : nullptr;
}

void log_line(const std::string& text) { std::cerr << "[msim] " << text << '\n'; }
}  // namespace

// ---- Map ------------------------------------------------------------------------------------------
std::shared_ptr<Map> Map::adopt(msim_map* handle) {
    auto result = std::make_shared<Map>();
    result->width = msim_map_width(handle);
    result->height = msim_map_height(handle);
    const size_t roadCount = msim_map_road_count(handle);
    const size_t connCount = msim_map_connection_count(handle);
    result->roads.resize(roadCount);
    if (roadCount) std::memcpy(static_cast<void*>(result->roads.data()), msim_map_roads(handle), roadCount * sizeof(Road));
    result->connections.assign(msim_map_connections(handle), msim_map_connections(handle) + connCount);
    result->roadPieces.reserve(roadCount * 2);
    for (const Road& road : result->roads) {  // red start / end markers (Map.cpp:130-132)
        result->roadPieces.push_back(RoadPiece{road.start.pos, {}, Rgba{1.0F, 0.0F, 0.0F, 1.0F}});
        result->roadPieces.push_back(RoadPiece{road.end.pos, {}, Rgba{1.0F, 0.0F, 0.0F, 1.0F}});
    }
    msim_map_free(handle);
    return result;
}

std::shared_ptr<Map> Map::load_from_file(const std::filesystem::path& path) {
    log_line("Loading map from '" + path.string() + "'...");
    msim_map* handle = nullptr;
    // the reference's map JSON (Map.cpp:28-150), a GeoJSON export (map/generate_map.py) or a binary map cache
    const int status = msim_map_load(path.c_str(), &handle);
    if (status == MSIM_ERR_IO) {
        log_line(msim_map_last_error());
        return nullptr;
    }
    if (status != MSIM_OK) throw std::runtime_error(msim_map_last_error());
    std::shared_ptr<Map> result = adopt(handle);
    log_line("Map loaded. Found " + std::to_string(result->roads.size()) + " roads with " + std::to_string(result->connections.size()) + " connections.");
    return result;
}

std::shared_ptr<Map> Map::generate_city(float width, float height, uint64_t seed) {
    msim_map* handle = nullptr;
    if (msim_map_generate_city(width, height, 35.0F, 0.3F, 0.12F, seed, &handle) != MSIM_OK) throw std::runtime_error(msim_map_last_error());
    return adopt(handle);
}

std::shared_ptr<Map> Map::generate_grid(uint32_t nx, uint32_t ny, float spacing) {
    msim_map* handle = nullptr;
    if (msim_map_generate_grid(nx, ny, spacing, &handle) != MSIM_OK) throw std::runtime_error(msim_map_last_error());
    return adopt(handle);
}

unsigned int Map::get_random_road_index() const {  // Map.cpp:152-157
    static std::random_device device;
    static std::mt19937 gen(device());
    std::uniform_int_distribution<unsigned int> distr(0, static_cast<unsigned int>(roads.size() - 1));
    return distr(gen);
}

void Map::select_road(size_t roadIndex) {  // Map.cpp:159-175
    assert(roadIndex < roads.size());
    const Rgba plain{1.0F, 0.0F, 0.0F, 1.0F};
    const Rgba chosen{0.0F, 1.0F, 0.0F, 1.0F};
    if (selectedRoad) {
        roadPieces[*selectedRoad * 2].color = plain;
        roadPieces[*selectedRoad * 2 + 1].color = plain;
    }
    selectedRoad = roadIndex;
    roadPieces[roadIndex * 2].color = chosen;
    roadPieces[roadIndex * 2 + 1].color = chosen;
}

// ---- configuration --------------------------------------------------------------------------------
SimulatorConfig SimulatorConfig::from_environment() {
    SimulatorConfig c;
    if (const char* v = env_or_null("MSIM_MAP")) c.mapPath = v;
    if (const char* v = env_or_null("MSIM_ENTITIES")) c.entities = std::strtoull(v, nullptr, 10);
    if (const char* v = env_or_null("MSIM_SEED")) c.seed = std::strtoull(v, nullptr, 10);
    if (const char* v = env_or_null("MSIM_TICKS")) c.tickLimit = std::strtoull(v, nullptr, 10);
    if (const char* v = env_or_null("MSIM_COLLISIONS")) c.collisions = std::strcmp(v, "0") != 0;
    if (const char* v = env_or_null("MSIM_DEVICE")) c.device = std::atoi(v);
    if (const char* v = env_or_null("MSIM_CSV")) c.csvPath = v;
    if (const char* v = env_or_null("MSIM_ASYNC_READBACK")) c.asyncReadback = std::strcmp(v, "0") != 0;
    return c;
}

// ---- Simulator ------------------------------------------------------------------------------------
Simulator::Simulator() : Simulator(SimulatorConfig::from_environment()) {}

Simulator::Simulator(SimulatorConfig cfg) : config(std::move(cfg)) { prepare_log_csv_file(); }

Simulator::~Simulator() {
    if (simThread) stop_worker();
    if (handle) msim_destroy(handle);
    if (logFile) logFile->close();
}

void Simulator::check(int status, const char* what) const {
    if (status == MSIM_OK) return;
    // Kompute throws on Vulkan failure; the C ABI reports a status, which becomes an exception here.
    throw std::runtime_error(std::string(what) + ": " + msim_status_string(status) + " - " + msim_last_error(handle));
}

void Simulator::init() {  // Simulator.cpp:44-108
    assert(!initialized);
    const std::string spec = config.mapPath.string();
    if (spec == "synthetic:city") {
        map = Map::generate_city(29007.4609F, 16463.7656F, 2022);  // world size: shader_validation/src/main.cpp:155
    } else if (spec.rfind("synthetic:grid:", 0) == 0) {
        unsigned nx = 0, ny = 0;
        if (std::sscanf(spec.c_str() + 15, "%ux%u", &nx, &ny) != 2) throw std::runtime_error("bad grid spec, want synthetic:grid:<nx>x<ny>");
        map = Map::generate_grid(nx, ny, 20.0F);
    } else {
        map = Map::load_from_file(config.mapPath);
    }
    if (!map) throw std::runtime_error("map '" + spec + "' could not be loaded");

    add_entities();

    quadTreeNodes->resize(gpu_quad_tree::calc_node_count(QUAD_TREE_MAX_DEPTH));
    (*quadTreeNodes)[0].width = map->width;  // gpu_quad_tree::init_node_zero (GpuQuadTree.cpp:6-10)
    (*quadTreeNodes)[0].height = map->height;
    (*quadTreeNodes)[0].contentType = gpu_quad_tree::NextType::ENTITY;

    pushConsts.worldSizeX = map->width;  // Simulator.cpp:94-101
    pushConsts.worldSizeY = map->height;
    pushConsts.nodeCount = static_cast<uint32_t>(quadTreeNodes->size());
    pushConsts.maxDepth = QUAD_TREE_MAX_DEPTH;
    pushConsts.entityNodeCap = QUAD_TREE_ENTITY_NODE_CAP;
    pushConsts.collisionRadius = config.collisionRadius;
    pushConsts.tick = 1;

    msim_config mc{};
    mc.abi_version = MSIM_ABI_VERSION;
    mc.device = config.device;
    mc.flags = config.collisions ? 0U : static_cast<uint32_t>(MSIM_FLAG_NO_COLLISIONS);
    mc.world_w = map->width;
    mc.world_h = map->height;
    mc.collision_radius = config.collisionRadius;
    mc.quadtree_max_depth = QUAD_TREE_MAX_DEPTH;
    mc.quadtree_node_cap = QUAD_TREE_ENTITY_NODE_CAP;
    mc.roads = reinterpret_cast<const msim_road*>(map->roads.data());
    mc.road_count = map->roads.size();
    mc.connections = map->connections.data();
    mc.connection_count = map->connections.size();
    mc.entities = reinterpret_cast<const msim_entity*>(entities->data());
    mc.entity_count = entities->size();
    // kp::Manager + 7 tensors + algorithm + the one-off OpTensorSyncDevice (Simulator.cpp:52-103,191-192)
    const int status = msim_create(&mc, &handle);
    if (status != MSIM_OK) throw std::runtime_error(std::string("msim_create: ") + msim_status_string(status) + " - " + msim_last_error(nullptr));
    initialized = true;
}

bool Simulator::is_initialized() const { return initialized; }

void Simulator::add_entities() {  // Simulator.cpp:114-129
    assert(map);
    entities->resize(config.entities);
    uint64_t seed = config.seed;
    if (seed == 0) {  // the reference seeds from std::random_device (Entity.cpp:17-18,45-47,52-53)
        std::random_device device;
        seed = (static_cast<uint64_t>(device()) << 16) ^ device();
        if (seed == 0) seed = 1;
    }
    const int status = msim_entities_init(reinterpret_cast<const msim_road*>(map->roads.data()), map->roads.size(), config.entities, seed, nullptr,
                                          reinterpret_cast<msim_entity*>(entities->data()));
    if (status != MSIM_OK) throw std::runtime_error(msim_map_last_error());
}

std::shared_ptr<Simulator>& Simulator::get_instance() {  // Simulator.cpp:131-137
    static std::shared_ptr<Simulator> instance = std::make_shared<Simulator>();
    if (!instance->is_initialized()) instance->init();
    return instance;
}

SimulatorState Simulator::get_state() const { return state; }

std::shared_ptr<std::vector<Entity>> Simulator::get_entities() {  // Simulator.cpp:143-147: take and null
    std::lock_guard<std::mutex> guard(handoffMutex);
    std::shared_ptr<std::vector<Entity>> result = std::move(entities);
    entities = nullptr;
    return result;
}

std::shared_ptr<std::vector<gpu_quad_tree::Node>> Simulator::get_quad_tree_nodes() {  // Simulator.cpp:149-153
    std::lock_guard<std::mutex> guard(handoffMutex);
    std::shared_ptr<std::vector<gpu_quad_tree::Node>> result = std::move(quadTreeNodes);
    quadTreeNodes = nullptr;
    return result;
}

const std::shared_ptr<Map> Simulator::get_map() const { return map; }

void Simulator::start_worker() {  // Simulator.cpp:159-167
    assert(initialized);
    assert(state == SimulatorState::STOPPED);
    assert(!simThread);
    log_line("Starting simulation thread...");
    state = SimulatorState::RUNNING;
    simThread = std::make_unique<std::thread>(&Simulator::sim_worker, this);
}

void Simulator::stop_worker() {  // Simulator.cpp:169-183
    assert(initialized);
    assert(simThread);
    log_line("Stopping simulation thread...");
    {
        std::lock_guard<std::mutex> guard(waitMutex);
        state = SimulatorState::JOINING;
    }
    waitCondVar.notify_all();
    if (simThread->joinable()) simThread->join();
    simThread = nullptr;
    state = SimulatorState::STOPPED;
    log_line("Simulation thread stopped.");
}

void Simulator::sim_worker() {  // Simulator.cpp:185-211 (the upload already happened in msim_create)
    assert(initialized);
    log_line("Simulation thread started.");
    std::unique_lock<std::mutex> lk(waitMutex);
    while (state == SimulatorState::RUNNING) {
        if (!simulating || reached_tick_limit()) waitCondVar.wait_for(lk, std::chrono::milliseconds(50));
        if (!simulating || reached_tick_limit()) continue;
        lk.unlock();
        sim_tick();
        lk.lock();
    }
}

void Simulator::sim_tick() {  // Simulator.cpp:213-279
    using clock = std::chrono::high_resolution_clock;
    const clock::time_point tickStart = clock::now();

    // update dispatch (even tick): calcSeq->eval<OpAlgoDispatch>(algo, pushConsts), blocking (:220-224)
    pushConsts.tick++;
    const clock::time_point updateStart = clock::now();
    check(msim_dispatch(handle, reinterpret_cast<const msim_push_consts*>(&pushConsts)), "update dispatch");
    const std::chrono::nanoseconds durationUpdate = clock::now() - updateStart;
    updateTickHistory.add_time(durationUpdate);

    // collision dispatch (odd tick) (:231-235); with collisions off only the tick counter advances
    pushConsts.tick++;
    const clock::time_point collisionStart = clock::now();
    if (config.collisions) check(msim_dispatch(handle, reinterpret_cast<const msim_push_consts*>(&pushConsts)), "collision dispatch");
    const std::chrono::nanoseconds durationCollision = clock::now() - collisionStart;
    collisionDetectionTickHistory.add_time(durationCollision);

    write_log_csv_file(pushConsts.tick, durationUpdate, durationCollision, durationUpdate + durationCollision);

    // readbacks only when the consumer took the previous copy (:248-268); the unconditional 20 MB
    // per-tick readback of quadtree links (:258,270-272) has no consumer and is not reproduced
    bool wantEntities = false, wantNodes = false;
    {
        std::lock_guard<std::mutex> guard(handoffMutex);
        wantEntities = !entities;
        wantNodes = !quadTreeNodes;
    }
    if (config.asyncReadback) {
        // the reference overlaps nothing here (evalAsync + evalAwait back to back, :250-253); with a snapshot the 64 B x N copy
        // leaves through the copy engine while the next ticks run, and is handed over once it has landed
        if (snapshotPending && wantEntities) {
            int ready = 0;
            check(msim_snapshot_poll(handle, &ready), "entity readback (poll)");
            if (ready) {
                const msim_entity* src = nullptr;
                uint64_t n = 0;
                check(msim_snapshot_end(handle, &src, &n), "entity readback (collect)");
                snapshotPending = false;
                const Entity* first = reinterpret_cast<const Entity*>(src);
                auto fresh = std::make_shared<std::vector<Entity>>(first, first + n);
                std::lock_guard<std::mutex> guard(handoffMutex);
                entities = std::move(fresh);
            }
        }
        if (!snapshotPending && wantEntities) {
            check(msim_snapshot_begin(handle), "entity readback (begin)");
            snapshotPending = true;
        }
    } else if (wantEntities) {
        auto fresh = std::make_shared<std::vector<Entity>>(config.entities);
        check(msim_read_entities(handle, reinterpret_cast<msim_entity*>(fresh->data()), fresh->size()), "entity readback");
        std::lock_guard<std::mutex> guard(handoffMutex);
        entities = std::move(fresh);
    }
    if (wantNodes) {
        auto fresh = std::make_shared<std::vector<gpu_quad_tree::Node>>(gpu_quad_tree::calc_node_count(QUAD_TREE_MAX_DEPTH));
        uint64_t count = 0;
        check(msim_read_quadtree_nodes(handle, reinterpret_cast<msim_quadtree_node*>(fresh->data()), fresh->size(), &count), "quadtree readback");
        std::lock_guard<std::mutex> guard(handoffMutex);
        quadTreeNodes = std::move(fresh);
    }
    uint32_t debugData[10];
    check(msim_read_debug(handle, debugData), "debug readback");  // tensorDebugData (:273)

    completedTicks++;
    tpsHistory.add_time(clock::now() - tickStart);
    tps.tick();
}

void Simulator::read_entities_now(std::vector<Entity>& out) {
    out.resize(config.entities);
    check(msim_read_entities(handle, reinterpret_cast<msim_entity*>(out.data()), out.size()), "entity readback");
}

void Simulator::continue_simulation() {  // Simulator.cpp:281-287
    if (simulating) return;
    {
        std::lock_guard<std::mutex> guard(waitMutex);
        simulating = true;
    }
    waitCondVar.notify_all();
}

void Simulator::pause_simulation() {  // Simulator.cpp:289-295
    if (!simulating) return;
    {
        std::lock_guard<std::mutex> guard(waitMutex);
        simulating = false;
    }
    waitCondVar.notify_all();
}

bool Simulator::is_simulating() const { return simulating; }
const utils::TickRate& Simulator::get_tps() const { return tps; }
const utils::TickDurationHistory& Simulator::get_tps_history() const { return tpsHistory; }
const utils::TickDurationHistory& Simulator::get_update_tick_history() const { return updateTickHistory; }
const utils::TickDurationHistory& Simulator::get_collision_detection_tick_history() const { return collisionDetectionTickHistory; }

void Simulator::prepare_log_csv_file() {  // Simulator.cpp:337-347
    const std::filesystem::path path = config.csvPath.empty() ? std::filesystem::path(std::to_string(config.entities) + ".csv") : config.csvPath;
    logFile = std::make_unique<std::ofstream>(path, std::ios::out | std::ios::app);
}

std::string Simulator::get_time_stamp() {  // hh:mm:ss.mmm of the current UTC day (Simulator.cpp:349-387)
    using namespace std::chrono;
    const auto sinceEpoch = system_clock::now().time_since_epoch();
    const auto msOfDay = duration_cast<milliseconds>(sinceEpoch) % hours(24);
    char text[32];
    std::snprintf(text, sizeof(text), "%02lld:%02lld:%02lld.%03lld", static_cast<long long>(duration_cast<hours>(msOfDay).count()),
                  static_cast<long long>(duration_cast<minutes>(msOfDay).count() % 60), static_cast<long long>(duration_cast<seconds>(msOfDay).count() % 60),
                  static_cast<long long>(msOfDay.count() % 1000));
    return text;
}

void Simulator::write_log_csv_file(uint32_t tick, std::chrono::nanoseconds durationUpdate, std::chrono::nanoseconds durationCollision,
                                   std::chrono::nanoseconds durationAll) {  // Simulator.cpp:389-400: time;tick/2;secUpdate;secCollision;secAll
    // microsecond resolution instead of the reference's whole milliseconds: a B200 tick is sub-millisecond
    const auto sec = [](std::chrono::nanoseconds d) { return static_cast<double>(std::chrono::duration_cast<std::chrono::microseconds>(d).count()) / 1e6; };
    const std::string row = get_time_stamp() + ";" + std::to_string(tick / 2) + ";" + std::to_string(sec(durationUpdate)) + ";" +
                            std::to_string(sec(durationCollision)) + ";" + std::to_string(sec(durationAll)) + "\n";
    if (logFile && logFile->good()) {
        (*logFile) << row;
        if ((tick / 2) % 64 == 0) logFile->flush();
    }
    if (!config.quiet) std::cerr << row;
}
}  // namespace sim
