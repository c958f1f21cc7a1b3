// Simulator.hpp — drop-in sim::Simulator: the public interface of the reference's host driver
// (/root/reference/src/sim/Simulator.hpp:92-119) kept verbatim — same names, signatures, return types
// and hand-off protocol — with the Kompute/Vulkan members replaced by one msim_handle (the C ABI of
// libmsim_cuda.so).  Callers that keep compiling unchanged: src/main.cpp:22-30,36-44 and
// src/ui/widgets/*.cpp.
//
// What the reference hard-codes becomes run-time configuration with the reference's values as
// defaults (SimulatorConfig): entity count (Simulator.hpp:33), quadtree depth/cap (:37-38), collision
// radius (:43), map path (Simulator.cpp:55), plus seed / tick limit / collisions on-off / device.
#pragma once

#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "DataModel.hpp"
#include "Timing.hpp"

namespace sim {
enum class SimulatorState { STOPPED, RUNNING, JOINING };

constexpr size_t MAX_ENTITIES = 1000000;           // default entity count (reference: compile-time)
constexpr float MAX_RENDER_RESOLUTION_X = 8192;
constexpr float MAX_RENDER_RESOLUTION_Y = 8192;
constexpr size_t QUAD_TREE_MAX_DEPTH = 8;
constexpr size_t QUAD_TREE_ENTITY_NODE_CAP = 10;
constexpr float COLLISION_RADIUS = 10;  // metres

struct SimulatorConfig {
    std::filesystem::path mapPath{"munich.json"};  // env MSIM_MAP; "synthetic:city" / "synthetic:grid:<nx>x<ny>" generate one
    size_t entities{MAX_ENTITIES};                 // env MSIM_ENTITIES
    uint64_t seed{0};                              // env MSIM_SEED; 0 = std::random_device like the reference
    uint64_t tickLimit{0};                         // env MSIM_TICKS; 0 = run until stopped (reference behaviour)
    bool collisions{true};                         // env MSIM_COLLISIONS=0: only the move dispatch runs
    int device{0};                                 // env MSIM_DEVICE
    float collisionRadius{COLLISION_RADIUS};
    std::filesystem::path csvPath{};               // empty = "<entities>.csv" (Simulator.cpp:337-340)
    bool quiet{false};                             // do not echo CSV rows to stderr
    bool asyncReadback{false};                     // env MSIM_ASYNC_READBACK=1: entity readback through msim_snapshot_* (the copy overlaps
                                                   // the following ticks; get_entities() then lags the simulation by a few ticks)

    static SimulatorConfig from_environment();
};

class Simulator {
 public:
    Simulator();
    explicit Simulator(SimulatorConfig config);
    ~Simulator();

    Simulator(Simulator&&) = delete;
    Simulator(const Simulator&) = delete;
    Simulator& operator=(Simulator&&) = delete;
    Simulator& operator=(const Simulator&) = delete;

    void init();

    static std::shared_ptr<Simulator>& get_instance();
    [[nodiscard]] SimulatorState get_state() const;
    void start_worker();
    void stop_worker();

    void continue_simulation();
    void pause_simulation();
    [[nodiscard]] bool is_simulating() const;
    [[nodiscard]] const utils::TickRate& get_tps() const;
    [[nodiscard]] const utils::TickDurationHistory& get_tps_history() const;
    [[nodiscard]] const utils::TickDurationHistory& get_update_tick_history() const;
    [[nodiscard]] const utils::TickDurationHistory& get_collision_detection_tick_history() const;
    std::shared_ptr<std::vector<Entity>> get_entities();
    std::shared_ptr<std::vector<gpu_quad_tree::Node>> get_quad_tree_nodes();
    [[nodiscard]] const std::shared_ptr<Map> get_map() const;

    [[nodiscard]] bool is_initialized() const;

    // additions (not in the reference)
    [[nodiscard]] uint64_t get_completed_ticks() const { return completedTicks; }
    [[nodiscard]] bool reached_tick_limit() const { return config.tickLimit != 0 && completedTicks >= config.tickLimit; }
    void read_entities_now(std::vector<Entity>& out);  // blocking readback outside the worker (tests / headless dump)
    [[nodiscard]] const SimulatorConfig& get_config() const { return config; }

 private:
    void sim_worker();
    void sim_tick();
    void add_entities();
    void prepare_log_csv_file();
    void write_log_csv_file(uint32_t tick, std::chrono::nanoseconds durationUpdate, std::chrono::nanoseconds durationCollision,
                            std::chrono::nanoseconds durationAll);
    static std::string get_time_stamp();
    void check(int status, const char* what) const;

    SimulatorConfig config;
    bool initialized{false};
    std::unique_ptr<std::ofstream> logFile{nullptr};

    std::unique_ptr<std::thread> simThread{nullptr};
    SimulatorState state{SimulatorState::STOPPED};
    std::mutex waitMutex{};
    std::condition_variable waitCondVar{};
    bool simulating{false};

    utils::TickDurationHistory tpsHistory{};
    utils::TickRate tps{};
    utils::TickDurationHistory updateTickHistory{};
    utils::TickDurationHistory collisionDetectionTickHistory{};

    msim_handle* handle{nullptr};  // replaces kp::Manager / kp::Tensor / kp::Algorithm / kp::Sequence
    PushConsts pushConsts{};
    uint64_t completedTicks{0};
    bool snapshotPending{false};  // asyncReadback: a msim_snapshot_begin has not been collected yet

    std::mutex handoffMutex{};  // the reference hands these two pointers over unsynchronised
    std::shared_ptr<std::vector<Entity>> entities{std::make_shared<std::vector<Entity>>()};
    std::shared_ptr<std::vector<gpu_quad_tree::Node>> quadTreeNodes{std::make_shared<std::vector<gpu_quad_tree::Node>>()};
    std::shared_ptr<Map> map{nullptr};
};
}  // namespace sim
