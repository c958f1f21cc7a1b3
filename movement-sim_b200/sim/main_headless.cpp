// main_headless.cpp — headless runner over the drop-in sim::Simulator.  The reference's main()
// (/root/reference/src/main.cpp:10-56) knows one flag, --headless, and then loops forever; this runner
// keeps that flag and adds what the benchmarks need: map, entity count, seed,
// tick limit, collisions on/off, device, CSV path and a binary dump of the final entity buffer.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>

#include "Simulator.hpp"

namespace {
void usage() {
    std::cerr << "usage: msim_headless [--headless] [--map PATH|synthetic:city|synthetic:grid:NXxNY] [--entities N] [--seed S]\n"
                 "                     [--ticks T] [--no-collisions] [--device D] [--csv PATH] [--dump PATH] [--quiet]\n"
                 "                     [--consume-entities] [--async-readback]\n";
}
}  // namespace

int main(int argc, char** argv) {
    sim::SimulatorConfig cfg = sim::SimulatorConfig::from_environment();
    std::string dumpPath;
    bool consumeEntities = false;  // take the entity buffer like the UI does every frame (EntityGlObject.cpp:17,22): exercises the readback path
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) {
                usage();
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--headless") continue;  // the only mode this runner has
        else if (a == "--map") cfg.mapPath = next();
        else if (a == "--entities") cfg.entities = std::strtoull(next(), nullptr, 10);
        else if (a == "--seed") cfg.seed = std::strtoull(next(), nullptr, 10);
        else if (a == "--ticks") cfg.tickLimit = std::strtoull(next(), nullptr, 10);
        else if (a == "--no-collisions") cfg.collisions = false;
        else if (a == "--device") cfg.device = std::atoi(next());
        else if (a == "--csv") cfg.csvPath = next();
        else if (a == "--dump") dumpPath = next();
        else if (a == "--quiet") cfg.quiet = true;
        else if (a == "--consume-entities") consumeEntities = true;
        else if (a == "--async-readback") cfg.asyncReadback = true;
        else {
            usage();
            return 2;
        }
    }
    try {
        sim::Simulator simulator(cfg);
        simulator.init();
        simulator.start_worker();
        simulator.continue_simulation();
        const auto begin = std::chrono::steady_clock::now();
        uint64_t frames = 0;
        while (!simulator.reached_tick_limit()) {
            if (consumeEntities && simulator.get_entities()) frames++;  // the worker refills the buffer after a later tick
            std::this_thread::sleep_for(std::chrono::milliseconds(cfg.tickLimit ? 2 : 100));
        }
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
        simulator.pause_simulation();
        simulator.stop_worker();
        const double updates = static_cast<double>(cfg.entities) * static_cast<double>(simulator.get_completed_ticks());
        std::cout << "ticks=" << simulator.get_completed_ticks() << " entities=" << cfg.entities << " seconds=" << seconds
                  << " entity_updates_per_s=" << (seconds > 0 ? updates / seconds : 0.0) << " avg_update=" << simulator.get_update_tick_history().get_avg_time_str()
                  << " avg_collision=" << simulator.get_collision_detection_tick_history().get_avg_time_str() << " entity_frames=" << frames << '\n';
        if (!dumpPath.empty()) {
            std::vector<sim::Entity> out;
            simulator.read_entities_now(out);
            std::FILE* f = std::fopen(dumpPath.c_str(), "wb");
            if (!f || std::fwrite(out.data(), sizeof(sim::Entity), out.size(), f) != out.size()) {
                std::cerr << "cannot write " << dumpPath << '\n';
                return 1;
            }
            std::fclose(f);
        }
    } catch (const std::exception& e) {
        std::cerr << "msim_headless: " << e.what() << '\n';
        return 1;
    }
    return 0;
}
