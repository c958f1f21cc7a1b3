// Timing.hpp — utils::TickRate and utils::TickDurationHistory as the overlay widget reads them
// (/root/reference/src/utils/TickRate.*, TickDurationHistory.*; consumer
// src/ui/widgets/SimulationOverlayWidget.cpp:56-67): ticks per second over the last second and a
// 100-sample rolling mean of durations.
#pragma once

#include <array>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <string>

namespace utils {
class TickRate {
 public:
    void tick() {
        const auto now = std::chrono::steady_clock::now();
        counter_++;
        if (now - windowStart_ >= std::chrono::seconds(1)) {
            ticksPerSecond_ = static_cast<double>(counter_) / std::chrono::duration<double>(now - windowStart_).count();
            counter_ = 0;
            windowStart_ = now;
        }
    }
    [[nodiscard]] double get_ticks() const { return ticksPerSecond_; }

 private:
    std::chrono::steady_clock::time_point windowStart_{std::chrono::steady_clock::now()};
    size_t counter_{0};
    double ticksPerSecond_{0};
};

class TickDurationHistory {
 public:
    static constexpr size_t WINDOW = 100;
    void add_time(std::chrono::nanoseconds d) {
        sum_ += d - samples_[next_];
        samples_[next_] = d;
        next_ = (next_ + 1) % WINDOW;
        if (count_ < WINDOW) count_++;
    }
    [[nodiscard]] std::chrono::nanoseconds get_avg_time() const { return count_ ? sum_ / static_cast<int64_t>(count_) : std::chrono::nanoseconds(0); }
    // the overlay string, as the reference prints it (utils/TickDurationHistory.cpp:36-55): two decimals in the largest unit of
    // ns / us / ms / s that keeps the value >= 1, e.g. "12.30us"
    [[nodiscard]] std::string get_avg_time_str() const {
        static constexpr struct { double ns; const char* unit; } SCALES[] = {{1e9, "s"}, {1e6, "ms"}, {1e3, "us"}, {1.0, "ns"}};
        const double t = static_cast<double>(get_avg_time().count());
        for (const auto& sc : SCALES) {
            if (t >= sc.ns || sc.ns == 1.0) {
                char buf[48];
                std::snprintf(buf, sizeof(buf), "%.2f%s", t / sc.ns, sc.unit);
                return buf;
            }
        }
        return "0.00ns";
    }

 private:
    std::array<std::chrono::nanoseconds, WINDOW> samples_{};
    std::chrono::nanoseconds sum_{0};
    size_t next_{0};
    size_t count_{0};
};
}  // namespace utils
