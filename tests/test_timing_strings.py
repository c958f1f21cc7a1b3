"""utils::TickDurationHistory::get_avg_time_str keeps the reference's overlay format
(/root/reference/src/utils/TickDurationHistory.cpp:36-55): "{:.2f}{unit}" in the largest of ns / us / ms / s that keeps the value >= 1."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <iostream>
#include "Timing.hpp"
int main() {
    const long long cases[] = {0, 999, 1000, 12300, 999999, 1500000, 999999999, 2000000000LL};
    for (long long ns : cases) {
        utils::TickDurationHistory h;
        h.add_time(std::chrono::nanoseconds(ns));
        std::cout << h.get_avg_time_str() << "\n";
    }
    utils::TickDurationHistory h;  // rolling mean over the window
    h.add_time(std::chrono::nanoseconds(10000));
    h.add_time(std::chrono::nanoseconds(14600));
    std::cout << h.get_avg_time_str() << "\n";
}
"""


def test_avg_time_str_scales_the_unit():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        with open(src, "w") as f:
            f.write(SRC)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "movement-sim_b200", "sim"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    assert out == ["0.00ns", "999.00ns", "1.00us", "12.30us", "1000.00us", "1.50ms", "1000.00ms", "2.00s", "12.30us"]
