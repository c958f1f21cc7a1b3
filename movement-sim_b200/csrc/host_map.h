// host_map.h — internals shared by the host-only translation units of libmsim_cuda.so (host_map.cpp: map JSON,
// synthetic generators, entity initialiser; host_mapgen.cpp: GeoJSON -> map pipeline and the binary map cache).
#pragma once

#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/msim.h"

struct msim_map {
    float width{0};
    float height{0};
    std::vector<msim_road> roads;
    std::vector<uint32_t> connections;
};

namespace msim_host {

// records the message msim_map_last_error() returns on this thread; returns `code`
int map_fail(int code, const std::string& msg);
// whole file into `text`; false (with the reference loader's message, Map.cpp:30-38) when it cannot be opened
bool read_file(const char* path, std::string& text);

// ---------------------------------------------------------------------------------------------
// Minimal JSON reader for the map schema (the reference uses nlohmann::json; any conforming parser
// yields the same doubles, which are then narrowed to float exactly like json::get_to<float>).
// ---------------------------------------------------------------------------------------------
class JsonCursor {
 public:
    JsonCursor(const char* begin, const char* end) : p_(begin), end_(end) {}

    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) p_++;
    }
    bool eof() {
        ws();
        return p_ >= end_;
    }
    char peek() {
        ws();
        return p_ < end_ ? *p_ : '\0';
    }
    bool consume(char c) {
        if (peek() == c) {
            p_++;
            return true;
        }
        return false;
    }
    void expect(char c) {
        if (!consume(c)) fail(std::string("expected '") + c + "'");
    }
    std::string string() {
        expect('"');
        std::string out;
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\' && p_ + 1 < end_) {
                p_++;
                switch (*p_) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'u': p_ += 4; out.push_back('?'); break;
                    default: out.push_back(*p_); break;
                }
                p_++;
            } else {
                out.push_back(*p_++);
            }
        }
        if (p_ >= end_) fail("unterminated string");
        p_++;
        return out;
    }
    double number() {
        ws();
        char* stop = nullptr;
        errno = 0;
        const double v = std::strtod(p_, &stop);
        if (stop == p_) fail("expected a number");
        p_ = stop;
        return v;
    }
    void skip_value() {
        const char c = peek();
        if (c == '{') {
            p_++;
            if (consume('}')) return;
            do {
                (void)string();
                expect(':');
                skip_value();
            } while (consume(','));
            expect('}');
        } else if (c == '[') {
            p_++;
            if (consume(']')) return;
            do {
                skip_value();
            } while (consume(','));
            expect(']');
        } else if (c == '"') {
            (void)string();
        } else if (c == 't' || c == 'f' || c == 'n') {
            while (p_ < end_ && std::isalpha(static_cast<unsigned char>(*p_))) p_++;
        } else {
            (void)number();
        }
    }
    [[noreturn]] void fail(const std::string& what) { throw std::runtime_error("Failed to parse map. JSON syntax: " + what); }

    const char* pos() const { return p_; }

 private:
    const char* p_;
    const char* end_;
};

}  // namespace msim_host
