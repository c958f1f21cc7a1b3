// host_map.cpp — host-side map model of libmsim_cuda.so: JSON load/save, synthetic generators and
// the seeded entity initialiser.  No CUDA in this file.
//
// Mirrors (behaviour, not code) of the reference:
//   Map::load_from_file          /root/reference/src/sim/Map.cpp:28-150   (schema, float narrowing,
//                                zero-length-road skip :124-128, error texts :43-117)
//   connection-table layout      /root/reference/map/generate_map.py:234-258
//   Simulator::add_entities      /root/reference/src/sim/Simulator.cpp:114-129
//   Map::get_random_road_index   /root/reference/src/sim/Map.cpp:152-157
//   Rgba::random_color / Vec4U::random_vec   /root/reference/src/sim/Entity.cpp:44-60
#include "host_map.h"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace msim_host {
namespace {
thread_local std::string g_map_error;
}
int map_fail(int code, const std::string& msg) {
    g_map_error = msg;
    return code;
}
const std::string& map_error() { return g_map_error; }
bool read_file(const char* path, std::string& text) {
    std::FILE* f = std::fopen(path, "rb");
    if (!f) {
        map_fail(MSIM_ERR_IO, std::string("Failed to open map from '") + path + "'. File does not exist.");
        return false;
    }
    char buf[1 << 16];
    size_t got = 0;
    while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    std::fclose(f);
    return true;
}
}  // namespace msim_host

namespace {
using msim_host::JsonCursor;
using msim_host::map_fail;

struct Present {
    bool v{false};
};

[[noreturn]] void missing(const char* field) {
    throw std::runtime_error(std::string("Failed to parse map. '") + field + "' field missing.");
}

// {"lat":..,"long":..,"distLat":..,"distLong":..}  ->  x = distLat, y = distLong (Map.cpp:121-122)
void parse_point(JsonCursor& js, float& x, float& y) {
    bool hasLat = false, hasLong = false;
    js.expect('{');
    if (!js.consume('}')) {
        do {
            const std::string key = js.string();
            js.expect(':');
            if (key == "distLat") {
                x = static_cast<float>(js.number());
                hasLat = true;
            } else if (key == "distLong") {
                y = static_cast<float>(js.number());
                hasLong = true;
            } else {
                js.skip_value();
            }
        } while (js.consume(','));
        js.expect('}');
    }
    if (!hasLat) missing("distLat");
    if (!hasLong) missing("distLong");
}

uint32_t parse_u32(JsonCursor& js) {
    const double v = js.number();
    if (v < 0 || v > 4294967295.0) throw std::runtime_error("Failed to parse map. Unsigned value out of range.");
    return static_cast<uint32_t>(v);
}

void parse_road(JsonCursor& js, msim_map& m) {
    bool hStart = false, hEnd = false, hIs = false, hCs = false, hIe = false, hCe = false;
    msim_road r{};
    js.expect('{');
    if (!js.consume('}')) {
        do {
            const std::string key = js.string();
            js.expect(':');
            if (key == "start") {
                parse_point(js, r.start.pos[0], r.start.pos[1]);
                hStart = true;
            } else if (key == "end") {
                parse_point(js, r.end.pos[0], r.end.pos[1]);
                hEnd = true;
            } else if (key == "connIndexStart") {
                r.start.connected_index = parse_u32(js);
                hIs = true;
            } else if (key == "connCountStart") {
                r.start.connected_count = parse_u32(js);
                hCs = true;
            } else if (key == "connIndexEnd") {
                r.end.connected_index = parse_u32(js);
                hIe = true;
            } else if (key == "connCountEnd") {
                r.end.connected_count = parse_u32(js);
                hCe = true;
            } else {
                js.skip_value();
            }
        } while (js.consume(','));
        js.expect('}');
    }
    if (!hIs) missing("connIndexStart");
    if (!hCs) missing("connCountStart");
    if (!hStart) missing("start");
    if (!hIe) missing("connIndexEnd");
    if (!hCe) missing("connCountEnd");
    if (!hEnd) missing("end");
    // zero-length roads are dropped WITHOUT re-indexing the connection table
    if (r.start.pos[0] == r.end.pos[0] && r.start.pos[1] == r.end.pos[1]) return;
    m.roads.push_back(r);
}

void parse_map(JsonCursor& js, msim_map& m) {
    bool hW = false, hH = false, hRoads = false, hConn = false;
    js.expect('{');
    if (!js.consume('}')) {
        do {
            const std::string key = js.string();
            js.expect(':');
            if (key == "maxDistLat") {
                m.width = static_cast<float>(js.number());
                hW = true;
            } else if (key == "maxDistLong") {
                m.height = static_cast<float>(js.number());
                hH = true;
            } else if (key == "roads") {
                js.expect('[');
                if (!js.consume(']')) {
                    do {
                        parse_road(js, m);
                    } while (js.consume(','));
                    js.expect(']');
                }
                hRoads = true;
            } else if (key == "connectionRoadIndexList") {
                js.expect('[');
                if (!js.consume(']')) {
                    do {
                        m.connections.push_back(parse_u32(js));
                    } while (js.consume(','));
                    js.expect(']');
                }
                hConn = true;
            } else {
                js.skip_value();
            }
        } while (js.consume(','));
        js.expect('}');
    }
    if (!hW) missing("maxDistLat");
    if (!hH) missing("maxDistLong");
    if (!hRoads) missing("roads");
    if (!hConn) missing("connectionRoadIndexList");
}

// ---------------------------------------------------------------------------------------------
// Connection table in the layout of map/generate_map.py:234-258: one shared block per coordinate;
// every road of the coordinate is appended once, and a second time when the coordinate is the
// road's END; connIndex* = block start, connCount* = number of roads at the coordinate.
// ---------------------------------------------------------------------------------------------
struct Edge {
    uint32_t a, b;  // node ids: a = start, b = end
};

void emit_tables(const std::vector<float>& nx, const std::vector<float>& ny, const std::vector<Edge>& edges, msim_map& m) {
    const size_t nodeCount = nx.size();
    std::vector<uint32_t> degree(nodeCount, 0);
    for (const Edge& e : edges) {
        degree[e.a]++;
        degree[e.b]++;
    }
    std::vector<uint64_t> adjStart(nodeCount + 1, 0);
    for (size_t i = 0; i < nodeCount; i++) adjStart[i + 1] = adjStart[i] + degree[i];
    std::vector<uint32_t> adj(adjStart[nodeCount]);
    std::vector<uint64_t> fill(adjStart.begin(), adjStart.end() - 1);
    for (uint32_t r = 0; r < edges.size(); r++) {
        adj[fill[edges[r].a]++] = r;
        adj[fill[edges[r].b]++] = r;
    }
    m.roads.resize(edges.size());
    for (uint32_t r = 0; r < edges.size(); r++) {
        m.roads[r].start.pos[0] = nx[edges[r].a];
        m.roads[r].start.pos[1] = ny[edges[r].a];
        m.roads[r].end.pos[0] = nx[edges[r].b];
        m.roads[r].end.pos[1] = ny[edges[r].b];
    }
    m.connections.clear();
    m.connections.reserve(edges.size() * 3);
    for (uint32_t node = 0; node < nodeCount; node++) {
        if (degree[node] == 0) continue;
        const uint32_t blockStart = static_cast<uint32_t>(m.connections.size());
        for (uint64_t k = adjStart[node]; k < adjStart[node + 1]; k++) {
            const uint32_t r = adj[k];
            m.connections.push_back(r);
            if (edges[r].a == node) {
                m.roads[r].start.connected_index = blockStart;
                m.roads[r].start.connected_count = degree[node];
            } else {
                m.connections.push_back(r);  // generate_map.py:252 appends END-matching roads twice
                m.roads[r].end.connected_index = blockStart;
                m.roads[r].end.connected_count = degree[node];
            }
        }
    }
}

uint32_t uf_find(std::vector<uint32_t>& parent, uint32_t x) {
    while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
    }
    return x;
}
}  // namespace

extern "C" {

const char* msim_map_last_error(void) { return msim_host::map_error().c_str(); }

int msim_map_load_json(const char* path, msim_map** out) {
    if (!path || !out) return map_fail(MSIM_ERR_INVALID, "msim_map_load_json: null argument");
    *out = nullptr;
    std::string text;
    if (!msim_host::read_file(path, text)) return MSIM_ERR_IO;
    msim_map* m = new (std::nothrow) msim_map();
    if (!m) return map_fail(MSIM_ERR_OOM, "out of host memory");
    try {
        JsonCursor js(text.data(), text.data() + text.size());
        parse_map(js, *m);
    } catch (const std::exception& e) {
        delete m;
        return map_fail(MSIM_ERR_PARSE, e.what());
    }
    *out = m;
    return MSIM_OK;
}

int msim_map_save_json(const msim_map* m, const char* path) {
    if (!m || !path) return map_fail(MSIM_ERR_INVALID, "msim_map_save_json: null argument");
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return map_fail(MSIM_ERR_IO, std::string("cannot write '") + path + "'");
    // %.9g round-trips every binary32 value through a double-parsing reader.
    std::fprintf(f, "{\"minDistLat\": 0, \"maxDistLat\": %.9g, \"minDistLong\": 0, \"maxDistLong\": %.9g, \"roads\": [", m->width, m->height);
    for (size_t i = 0; i < m->roads.size(); i++) {
        const msim_road& r = m->roads[i];
        std::fprintf(f,
                     "%s{\"start\": {\"lat\": 0, \"long\": 0, \"distLat\": %.9g, \"distLong\": %.9g}, "
                     "\"end\": {\"lat\": 0, \"long\": 0, \"distLat\": %.9g, \"distLong\": %.9g}, "
                     "\"connIndexStart\": %u, \"connCountStart\": %u, \"connIndexEnd\": %u, \"connCountEnd\": %u}",
                     i ? ", " : "", r.start.pos[0], r.start.pos[1], r.end.pos[0], r.end.pos[1], r.start.connected_index,
                     r.start.connected_count, r.end.connected_index, r.end.connected_count);
    }
    std::fprintf(f, "], \"connectionRoadIndexList\": [");
    for (size_t i = 0; i < m->connections.size(); i++) std::fprintf(f, "%s%u", i ? ", " : "", m->connections[i]);
    std::fprintf(f, "]}\n");
    const bool ok = std::fclose(f) == 0;
    return ok ? MSIM_OK : map_fail(MSIM_ERR_IO, "write failed");
}

int msim_map_generate_city(float world_w, float world_h, float spacing, float jitter, float drop_prob, uint64_t seed, msim_map** out) {
    if (!out || !(world_w > 0) || !(world_h > 0) || !(spacing > 0) || jitter < 0 || jitter >= 0.5f || drop_prob < 0 || drop_prob >= 1)
        return map_fail(MSIM_ERR_INVALID, "msim_map_generate_city: bad parameters");
    *out = nullptr;
    const uint32_t gx = std::max<uint32_t>(2, static_cast<uint32_t>(std::floor(world_w / spacing)) + 1);
    const uint32_t gy = std::max<uint32_t>(2, static_cast<uint32_t>(std::floor(world_h / spacing)) + 1);
    if (static_cast<uint64_t>(gx) * gy > (1ull << 30)) return map_fail(MSIM_ERR_INVALID, "city grid too large");
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<float> uni(0.0f, 1.0f);
    const uint32_t nodeCount = gx * gy;
    std::vector<float> px(nodeCount), py(nodeCount);
    const float stepX = world_w / static_cast<float>(gx - 1), stepY = world_h / static_cast<float>(gy - 1);
    for (uint32_t j = 0; j < gy; j++) {
        for (uint32_t i = 0; i < gx; i++) {
            float x = static_cast<float>(i) * stepX + (uni(gen) - 0.5f) * 2.0f * jitter * stepX;
            float y = static_cast<float>(j) * stepY + (uni(gen) - 0.5f) * 2.0f * jitter * stepY;
            px[j * gx + i] = std::min(std::max(x, 0.0f), world_w);
            py[j * gx + i] = std::min(std::max(y, 0.0f), world_h);
        }
    }
    // candidate edges: right, down, and a sparse set of diagonals (degree 5+ junctions)
    std::vector<Edge> cand;
    cand.reserve(static_cast<size_t>(nodeCount) * 2);
    for (uint32_t j = 0; j < gy; j++) {
        for (uint32_t i = 0; i < gx; i++) {
            const uint32_t n = j * gx + i;
            auto add = [&](uint32_t other) {
                if (px[n] == px[other] && py[n] == py[other]) return;  // never emit zero-length roads
                if (uni(gen) < 0.5f) cand.push_back({n, other});
                else cand.push_back({other, n});
            };
            if (i + 1 < gx && uni(gen) >= drop_prob) add(n + 1);
            if (j + 1 < gy && uni(gen) >= drop_prob) add(n + gx);
            if (i + 1 < gx && j + 1 < gy && uni(gen) < 0.04f) add(n + gx + 1);
        }
    }
    // keep the largest connected component, like remove_not_connected (generate_map.py:184-231)
    std::vector<uint32_t> parent(nodeCount);
    std::iota(parent.begin(), parent.end(), 0u);
    for (const Edge& e : cand) {
        const uint32_t ra = uf_find(parent, e.a), rb = uf_find(parent, e.b);
        if (ra != rb) parent[ra] = rb;
    }
    std::vector<uint32_t> compSize(nodeCount, 0);
    for (const Edge& e : cand) compSize[uf_find(parent, e.a)]++;
    const uint32_t best = static_cast<uint32_t>(std::max_element(compSize.begin(), compSize.end()) - compSize.begin());
    std::vector<Edge> edges;
    edges.reserve(cand.size());
    for (const Edge& e : cand)
        if (uf_find(parent, e.a) == best) edges.push_back(e);
    msim_map* m = new (std::nothrow) msim_map();
    if (!m) return map_fail(MSIM_ERR_OOM, "out of host memory");
    m->width = world_w;
    m->height = world_h;
    emit_tables(px, py, edges, *m);
    *out = m;
    return MSIM_OK;
}

int msim_map_generate_grid(uint32_t nx, uint32_t ny, float spacing, msim_map** out) {
    if (!out || nx < 2 || ny < 2 || !(spacing > 0)) return map_fail(MSIM_ERR_INVALID, "msim_map_generate_grid: bad parameters");
    *out = nullptr;
    const uint64_t nodeCount = static_cast<uint64_t>(nx) * ny;
    const uint64_t edgeCount = static_cast<uint64_t>(nx - 1) * ny + static_cast<uint64_t>(ny - 1) * nx;
    if (nodeCount > 0xFFFFFFF0ull || edgeCount > 0xFFFFFFF0ull) return map_fail(MSIM_ERR_INVALID, "grid too large for 32-bit road indices");
    std::vector<float> px(nodeCount), py(nodeCount);
    for (uint32_t j = 0; j < ny; j++)
        for (uint32_t i = 0; i < nx; i++) {
            px[static_cast<size_t>(j) * nx + i] = spacing * static_cast<float>(i);
            py[static_cast<size_t>(j) * nx + i] = spacing * static_cast<float>(j);
        }
    std::vector<Edge> edges;
    edges.reserve(edgeCount);
    for (uint32_t j = 0; j < ny; j++)
        for (uint32_t i = 0; i < nx; i++) {
            const uint32_t n = j * nx + i;
            if (i + 1 < nx) edges.push_back({n, n + 1});
            if (j + 1 < ny) edges.push_back({n, n + nx});
        }
    msim_map* m = new (std::nothrow) msim_map();
    if (!m) return map_fail(MSIM_ERR_OOM, "out of host memory");
    m->width = spacing * static_cast<float>(nx - 1);
    m->height = spacing * static_cast<float>(ny - 1);
    emit_tables(px, py, edges, *m);
    *out = m;
    return MSIM_OK;
}

void msim_map_free(msim_map* m) { delete m; }
float msim_map_width(const msim_map* m) { return m ? m->width : 0.0f; }
float msim_map_height(const msim_map* m) { return m ? m->height : 0.0f; }
uint64_t msim_map_road_count(const msim_map* m) { return m ? m->roads.size() : 0; }
uint64_t msim_map_connection_count(const msim_map* m) { return m ? m->connections.size() : 0; }
const msim_road* msim_map_roads(const msim_map* m) { return m ? m->roads.data() : nullptr; }
const uint32_t* msim_map_connections(const msim_map* m) { return m ? m->connections.data() : nullptr; }

int msim_entities_init(const msim_road* roads, uint64_t road_count, uint64_t count, uint64_t seed, const float* box, msim_entity* out) {
    if (!roads || road_count == 0 || (!out && count)) return map_fail(MSIM_ERR_INVALID, "msim_entities_init: bad arguments");
    std::vector<uint32_t> pool;
    if (box) {
        for (uint64_t r = 0; r < road_count; r++) {
            auto inside = [&](const float* p) { return p[0] >= box[0] && p[0] <= box[2] && p[1] >= box[1] && p[1] <= box[3]; };
            if (inside(roads[r].start.pos) && inside(roads[r].end.pos)) pool.push_back(static_cast<uint32_t>(r));
        }
        if (pool.empty()) return map_fail(MSIM_ERR_INVALID, "msim_entities_init: no road inside the box");
    }
    const uint64_t choices = box ? pool.size() : road_count;
    std::mt19937 genRoad(static_cast<uint32_t>(seed));
    std::mt19937 genColor(static_cast<uint32_t>(seed + 1));
    std::mt19937 genState(static_cast<uint32_t>(seed + 2));
    std::uniform_int_distribution<unsigned int> pick(0, static_cast<unsigned int>(choices - 1));
    std::uniform_real_distribution<float> channel(0, 1.0);
    std::uniform_int_distribution<unsigned int> word(std::numeric_limits<unsigned int>::min(), std::numeric_limits<unsigned int>::max());
    for (uint64_t i = 0; i < count; i++) {
        const unsigned int chosen = pick(genRoad);
        const uint32_t roadIndex = box ? pool[chosen] : chosen;
        const msim_road& road = roads[roadIndex];
        msim_entity& e = out[i];
        e.color[0] = channel(genColor);
        e.color[1] = channel(genColor);
        e.color[2] = channel(genColor);
        e.color[3] = 1.0f;
        e.rand_state[0] = word(genState);
        e.rand_state[1] = word(genState);
        e.rand_state[2] = word(genState);
        e.rand_state[3] = word(genState);
        e.pos[0] = road.start.pos[0];
        e.pos[1] = road.start.pos[1];
        e.target[0] = road.end.pos[0];
        e.target[1] = road.end.pos[1];
        e.direction[0] = 0.0f;
        e.direction[1] = 0.0f;
        e.road_index = roadIndex;
        e.initialized = 0u;
    }
    return MSIM_OK;
}

// the road-index stream of msim_entities_init alone (same generator, same distribution, same box rule): where every entity of a seeded
// population starts, without materialising the population
int msim_entities_init_roads(const msim_road* roads, uint64_t road_count, uint64_t count, uint64_t seed, const float* box, uint32_t* road_index_out) {
    if (!roads || road_count == 0 || (!road_index_out && count)) return map_fail(MSIM_ERR_INVALID, "msim_entities_init_roads: bad arguments");
    std::vector<uint32_t> pool;
    if (box) {
        for (uint64_t r = 0; r < road_count; r++) {
            auto inside = [&](const float* p) { return p[0] >= box[0] && p[0] <= box[2] && p[1] >= box[1] && p[1] <= box[3]; };
            if (inside(roads[r].start.pos) && inside(roads[r].end.pos)) pool.push_back(static_cast<uint32_t>(r));
        }
        if (pool.empty()) return map_fail(MSIM_ERR_INVALID, "msim_entities_init_roads: no road inside the box");
    }
    const uint64_t choices = box ? pool.size() : road_count;
    std::mt19937 genRoad(static_cast<uint32_t>(seed));
    std::uniform_int_distribution<unsigned int> pick(0, static_cast<unsigned int>(choices - 1));
    for (uint64_t i = 0; i < count; i++) {
        const unsigned int chosen = pick(genRoad);
        road_index_out[i] = box ? pool[chosen] : chosen;
    }
    return MSIM_OK;
}

uint64_t msim_calc_node_count(uint32_t max_depth) {
    uint64_t total = 0, level = 1;
    for (uint32_t d = 0; d < max_depth; d++) {
        total += level;
        level *= 4;
    }
    return total;
}

uint32_t msim_abi_version(void) { return MSIM_ABI_VERSION; }

}  // extern "C"
