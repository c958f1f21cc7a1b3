// host_mapgen.cpp — map pipeline (include/msim_mapgen.h): GeoJSON -> road graph, binary map cache.  Host only.
//
// The generator restates /root/reference/map/generate_map.py (a script the reference's author runs by hand to
// turn an OpenStreetMap GeoJSON export into munich.json); each step below names the function it follows:
//   build_map                :163-182   LineString features -> one road per consecutive point pair
//   Map.get_min_lat_long     :132-150   reference point = (min lat, min long) over ALL pieces
//   Coordinate.calc_dist     :28-36     metres north / east of the reference point (haversine package)
//   remove_not_connected     :184-231   keep what a traversal from the first road's END reaches; every
//                                       discovered road is re-oriented to START at the vertex it was found from
//   build_road_connections   :233-258   one block per coordinate in first-seen order; END-matching roads twice
//   Map.update_min_max_dist  :107-130   world size
// Identity of a coordinate is exact equality of both binary64 numbers (Coordinate.__eq__, :38-39).
// Python set iteration order (the order of the pieces going into the traversal, the order of the roads inside
// one block) is arbitrary in the script; here it is file order resp. discovery order.  tests/test_mapgen.py
// runs the script itself (from /root/reference, with stand-ins for its two missing imports) to produce the
// golden structure this file is compared with, block contents as multisets.
#include "../../include/msim_mapgen.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <unordered_set>

#include "host_map.h"

namespace {
using msim_host::JsonCursor;
using msim_host::map_fail;

// ---- haversine (PyPI `haversine`, restated: same formula, same operation order, binary64) ----------------
constexpr double AVG_EARTH_RADIUS_KM = 6371.0088;
constexpr double DEG_TO_RAD = 3.14159265358979323846 / 180.0;  // CPython math.radians: x * (pi / 180)

double haversine_km(double lat1, double lng1, double lat2, double lng2) {
    lat1 *= DEG_TO_RAD;
    lng1 *= DEG_TO_RAD;
    lat2 *= DEG_TO_RAD;
    lng2 *= DEG_TO_RAD;
    const double lat = lat2 - lat1, lng = lng2 - lng1;
    // Python's `x ** 2` on floats is C pow(x, 2.0)
    const double d = std::pow(std::sin(lat * 0.5), 2.0) + std::cos(lat1) * std::cos(lat2) * std::pow(std::sin(lng * 0.5), 2.0);
    return 2.0 * AVG_EARTH_RADIUS_KM * std::asin(std::sqrt(d));
}

// ---- interned coordinates --------------------------------------------------------------------------------
struct CoordKey {
    uint64_t lat, lng;
    bool operator==(const CoordKey& o) const { return lat == o.lat && lng == o.lng; }
};
struct CoordKeyHash {
    size_t operator()(const CoordKey& k) const {
        uint64_t h = k.lat * 0x9E3779B97F4A7C15ull;
        h ^= (k.lng + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
        return static_cast<size_t>(h ^ (h >> 29));
    }
};
uint64_t bits_of(double v) {
    v += 0.0;  // -0.0 == 0.0 in the script's comparisons
    uint64_t u;
    std::memcpy(&u, &v, sizeof(u));
    return u;
}

struct Piece {
    uint32_t a, b;  // coordinate ids: start, end
};

struct Graph {
    std::vector<double> lat, lng;  // per coordinate id
    std::unordered_map<CoordKey, uint32_t, CoordKeyHash> ids;
    std::vector<Piece> pieces;     // file order

    uint32_t intern(double la, double lo) {
        const CoordKey k{bits_of(la), bits_of(lo)};
        auto it = ids.find(k);
        if (it != ids.end()) return it->second;
        const uint32_t id = static_cast<uint32_t>(lat.size());
        ids.emplace(k, id);
        lat.push_back(la);
        lng.push_back(lo);
        return id;
    }
};

// ---- build_map (:163-182) --------------------------------------------------------------------------------
// `span` is the raw text of a geometry's "coordinates" value; for a LineString that is [[a, b(, c)], ...]
void add_line_string(const char* begin, const char* end, Graph& g, msim_mapgen_stats& st) {
    JsonCursor js(begin, end);
    std::vector<uint32_t> pts;
    js.expect('[');
    if (!js.consume(']')) {
        do {
            js.expect('[');
            const double first = js.number();
            js.expect(',');
            const double second = js.number();
            while (js.consume(',')) js.skip_value();  // altitude etc.
            js.expect(']');
            pts.push_back(g.intern(first, second));  // :175 takes position[0] as lat, position[1] as long
        } while (js.consume(','));
        js.expect(']');
    }
    if (pts.size() < 2) return;  // "Found road with only one point. Ignoring." (:171-173)
    for (size_t i = 1; i < pts.size(); i++) {
        st.road_pieces++;
        if (pts[i - 1] == pts[i]) {  // "Skipping road, where start == end" (:177-179)
            st.skipped_zero++;
            continue;
        }
        g.pieces.push_back({pts[i - 1], pts[i]});
    }
}

void parse_geometry(JsonCursor& js, Graph& g, msim_mapgen_stats& st) {
    if (js.peek() != '{') {  // "geometry": null
        js.skip_value();
        return;
    }
    std::string type;
    const char* cBegin = nullptr;
    const char* cEnd = nullptr;
    js.expect('{');
    if (!js.consume('}')) {
        do {
            const std::string key = js.string();
            js.expect(':');
            if (key == "type" && js.peek() == '"') {
                type = js.string();
            } else if (key == "coordinates") {
                js.ws();
                cBegin = js.pos();
                js.skip_value();
                cEnd = js.pos();
            } else {
                js.skip_value();
            }
        } while (js.consume(','));
        js.expect('}');
    }
    if (type != "LineString") return;  // :168-169
    st.line_strings++;
    if (!cBegin) throw std::runtime_error("Failed to parse GeoJSON. LineString without 'coordinates'.");
    add_line_string(cBegin, cEnd, g, st);
}

void parse_geojson(JsonCursor& js, Graph& g, msim_mapgen_stats& st) {
    bool hasFeatures = false;
    js.expect('{');
    if (!js.consume('}')) {
        do {
            const std::string key = js.string();
            js.expect(':');
            if (key == "features") {
                hasFeatures = true;
                js.expect('[');
                if (!js.consume(']')) {
                    do {
                        st.features++;
                        bool hasGeometry = false;
                        js.expect('{');
                        if (!js.consume('}')) {
                            do {
                                const std::string fkey = js.string();
                                js.expect(':');
                                if (fkey == "geometry") {
                                    hasGeometry = true;
                                    parse_geometry(js, g, st);
                                } else {
                                    js.skip_value();
                                }
                            } while (js.consume(','));
                            js.expect('}');
                        }
                        if (!hasGeometry) throw std::runtime_error("Failed to parse GeoJSON. Feature without 'geometry'.");  // KeyError at :167
                    } while (js.consume(','));
                    js.expect(']');
                }
            } else {
                js.skip_value();
            }
        } while (js.consume(','));
        js.expect('}');
    }
    if (!hasFeatures) throw std::runtime_error("Failed to parse GeoJSON. 'features' field missing.");  // KeyError at :269
}

// the script asserts that no piece occurs twice (:180); here later copies are dropped and counted
void drop_duplicates(Graph& g, msim_mapgen_stats& st) {
    std::unordered_set<uint64_t> seen;
    seen.reserve(g.pieces.size() * 2);
    size_t w = 0;
    for (const Piece& p : g.pieces) {
        if (!seen.insert((static_cast<uint64_t>(p.a) << 32) | p.b).second) {
            st.skipped_duplicate++;
            continue;
        }
        g.pieces[w++] = p;
    }
    g.pieces.resize(w);
}

// ---- remove_not_connected (:184-231) ---------------------------------------------------------------------
// Returns the kept pieces in discovery order, re-oriented; `firstSeen` receives the coordinates in the order
// the script inserts them into connectionsMap (per popped road: start, then end).
std::vector<Piece> traverse_exact(std::vector<Piece> obj, std::vector<uint32_t>& firstSeen, size_t coordCount) {
    // literal emulation of the Python list the script deletes from while iterating it
    std::vector<uint32_t> list(obj.size());
    for (uint32_t i = 0; i < obj.size(); i++) list[i] = i;
    std::vector<uint32_t> next;
    std::vector<Piece> result;
    std::vector<char> seen(coordCount, 0);
    auto remove_first_equal = [&](const Piece& v) {  // list.remove(x): first element that compares equal
        for (size_t k = 0; k < list.size(); k++) {
            if (obj[list[k]].a == v.a && obj[list[k]].b == v.b) {
                list.erase(list.begin() + static_cast<std::ptrdiff_t>(k));
                return;
            }
        }
    };
    next.push_back(list[0]);
    list.erase(list.begin());
    while (!next.empty()) {
        const uint32_t cur = next.back();
        next.pop_back();
        const Piece c = obj[cur];
        if (!seen[c.a]) { seen[c.a] = 1; firstSeen.push_back(c.a); }
        if (!seen[c.b]) { seen[c.b] = 1; firstSeen.push_back(c.b); }
        result.push_back(c);
        for (size_t i = 0; i < list.size(); i++) {  // `for road in roads:` — the iterator is an index
            Piece& r = obj[list[i]];
            if (c.b == r.a) {
                next.push_back(list[i]);
                remove_first_equal(r);
            } else if (c.b == r.b) {
                r.b = r.a;
                r.a = c.b;
                next.push_back(list[i]);
                remove_first_equal(r);
            }
        }
    }
    return result;
}

std::vector<Piece> traverse_fast(const std::vector<Piece>& pieces, std::vector<uint32_t>& firstSeen, size_t coordCount) {
    // per-coordinate lists of touching pieces in file order (CSR)
    std::vector<uint64_t> start(coordCount + 1, 0);
    for (const Piece& p : pieces) {
        start[p.a + 1]++;
        start[p.b + 1]++;
    }
    for (size_t i = 0; i < coordCount; i++) start[i + 1] += start[i];
    std::vector<uint32_t> adj(start[coordCount]);
    std::vector<uint64_t> fill(start.begin(), start.end() - 1);
    for (uint32_t i = 0; i < pieces.size(); i++) {
        adj[fill[pieces[i].a]++] = i;
        adj[fill[pieces[i].b]++] = i;
    }
    std::vector<char> found(pieces.size(), 0), seen(coordCount, 0), expanded(coordCount, 0);
    std::vector<Piece> result, oriented(pieces);
    std::vector<uint32_t> next;
    next.push_back(0);
    found[0] = 1;
    while (!next.empty()) {
        const uint32_t cur = next.back();
        next.pop_back();
        const Piece c = oriented[cur];
        if (!seen[c.a]) { seen[c.a] = 1; firstSeen.push_back(c.a); }
        if (!seen[c.b]) { seen[c.b] = 1; firstSeen.push_back(c.b); }
        result.push_back(c);
        if (expanded[c.b]) continue;  // everything touching this vertex has been discovered already
        expanded[c.b] = 1;
        for (uint64_t k = start[c.b]; k < start[c.b + 1]; k++) {
            const uint32_t r = adj[k];
            if (found[r]) continue;
            found[r] = 1;
            if (oriented[r].a != c.b) {  // found from its end: turn it around (:216-222)
                oriented[r].b = oriented[r].a;
                oriented[r].a = c.b;
            }
            next.push_back(r);
        }
    }
    return result;
}

// ---- the whole script ------------------------------------------------------------------------------------
void generate(Graph& g, uint32_t flags, msim_map& m, msim_mapgen_stats& st) {
    drop_duplicates(g, st);
    if (g.pieces.empty()) throw std::runtime_error("Failed to build map. The GeoJSON holds no road.");  // IndexError at :188
    // get_min_lat_long (:132-150) over every piece, connected or not
    double refLat = DBL_MAX, refLong = DBL_MAX;
    for (const Piece& p : g.pieces) {
        refLat = std::min({refLat, g.lat[p.a], g.lat[p.b]});
        refLong = std::min({refLong, g.lng[p.a], g.lng[p.b]});
    }
    st.ref_lat = refLat;
    st.ref_long = refLong;

    std::vector<uint32_t> firstSeen;
    const std::vector<Piece> kept = (flags & MSIM_MAPGEN_EXACT_TRAVERSAL) ? traverse_exact(g.pieces, firstSeen, g.lat.size())
                                                                         : traverse_fast(g.pieces, firstSeen, g.lat.size());
    if (kept.size() > 0xFFFFFFF0ull) throw std::runtime_error("Failed to build map. More than 2^32 roads.");
    st.connected = kept.size();
    st.coordinates = firstSeen.size();

    // calc_dist (:34-36): metres along the meridian / along the equator from the reference point
    std::vector<double> distLat(g.lat.size(), 0.0), distLong(g.lat.size(), 0.0);
    for (const uint32_t c : firstSeen) {
        distLat[c] = haversine_km(refLat, 0.0, g.lat[c], 0.0) * 1000.0;
        distLong[c] = haversine_km(0.0, refLong, 0.0, g.lng[c]) * 1000.0;
    }
    // update_min_max_dist (:107-130)
    st.min_dist_lat = st.min_dist_long = DBL_MAX;
    st.max_dist_lat = st.max_dist_long = 0.0;
    for (const Piece& p : kept) {
        for (const uint32_t c : {p.a, p.b}) {
            st.max_dist_lat = std::max(st.max_dist_lat, distLat[c]);
            st.min_dist_lat = std::min(st.min_dist_lat, distLat[c]);
            st.max_dist_long = std::max(st.max_dist_long, distLong[c]);
            st.min_dist_long = std::min(st.min_dist_long, distLong[c]);
        }
    }

    // build_road_connections (:233-258); road index = discovery order (:236-237)
    m.width = static_cast<float>(st.max_dist_lat);    // Map.cpp:45-46
    m.height = static_cast<float>(st.max_dist_long);  // Map.cpp:51-52
    m.roads.assign(kept.size(), msim_road{});
    std::vector<uint64_t> blockStart(g.lat.size() + 1, 0);
    for (const Piece& p : kept) {
        blockStart[p.a + 1]++;
        blockStart[p.b + 1]++;
    }
    for (size_t i = 0; i < g.lat.size(); i++) blockStart[i + 1] += blockStart[i];
    std::vector<uint32_t> members(blockStart[g.lat.size()]);
    std::vector<uint64_t> fill(blockStart.begin(), blockStart.end() - 1);
    for (uint32_t r = 0; r < kept.size(); r++) {  // the order roads join a coordinate's set: pop order, start before end (:198-205)
        members[fill[kept[r].a]++] = r;
        members[fill[kept[r].b]++] = r;
        msim_road& road = m.roads[r];
        road.start.pos[0] = static_cast<float>(distLat[kept[r].a]);  // Map.cpp:121-122: x = distLat, y = distLong
        road.start.pos[1] = static_cast<float>(distLong[kept[r].a]);
        road.end.pos[0] = static_cast<float>(distLat[kept[r].b]);
        road.end.pos[1] = static_cast<float>(distLong[kept[r].b]);
    }
    const bool duplicateEnd = !(flags & MSIM_MAPGEN_NO_DUPLICATE_END);
    m.connections.clear();
    m.connections.reserve(kept.size() * 3);
    for (const uint32_t c : firstSeen) {
        if (m.connections.size() > 0xFFFFFFF0ull) throw std::runtime_error("Failed to build map. Connection table exceeds 2^32 entries.");
        const uint32_t connIndex = static_cast<uint32_t>(m.connections.size());
        const uint32_t count = static_cast<uint32_t>(blockStart[c + 1] - blockStart[c]);
        for (uint64_t k = blockStart[c]; k < blockStart[c + 1]; k++) {
            const uint32_t r = members[k];
            m.connections.push_back(r);
            if (kept[r].a == c) {
                m.roads[r].start.connected_index = connIndex;
                m.roads[r].start.connected_count = count;
            } else {
                if (duplicateEnd) m.connections.push_back(r);  // :252
                m.roads[r].end.connected_index = connIndex;
                m.roads[r].end.connected_count = count;
            }
        }
    }
}

// ---- binary cache ----------------------------------------------------------------------------------------
constexpr char MAGIC[8] = {'M', 'S', 'I', 'M', 'M', 'A', 'P', '1'};
struct CacheHeader {
    char magic[8];
    uint32_t version;
    uint32_t byte_order;  // 0x01020304 as written by the producer
    float width, height;
    uint64_t road_count;
    uint64_t connection_count;
};
static_assert(sizeof(CacheHeader) == 40, "cache header layout");

uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    // 8 bytes per step (word-wise FNV-1a variant): the lattice of config 4 is 2 GB of tables
    const unsigned char* p = static_cast<const unsigned char*>(data);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001B3ull;
    }
    for (; i < bytes; i++) h = (h ^ p[i]) * 0x100000001B3ull;
    return h;
}
uint64_t map_checksum(const CacheHeader& hd, const msim_map& m) {
    uint64_t h = 0xCBF29CE484222325ull;
    h = fnv1a(&hd, sizeof(hd), h);
    h = fnv1a(m.roads.data(), m.roads.size() * sizeof(msim_road), h);
    h = fnv1a(m.connections.data(), m.connections.size() * sizeof(uint32_t), h);
    return h;
}

bool has_suffix(const char* path, const char* suffix) {
    const size_t n = std::strlen(path), k = std::strlen(suffix);
    return n >= k && std::strcmp(path + n - k, suffix) == 0;
}
}  // namespace

extern "C" {

double msim_haversine_m(double lat1, double lng1, double lat2, double lng2) { return haversine_km(lat1, lng1, lat2, lng2) * 1000.0; }

int msim_map_from_geojson(const char* path, uint32_t flags, msim_map** out, msim_mapgen_stats* stats) {
    if (!path || !out) return map_fail(MSIM_ERR_INVALID, "msim_map_from_geojson: null argument");
    *out = nullptr;
    if (flags & ~(MSIM_MAPGEN_EXACT_TRAVERSAL | MSIM_MAPGEN_NO_DUPLICATE_END)) return map_fail(MSIM_ERR_INVALID, "msim_map_from_geojson: unknown flag");
    std::string text;
    if (!msim_host::read_file(path, text)) return MSIM_ERR_IO;
    msim_map* m = new (std::nothrow) msim_map();
    if (!m) return map_fail(MSIM_ERR_OOM, "out of host memory");
    msim_mapgen_stats st{};
    try {
        Graph g;
        JsonCursor js(text.data(), text.data() + text.size());
        parse_geojson(js, g, st);
        text.clear();
        text.shrink_to_fit();
        generate(g, flags, *m, st);
    } catch (const std::bad_alloc&) {
        delete m;
        return map_fail(MSIM_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        delete m;
        return map_fail(MSIM_ERR_PARSE, e.what());
    }
    if (stats) *stats = st;
    *out = m;
    return MSIM_OK;
}

int msim_map_save_binary(const msim_map* m, const char* path) {
    if (!m || !path) return map_fail(MSIM_ERR_INVALID, "msim_map_save_binary: null argument");
    CacheHeader hd{};
    std::memcpy(hd.magic, MAGIC, 8);
    hd.version = 1;
    hd.byte_order = 0x01020304u;
    hd.width = m->width;
    hd.height = m->height;
    hd.road_count = m->roads.size();
    hd.connection_count = m->connections.size();
    const uint64_t sum = map_checksum(hd, *m);
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return map_fail(MSIM_ERR_IO, std::string("cannot write '") + path + "'");
    bool ok = std::fwrite(&hd, sizeof(hd), 1, f) == 1;
    ok = ok && (m->roads.empty() || std::fwrite(m->roads.data(), sizeof(msim_road), m->roads.size(), f) == m->roads.size());
    ok = ok && (m->connections.empty() || std::fwrite(m->connections.data(), sizeof(uint32_t), m->connections.size(), f) == m->connections.size());
    ok = ok && std::fwrite(&sum, sizeof(sum), 1, f) == 1;
    ok = (std::fclose(f) == 0) && ok;
    return ok ? MSIM_OK : map_fail(MSIM_ERR_IO, std::string("write to '") + path + "' failed");
}

int msim_map_load_binary(const char* path, msim_map** out) {
    if (!path || !out) return map_fail(MSIM_ERR_INVALID, "msim_map_load_binary: null argument");
    *out = nullptr;
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return map_fail(MSIM_ERR_IO, std::string("Failed to open map from '") + path + "'. File does not exist.");
    CacheHeader hd{};
    msim_map* m = nullptr;
    int rc = MSIM_OK;
    std::string why;
    do {
        if (std::fread(&hd, sizeof(hd), 1, f) != 1 || std::memcmp(hd.magic, MAGIC, 8) != 0) { rc = MSIM_ERR_PARSE; why = "not a map cache file (bad magic)"; break; }
        if (hd.byte_order != 0x01020304u) { rc = MSIM_ERR_UNSUPPORTED; why = "map cache written with another byte order"; break; }
        if (hd.version != 1) { rc = MSIM_ERR_UNSUPPORTED; why = "map cache version " + std::to_string(hd.version) + " is not supported"; break; }
        // the counts must agree with the file size before anything is allocated
        if (std::fseek(f, 0, SEEK_END) != 0) { rc = MSIM_ERR_IO; why = "seek failed"; break; }
        const long long size = std::ftell(f);
        const unsigned long long need = sizeof(hd) + 8ull;
        if (hd.road_count > 0xFFFFFFFFull || hd.connection_count > 0xFFFFFFFFull ||
            size < 0 || static_cast<unsigned long long>(size) != need + hd.road_count * sizeof(msim_road) + hd.connection_count * sizeof(uint32_t)) {
            rc = MSIM_ERR_PARSE; why = "map cache is truncated or its header is corrupt"; break;
        }
        std::fseek(f, static_cast<long>(sizeof(hd)), SEEK_SET);
        m = new (std::nothrow) msim_map();
        if (!m) { rc = MSIM_ERR_OOM; why = "out of host memory"; break; }
        try {
            m->roads.resize(hd.road_count);
            m->connections.resize(hd.connection_count);
        } catch (const std::bad_alloc&) { rc = MSIM_ERR_OOM; why = "out of host memory"; break; }
        m->width = hd.width;
        m->height = hd.height;
        uint64_t sum = 0;
        bool ok = m->roads.empty() || std::fread(m->roads.data(), sizeof(msim_road), m->roads.size(), f) == m->roads.size();
        ok = ok && (m->connections.empty() || std::fread(m->connections.data(), sizeof(uint32_t), m->connections.size(), f) == m->connections.size());
        ok = ok && std::fread(&sum, sizeof(sum), 1, f) == 1;
        if (!ok) { rc = MSIM_ERR_IO; why = "read failed"; break; }
        if (sum != map_checksum(hd, *m)) { rc = MSIM_ERR_PARSE; why = "map cache checksum mismatch"; break; }
    } while (false);
    std::fclose(f);
    if (rc != MSIM_OK) {
        delete m;
        return map_fail(rc, std::string("Failed to load map cache '") + path + "': " + why);
    }
    *out = m;
    return MSIM_OK;
}

int msim_map_from_arrays(float width, float height, const msim_road* roads, uint64_t road_count, const uint32_t* connections,
                         uint64_t connection_count, msim_map** out) {
    if (!out || (road_count && !roads) || (connection_count && !connections)) return map_fail(MSIM_ERR_INVALID, "msim_map_from_arrays: null argument");
    *out = nullptr;
    msim_map* m = new (std::nothrow) msim_map();
    if (!m) return map_fail(MSIM_ERR_OOM, "out of host memory");
    try {
        m->roads.assign(roads, roads + road_count);
        m->connections.assign(connections, connections + connection_count);
    } catch (const std::bad_alloc&) {
        delete m;
        return map_fail(MSIM_ERR_OOM, "out of host memory");
    }
    m->width = width;
    m->height = height;
    *out = m;
    return MSIM_OK;
}

int msim_map_load(const char* path, msim_map** out) {
    if (!path || !out) return map_fail(MSIM_ERR_INVALID, "msim_map_load: null argument");
    *out = nullptr;
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return map_fail(MSIM_ERR_IO, std::string("Failed to open map from '") + path + "'. File does not exist.");
    char head[8] = {0};
    const size_t got = std::fread(head, 1, sizeof(head), f);
    std::fclose(f);
    if (got == sizeof(head) && std::memcmp(head, MAGIC, 8) == 0) return msim_map_load_binary(path, out);
    if (has_suffix(path, ".geojson")) return msim_map_from_geojson(path, 0u, out, nullptr);
    return msim_map_load_json(path, out);
}

}  // extern "C"
