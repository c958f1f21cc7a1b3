/*
 * msim_mapgen.h — map pipeline of libmsim_cuda.so (host only, no GPU needed): the data format on the
 * input side of the hot path.
 *
 *   GeoJSON (OpenStreetMap export) --msim_map_from_geojson--> msim_map --msim_map_save_binary--> cache file
 *                                                                ^------msim_map_load-------------'
 *
 * The reference builds its map with a Python script (map/generate_map.py, run by hand, writes munich.json)
 * and parses that JSON at every start (src/sim/Map.cpp:28-150).  Here both steps are native: the generator
 * follows generate_map.py function by function (cited below), and a binary cache replaces JSON parsing for
 * large graphs (the 33.5 M-road lattice of BASELINE config 4 is 1 GB of road records).
 *
 * Third-party arithmetic the reference script pulls in and that is NOT under the reference checkout:
 * the PyPI package `haversine` (version unpinned, map/README.md) — great-circle distance
 *     d = sin^2(dlat/2) + cos(lat1) cos(lat2) sin^2(dlng/2);   km = 2 * 6371.0088 * asin(sqrt(d))
 * restated in host_mapgen.cpp with the same operation order in binary64.
 */
#ifndef MSIM_MAPGEN_H
#define MSIM_MAPGEN_H

#include "msim.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    /* remove_not_connected (map/generate_map.py:184-231) walks the graph while deleting from the Python list it
     * iterates, so the element after every removed road is skipped in that pass and some connected roads are
     * dropped depending on list order.  Default: a plain traversal that keeps every road reachable from the
     * first road's end point (same orientation rule, same LIFO order).  With this flag the list mutation is
     * emulated literally (O(roads^2), like the script): road order in == road order the script would see. */
    MSIM_MAPGEN_EXACT_TRAVERSAL = 1u << 0,
    /* build_road_connections (map/generate_map.py:244-258) appends a road twice to a coordinate's block when the
     * coordinate is the road's END.  Default: replicate.  With this flag every road appears once. */
    MSIM_MAPGEN_NO_DUPLICATE_END = 1u << 1
};

typedef struct msim_mapgen_stats {
    uint64_t features;         /* GeoJSON features seen */
    uint64_t line_strings;     /* of which LineString geometries (everything else is ignored, :168-169) */
    uint64_t road_pieces;      /* consecutive point pairs turned into roads (:175-181) */
    uint64_t skipped_zero;     /* pieces with start == end (:177-179) */
    uint64_t skipped_duplicate;/* pieces equal to an earlier one (the script asserts there are none, :180) */
    uint64_t connected;        /* roads kept by remove_not_connected */
    uint64_t coordinates;      /* distinct junction coordinates = blocks of the connection table */
    double min_dist_lat, max_dist_lat, min_dist_long, max_dist_long; /* update_min_max_dist (:107-130) */
    double ref_lat, ref_long;  /* get_min_lat_long (:132-150) */
} msim_mapgen_stats;

/* map/generate_map.py end to end: build_map (:163-182), get_min_lat_long (:132-150), remove_not_connected
 * (:184-231), build_road_connections (:233-258), update_min_max_dist (:107-130).  The first number of every
 * GeoJSON position is taken as "lat" and the second as "long", exactly as the script does (:175).
 * Road i of the result is the i-th road the traversal discovered — the index build_road_connections assigns
 * (:236-237) — so connection entries and road records agree.  (The script itself then writes its roads in
 * Python set order, :158, which scrambles that correspondence; that accident is not reproduced.)
 * Positions: x = distLat, y = distLong narrowed to binary32 like Map.cpp:121-122; width/height = maxDistLat /
 * maxDistLong (Map.cpp:45-52).  stats may be NULL. */
int msim_map_from_geojson(const char* path, uint32_t flags, msim_map** out, msim_mapgen_stats* stats);

/* Binary map cache: little-endian, "MSIMMAP1" magic, counts, 32-byte road records and u32 connection entries
 * verbatim, FNV-1a checksum.  Loads with two freads instead of a JSON parse. */
int msim_map_save_binary(const msim_map* m, const char* path);
int msim_map_load_binary(const char* path, msim_map** out);
/* Any supported file: binary cache (by magic), GeoJSON (by ".geojson" suffix, default flags), else the
 * reference's map JSON (msim_map_load_json). */
int msim_map_load(const char* path, msim_map** out);

/* A map object from caller-owned tables (copied): lets any producer use the cache writer / JSON writer. */
int msim_map_from_arrays(float width, float height, const msim_road* roads, uint64_t road_count, const uint32_t* connections,
                         uint64_t connection_count, msim_map** out);

/* Great-circle distance in metres as generate_map.py uses it (haversine(...) * 1000, :28-36). */
double msim_haversine_m(double lat1, double lng1, double lat2, double lng2);

#ifdef __cplusplus
}
#endif
#endif /* MSIM_MAPGEN_H */
