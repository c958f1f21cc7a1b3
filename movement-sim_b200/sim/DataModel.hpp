// DataModel.hpp — host PODs of the drop-in sim::Simulator, layout-identical to the reference's
//   sim::Vec2 / Vec4U / Rgba / Entity   (/root/reference/src/sim/Entity.hpp:5-46)
//   sim::Coordinate / Road / RoadPiece / Map   (/root/reference/src/sim/Map.hpp:12-47)
//   sim::gpu_quad_tree::Node / NextType   (/root/reference/src/sim/GpuQuadTree.hpp:8-41)
//   sim::PushConsts   (/root/reference/src/sim/PushConsts.hpp:9-20)
// so that the UI code that consumes them (vertex layouts in src/ui/widgets/opengl/*.cpp) keeps working
// unchanged.  A maintainer integrating into the reference tree keeps the reference's own headers;
// these exist so that this repository builds stand-alone.  Sizes/offsets are static_asserted against
// the C ABI structs (include/msim.h), which are what actually crosses the library boundary.
#pragma once

#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <memory>
#include <optional>
#include <vector>

#include "../../include/msim.h"
#include "../../include/msim_mapgen.h"

namespace sim {

struct Vec2 {
    float x{0};
    float y{0};
};

struct Vec4U {
    unsigned int x{0}, y{0}, z{0}, w{0};
};

struct Rgba {
    float r{0}, g{0}, b{0}, a{0};
};

struct alignas(64) Entity {
    Rgba color{1.0F, 0.0F, 0.0F, 1.0F};
    Vec4U randomState{};
    Vec2 pos{};
    Vec2 target{};
    Vec2 direction{};
    unsigned int roadIndex{0};
    unsigned int initialized{0};
};

struct Coordinate {
    Vec2 pos{};
    unsigned int connectedIndex{0};
    unsigned int connectedCount{0};
};

struct alignas(32) Road {
    Coordinate start;
    Coordinate end;
};

struct alignas(32) RoadPiece {  // render-only (src/ui/widgets/opengl/MapGlObject.cpp)
    Vec2 pos;
    Vec2 padding;
    Rgba color;
};

static_assert(sizeof(Entity) == sizeof(msim_entity) && sizeof(Entity) == 64);
static_assert(offsetof(Entity, color) == offsetof(msim_entity, color));
static_assert(offsetof(Entity, randomState) == offsetof(msim_entity, rand_state));
static_assert(offsetof(Entity, pos) == offsetof(msim_entity, pos) && offsetof(Entity, pos) == 32);
static_assert(offsetof(Entity, target) == offsetof(msim_entity, target));
static_assert(offsetof(Entity, direction) == offsetof(msim_entity, direction));
static_assert(offsetof(Entity, roadIndex) == offsetof(msim_entity, road_index));
static_assert(offsetof(Entity, initialized) == offsetof(msim_entity, initialized));
static_assert(sizeof(Road) == sizeof(msim_road) && sizeof(Road) == 32 && sizeof(Coordinate) == 16);
static_assert(sizeof(RoadPiece) == 32);

class Map {
 public:
    float width{0};
    float height{0};
    std::vector<Road> roads;
    std::vector<RoadPiece> roadPieces;
    std::vector<unsigned int> connections;
    std::optional<size_t> selectedRoad{std::nullopt};

    // nullptr when the file is missing (Map.cpp:30-38); throws std::runtime_error on schema errors (Map.cpp:43-117)
    static std::shared_ptr<Map> load_from_file(const std::filesystem::path& path);
    // stand-ins for the missing munich.json / config 4 (not in the reference)
    static std::shared_ptr<Map> generate_city(float width, float height, uint64_t seed);
    static std::shared_ptr<Map> generate_grid(uint32_t nx, uint32_t ny, float spacing);

    [[nodiscard]] unsigned int get_random_road_index() const;
    void select_road(size_t roadIndex);

 private:
    static std::shared_ptr<Map> adopt(msim_map* handle);
};

namespace gpu_quad_tree {
enum class NextType : uint32_t { INVALID = 0, NODE = 1, ENTITY = 2 };

struct alignas(64) Node {
    int32_t acquireLock{0}, writeLock{0}, readerLock{0};
    float offsetX{0}, offsetY{0}, width{0}, height{0};
    NextType contentType{NextType::INVALID};
    uint32_t entityCount{0};
    uint32_t first{0};
    uint32_t prevNodeIndex{0};
    uint32_t nextTL{0}, nextTR{0}, nextBL{0}, nextBR{0};
    uint32_t padding{0};
};
static_assert(sizeof(Node) == sizeof(msim_quadtree_node) && sizeof(Node) == 64);
static_assert(offsetof(Node, nextTL) == offsetof(msim_quadtree_node, next_tl));

inline size_t calc_node_count(size_t maxDepth) { return static_cast<size_t>(msim_calc_node_count(static_cast<uint32_t>(maxDepth))); }
}  // namespace gpu_quad_tree

#pragma pack(push, 1)
struct PushConsts {
    float worldSizeX{0};
    float worldSizeY{0};
    uint32_t nodeCount{0};
    uint32_t maxDepth{0};
    uint32_t entityNodeCap{0};
    float collisionRadius{0};
    uint32_t tick{0};
};
#pragma pack(pop)
static_assert(sizeof(PushConsts) == sizeof(msim_push_consts) && sizeof(PushConsts) == 28);

}  // namespace sim
