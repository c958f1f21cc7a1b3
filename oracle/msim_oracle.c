/*
 * msim_oracle.c — CPU ORACLE (test infrastructure, NOT product code; see msim_oracle.h header).
 *
 * Restates, in plain C with IEEE binary32 arithmetic, the per-entity semantics of the reference
 * compute shader /root/reference/src/sim/shader/random_move.comp.  Every function cites the lines
 * it follows.  Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 *
 * Parity pin: both halves are pinned against the reference's own code compiled into oracle/_ref — movement / RNG / road
 * choice against the shader's own text (libref_shader_move.so), collision flags against the reference's CPU quadtree
 * (libref_quadtree.so) — see msim_oracle.h.
 */
#include "msim_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

/* random_move.comp:750 */
static const float ORC_SPEED = 1.4f;

/* ------------------------------------------------------------------------------------------ */
/* RNG — random_move.comp:725-746                                                              */
/* ------------------------------------------------------------------------------------------ */

/* random_move.comp:725-736.  state = (x,y,z,w) = s[0..3]. */
uint32_t orc_xorshift128(uint32_t s[4]) {
    uint32_t t = s[3];
    const uint32_t x = s[0];
    s[3] = s[2];
    s[2] = s[1];
    s[1] = x;
    t ^= t << 11;
    t ^= t >> 8;
    s[0] = t ^ x ^ (x >> 19);
    return s[0];
}

/* random_move.comp:738-741.  uint->float is round-to-nearest-even (C cast on x86-64 is RNE);
 * 1.0/4294967296.0 = 2^-32 is exact, so the product is exact.  May return exactly 1.0f. */
float orc_next_float(uint32_t s[4]) {
    const uint32_t u = orc_xorshift128(s);
    const float f = (float)u;
    return f * 2.3283064365386962890625e-10f;
}

/* random_move.comp:743-746: uint(ceil(float(min) + (next_float * float(max - min + 1)))) - 1.
 * The multiply and the add are two separately rounded binary32 operations (no FMA). */
uint32_t orc_next_range(uint32_t s[4], uint32_t lo, uint32_t hi) {
    const float span = (float)(hi - lo + 1u);
    volatile float prod = orc_next_float(s) * span; /* volatile: forbid contraction whatever the flags */
    const float sum = (float)lo + prod;
    return (uint32_t)ceilf(sum) - 1u;
}

/* ------------------------------------------------------------------------------------------ */
/* Movement — random_move.comp:778-852                                                         */
/* ------------------------------------------------------------------------------------------ */

static inline float orc_length2(float dx, float dy) {
    /* GLSL length(): sqrt(dx*dx + dy*dy); products and sum individually rounded */
    volatile float xx = dx * dx;
    volatile float yy = dy * dy;
    const float s = xx + yy;
    return sqrtf(s);
}

/* random_move.comp:830-839 */
static inline void orc_update_direction(orc_entity* e, float px, float py) {
    const float dx = e->target[0] - px;
    const float dy = e->target[1] - py;
    const float len = orc_length2(dx, dy);
    if (len == 0.0f) {
        e->dir[0] = 0.0f;
        e->dir[1] = 0.0f;
        return;
    }
    const float nx = dx / len;
    const float ny = dy / len;
    e->dir[0] = nx * ORC_SPEED;
    e->dir[1] = ny * ORC_SPEED;
}

/* connections[] read with the canonical out-of-bounds rule of SURVEY App. B1: the table behaves as
 * if padded with zeros, i.e. any index past the end yields road 0. */
static inline uint32_t orc_conn(const orc_map* m, uint64_t idx, orc_move_stats* st) {
    if (idx >= m->connection_count) {
        if (st) st->oob_reads++;
        return 0u;
    }
    return m->connections[idx];
}

/* random_move.comp:778-828 */
static inline void orc_new_target(orc_entity* e, const orc_map* m, orc_move_stats* st) {
    const orc_road* cur = &m->roads[e->road];
    const orc_coord* c;
    if (e->target[0] == cur->start.pos[0] && e->target[1] == cur->start.pos[1]) {
        c = &cur->start;
        if (c->conn_count <= 1u) { /* :786-789 turn around */
            e->target[0] = cur->end.pos[0];
            e->target[1] = cur->end.pos[1];
            if (st) st->uturns++;
            return;
        }
    } else {
        c = &cur->end;
        if (c->conn_count <= 1u) { /* :794-797 */
            e->target[0] = cur->start.pos[0];
            e->target[1] = cur->start.pos[1];
            if (st) st->uturns++;
            return;
        }
    }

    uint32_t next_road;
    if (c->conn_count == 2u) { /* :802-804 */
        next_road = orc_conn(m, (uint64_t)c->conn_index + 1u, st);
    } else { /* :805-810 */
        const uint32_t off = orc_next_range(e->rng, 1u, c->conn_count);
        if (st) st->rng_draws++;
        next_road = orc_conn(m, (uint64_t)c->conn_index + (uint64_t)off, st);
    }

    const orc_road* nr = &m->roads[next_road]; /* :813-820 */
    if (nr->start.pos[0] == e->target[0] && nr->start.pos[1] == e->target[1]) {
        e->target[0] = nr->end.pos[0];
        e->target[1] = nr->end.pos[1];
    } else {
        e->target[0] = nr->start.pos[0];
        e->target[1] = nr->start.pos[1];
    }
    e->road = next_road;
}

/* update_direction(index, pos) followed by move(index) and the position store of
 * quad_tree_update — random_move.comp:870-873, :841-852, :499-503/:527. */
static inline void orc_move_one(orc_entity* e, const orc_map* m, orc_move_stats* st) {
    orc_update_direction(e, e->pos[0], e->pos[1]);
    const float ex = e->pos[0] - e->target[0];
    const float ey = e->pos[1] - e->target[1];
    const float dist = orc_length2(ex, ey); /* GLSL distance(pos, target) */
    if (dist > ORC_SPEED) {
        e->pos[0] = e->pos[0] + e->dir[0];
        e->pos[1] = e->pos[1] + e->dir[1];
        return;
    }
    const float nx = e->target[0];
    const float ny = e->target[1];
    if (st) st->arrivals++;
    orc_new_target(e, m, st);
    orc_update_direction(e, nx, ny);
    e->pos[0] = nx;
    e->pos[1] = ny;
}

static void orc_move_range(orc_entity* e, size_t lo, size_t hi, const orc_map* m, orc_move_stats* st) {
    for (size_t i = lo; i < hi; i++) {
        if (e[i].initialized == 0u) { /* random_move.comp:863-867 */
            e[i].initialized = 1u;
            if (st) st->initialised++;
            continue;
        }
        if (st) st->moved++;
        orc_move_one(&e[i], m, st);
    }
}

void orc_move_pass(orc_entity* e, size_t n, const orc_map* map, orc_move_stats* stats) {
    orc_move_range(e, 0, n, map, stats);
}

typedef struct {
    orc_entity* e;
    size_t lo, hi;
    const orc_map* map;
    orc_move_stats st;
} orc_move_job;

static void* orc_move_thread(void* p) {
    orc_move_job* j = (orc_move_job*)p;
    orc_move_range(j->e, j->lo, j->hi, j->map, &j->st);
    return NULL;
}

void orc_move_pass_mt(orc_entity* e, size_t n, const orc_map* map, int threads, orc_move_stats* stats) {
    if (threads < 1) threads = 1;
    orc_move_job* jobs = (orc_move_job*)calloc((size_t)threads, sizeof(orc_move_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].e = e;
        jobs[t].lo = n * (size_t)t / (size_t)threads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        jobs[t].map = map;
        pthread_create(&tid[t], NULL, orc_move_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) {
        pthread_join(tid[t], NULL);
        if (stats) {
            stats->moved += jobs[t].st.moved;
            stats->arrivals += jobs[t].st.arrivals;
            stats->rng_draws += jobs[t].st.rng_draws;
            stats->uturns += jobs[t].st.uturns;
            stats->oob_reads += jobs[t].st.oob_reads;
            stats->initialised += jobs[t].st.initialised;
        }
    }
    free(jobs);
    free(tid);
}

/* ------------------------------------------------------------------------------------------ */
/* Collisions — random_move.comp:545-562, :875-877                                             */
/* ------------------------------------------------------------------------------------------ */

/* random_move.comp:551-562 */
int orc_in_range(const float a[2], const float b[2], float max_distance) {
    const float dx = fabsf(b[0] - a[0]);
    if (dx > max_distance) return 0;
    const float dy = fabsf(b[1] - a[1]);
    if (dy > max_distance) return 0;
    /* distance(v1, v2) = length(v1 - v2); (v1-v2)^2 == (v2-v1)^2 bit for bit */
    return orc_length2(a[0] - b[0], a[1] - b[1]) < max_distance;
}

static inline void orc_paint(orc_entity* e, int hit) {
    /* :876 green, :546-547 blue */
    e->color[0] = 0.0f;
    e->color[1] = hit ? 0.0f : 1.0f;
    e->color[2] = hit ? 1.0f : 0.0f;
    e->color[3] = 1.0f;
}

uint64_t orc_collide_pass_brute(orc_entity* e, size_t n, float radius) {
    unsigned char* live = (unsigned char*)malloc(n ? n : 1);
    unsigned char* hit = (unsigned char*)calloc(n ? n : 1, 1);
    for (size_t i = 0; i < n; i++) live[i] = e[i].initialized != 0u;
    uint64_t pairs = 0;
    for (size_t i = 0; i < n; i++) {
        if (!live[i]) continue;
        for (size_t j = 0; j < i; j++) {
            if (!live[j]) continue;
            if (orc_in_range(e[j].pos, e[i].pos, radius)) {
                hit[i] = 1;
                hit[j] = 1;
                pairs++;
            }
        }
    }
    for (size_t i = 0; i < n; i++) {
        if (live[i]) orc_paint(&e[i], hit[i]);
        else e[i].initialized = 1u;
    }
    free(live);
    free(hit);
    return pairs;
}

/* Host cell grid (counting sort into CSR).  Cell edge = 1.25 * radius, indices computed in double:
 * deliberately NOT the CUDA path's key function, so the two are independent. */
typedef struct {
    const orc_entity* e;
    size_t n;
    float radius;
    double inv_cell;
    int64_t ncx, ncy;
    const uint32_t* cell_start; /* ncx*ncy + 1 */
    const uint32_t* order;      /* live entity ids grouped by cell */
    const int64_t* cell_of;     /* per entity, -1 if not live */
    unsigned char* hit;
    size_t lo, hi;
    uint64_t pairs;
} orc_grid_job;

static inline int64_t orc_cell_axis(double v, double inv, int64_t nc) {
    int64_t c = (int64_t)floor(v * inv);
    if (c < 0) c = 0;
    if (c >= nc) c = nc - 1;
    return c;
}

static void* orc_grid_thread(void* p) {
    orc_grid_job* j = (orc_grid_job*)p;
    uint64_t pairs = 0;
    for (size_t i = j->lo; i < j->hi; i++) {
        const int64_t c = j->cell_of[i];
        if (c < 0) continue;
        const int64_t cx = c % j->ncx, cy = c / j->ncx;
        int any = 0;
        for (int64_t yy = cy - 1; yy <= cy + 1; yy++) {
            if (yy < 0 || yy >= j->ncy) continue;
            for (int64_t xx = cx - 1; xx <= cx + 1; xx++) {
                if (xx < 0 || xx >= j->ncx) continue;
                const int64_t cc = yy * j->ncx + xx;
                for (uint32_t k = j->cell_start[cc]; k < j->cell_start[cc + 1]; k++) {
                    const uint32_t o = j->order[k];
                    if (o == i) continue;
                    if (orc_in_range(j->e[o].pos, j->e[i].pos, j->radius)) {
                        any = 1;
                        if (o < i) pairs++; /* count each unordered pair once */
                    }
                }
            }
        }
        j->hit[i] = (unsigned char)any;
    }
    j->pairs = pairs;
    return NULL;
}

uint64_t orc_collide_pass_grid_mt(orc_entity* e, size_t n, float world_w, float world_h, float radius, int threads) {
    if (threads < 1) threads = 1;
    if (n == 0) return 0;
    double cell = 1.25 * (double)radius;
    if (!(cell > 0.0)) cell = 1.0;
    /* bound the table: at most 2^26 cells */
    double maxx = world_w, maxy = world_h;
    for (size_t i = 0; i < n; i++) { /* tolerate entities outside the nominal world */
        if (e[i].pos[0] > maxx) maxx = e[i].pos[0];
        if (e[i].pos[1] > maxy) maxy = e[i].pos[1];
    }
    if (maxx < 1.0) maxx = 1.0;
    if (maxy < 1.0) maxy = 1.0;
    while ((maxx / cell + 1.0) * (maxy / cell + 1.0) > 67108864.0) cell *= 2.0;
    const int64_t ncx = (int64_t)floor(maxx / cell) + 1, ncy = (int64_t)floor(maxy / cell) + 1;
    const double inv = 1.0 / cell;
    const size_t ncell = (size_t)(ncx * ncy);

    int64_t* cell_of = (int64_t*)malloc(n * sizeof(int64_t));
    uint32_t* start = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    uint32_t* order = (uint32_t*)malloc(n * sizeof(uint32_t));
    unsigned char* hit = (unsigned char*)calloc(n, 1);
    for (size_t i = 0; i < n; i++) {
        if (e[i].initialized == 0u) {
            cell_of[i] = -1;
            continue;
        }
        const int64_t c = orc_cell_axis(e[i].pos[1], inv, ncy) * ncx + orc_cell_axis(e[i].pos[0], inv, ncx);
        cell_of[i] = c;
        start[c + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) start[c + 1] += start[c];
    uint32_t* fill = (uint32_t*)malloc(ncell * sizeof(uint32_t));
    memcpy(fill, start, ncell * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++)
        if (cell_of[i] >= 0) order[fill[cell_of[i]]++] = (uint32_t)i;
    free(fill);

    orc_grid_job* jobs = (orc_grid_job*)calloc((size_t)threads, sizeof(orc_grid_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        orc_grid_job* j = &jobs[t];
        j->e = e; j->n = n; j->radius = radius; j->inv_cell = inv; j->ncx = ncx; j->ncy = ncy;
        j->cell_start = start; j->order = order; j->cell_of = cell_of; j->hit = hit;
        j->lo = n * (size_t)t / (size_t)threads;
        j->hi = n * (size_t)(t + 1) / (size_t)threads;
        if (threads == 1) orc_grid_thread(j);
        else pthread_create(&tid[t], NULL, orc_grid_thread, j);
    }
    uint64_t pairs = 0;
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(tid[t], NULL);
        pairs += jobs[t].pairs;
    }
    for (size_t i = 0; i < n; i++) {
        if (cell_of[i] >= 0) orc_paint(&e[i], hit[i]);
        else e[i].initialized = 1u; /* init branch, random_move.comp:863-867 */
    }
    free(jobs); free(tid); free(cell_of); free(start); free(order); free(hit);
    return pairs;
}

uint64_t orc_collide_pass_grid(orc_entity* e, size_t n, float world_w, float world_h, float radius) {
    return orc_collide_pass_grid_mt(e, n, world_w, world_h, radius, 1);
}

/* random_move.comp:860-879 with the host's tick counter (src/sim/Simulator.cpp:220-235) */
uint64_t orc_dispatch(orc_entity* e, size_t n, const orc_map* map, float radius, uint32_t tick, orc_move_stats* stats) {
    if ((tick % 2u) == 0u) {
        orc_move_pass(e, n, map, stats);
        return 0;
    }
    return orc_collide_pass_grid(e, n, map->world_w, map->world_h, radius);
}

/* src/sim/GpuQuadTree.cpp:11-17 */
size_t orc_calc_node_count(size_t max_depth) {
    size_t result = 0, p = 1;
    for (size_t i = 0; i < max_depth; i++) {
        result += p;
        p *= 4;
    }
    return result;
}
