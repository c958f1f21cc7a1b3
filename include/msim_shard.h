/*
 * msim_shard.h — multi-GPU extension of the C ABI.  The reference runs on one device
 * (one kp::Manager, src/sim/Simulator.cpp:52); sharding is new work, so nothing here replaces a
 * reference call — it extends msim.h so that N handles (one per GPU, one process per GPU) together
 * produce exactly what one handle would.
 *
 * Partition: the neighbour grid's cell rows are split into contiguous bands, one per rank, the road
 * graph is replicated.  Per sim tick and rank:
 *     msim_enqueue_move            owned entities move (msim.h)
 *     msim_shard_pack              leavers -> migrant records, boundary-row entities -> halo, into two
 *                                  fixed-size DEVICE buffers (for the rank below / above)
 *     <exchange>                   caller sends/receives the buffers (NCCL send/recv via torch.distributed)
 *     msim_shard_integrate         arrivals join the owned set, halo + own leavers become ghosts
 *     msim_enqueue_collide         collision pass over owned + ghosts; ghosts get no flag, count no pair
 * The global unique-pair count is the sum of the ranks' last_pair_count (every pair is counted by the
 * rank that owns its higher-keyed member).  With collisions off no exchange is needed at all: shard by
 * entity range and use plain handles.
 */
#ifndef MSIM_SHARD_H
#define MSIM_SHARD_H

#include "msim.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Size in bytes of one exchange buffer: 32-byte header + migrant_capacity x 72 + halo_capacity x 8. */
uint64_t msim_shard_buffer_bytes(uint32_t migrant_capacity, uint32_t halo_capacity);

/* Turns a handle into one shard.  gids[i] = global id of resident entity i (host array, `count` ==
 * resident entity count).  Capacities bound what one tick may send to ONE neighbour; overflow is
 * reported as MSIM_ERR_CAPACITY by msim_shard_integrate.  The handle's entity_capacity must leave room
 * for arrivals and ghosts. */
int msim_shard_enable(msim_handle* h, const uint32_t* gids, uint64_t count, uint32_t migrant_capacity, uint32_t halo_capacity);

/* After a move pass: this rank owns cell rows [row_lo, row_hi).  send_down / send_up are DEVICE
 * buffers of msim_shard_buffer_bytes() (NULL when there is no neighbour on that side).  Enqueue only. */
int msim_shard_pack(msim_handle* h, uint32_t row_lo, uint32_t row_hi, void* send_down, void* send_up);

/* msim_enqueue_move + msim_shard_pack in ONE kernel: the move kernel classifies the entities it has just
 * moved and writes leavers / halo straight into the buffers, so the population is not read a second
 * time.  Migrant records then carry the state before the next-waypoint pass (arrival bit set); that pass
 * runs after msim_shard_integrate, beside the collision query, on the GPU that owns the entity by then.
 * Until the integrate call the handle accepts no move pass and no entity readback.  Enqueue only. */
int msim_shard_move_pack(msim_handle* h, uint32_t row_lo, uint32_t row_hi, void* send_down, void* send_up);

/* ---- peer-memory exchange: no collective call per tick ------------------------------------------------
 * The fused move + pack kernel writes leavers and halo DIRECTLY into the neighbours' receive buffers
 * (NVLink / NVSwitch peer stores and atomics), its last CTA raises a flag in the neighbour's memory after a
 * system-scope fence, and the neighbour's integrate kernel spins on that flag before it reads.  Two
 * receive buffers per side are used alternately, so tick t+1's stores never meet tick t's reads.  Per tick
 * and rank the host enqueues msim_shard_p2p_move_pack, msim_shard_p2p_integrate, msim_enqueue_collide and
 * never waits.  A neighbour that does not signal within MSIM_P2P_TIMEOUT_MS (default 10 000) is reported
 * as MSIM_ERR_INTERNAL by the next synchronising call; the GPU never hangs on it.
 *   msim_shard_p2p_create         allocates this handle's receive arena; returns its CUDA IPC handle
 *                                 (MSIM_P2P_HANDLE_BYTES bytes, for neighbours in other processes) and / or
 *                                 its device pointer (for neighbours driven by this process)
 *   msim_shard_p2p_connect        opens the neighbours' arenas from their IPC handles (NULL = no neighbour)
 *   msim_shard_p2p_connect_local  same, from device pointers of arenas living in this process.  Handles connected
 *                                 this way may share one stream if every band's move_pack is enqueued before
 *                                 any band's integrate (the flag wait would otherwise wait for a kernel queued
 *                                 behind itself); handles connected over IPC publish, wait and integrate in ONE
 *                                 kernel launched by msim_shard_p2p_integrate
 * All ranks must use the same migrant / halo capacities (they fix the buffer layout). */
#define MSIM_P2P_HANDLE_BYTES 64
int msim_shard_p2p_create(msim_handle* h, void* ipc_handle_out, void** arena_out);
int msim_shard_p2p_connect(msim_handle* h, const void* down_ipc_handle, const void* up_ipc_handle);
int msim_shard_p2p_connect_local(msim_handle* h, void* down_arena, void* up_arena);
int msim_shard_p2p_move_pack(msim_handle* h, uint32_t row_lo, uint32_t row_hi);
int msim_shard_p2p_integrate(msim_handle* h);

/* After the exchange: recv_down / recv_up are the DEVICE buffers received from the rank below / above
 * (NULL = none).  One host round trip (counts + hole list).  Returns the new owned / ghost counts. */
int msim_shard_integrate(msim_handle* h, const void* recv_down, const void* recv_up, uint64_t* owned, uint64_t* ghosts);

/* Asynchronous variant: the same integration done by kernels on the handle's stream, NO host round trip.
 * The owned / ghost counts stay in device memory and every following kernel of the handle reads them
 * there, so a whole sharded tick (move, pack, exchange, integrate, collide) can be enqueued without
 * waiting for the GPU.  Overflow is reported by the next call that synchronises (msim_sync, msim_get_stats,
 * msim_read_*, msim_shard_counts). */
int msim_shard_integrate_async(msim_handle* h, const void* recv_down, const void* recv_up);
/* Current owned / ghost counts (synchronises the stream when asynchronous ticks are outstanding). */
int msim_shard_counts(msim_handle* h, uint64_t* owned, uint64_t* ghosts);

/* Global ids of the owned entities, in the order msim_read_entities returns them. */
int msim_shard_read_gids(msim_handle* h, uint32_t* dst, uint64_t count);

/* Owned entities per cell row (rows == msim_stats.grid_cells_y): input of the band re-balancer. */
int msim_shard_row_histogram(msim_handle* h, uint32_t* dst, uint32_t rows);

/* Host helper, no GPU: the cell row every position falls into, computed exactly as the device does,
 * plus the grid dimensions — lets the caller build the initial partition. */
int msim_grid_rows(float world_w, float world_h, float radius, const float* xy, uint64_t count, uint32_t* rows_out,
                   uint32_t* cells_x, uint32_t* cells_y);

/* Host helper, no GPU: the neighbour grid the device builds for this world and radius - 1 / cell edge (the edge lies slightly above the
 * radius so that two points closer than the radius always fall into adjacent cells in spite of the binary32 rounding of pos * inv_cell) and
 * the exact binary32 threshold T with (d2 < T) <=> (sqrtf(d2) < radius), which lets the query skip the square root.  Any pointer may be NULL. */
int msim_grid_params(float world_w, float world_h, float radius, float* inv_cell, float* hit_threshold, uint32_t* cells_x, uint32_t* cells_y);

#ifdef __cplusplus
}
#endif
#endif /* MSIM_SHARD_H */
