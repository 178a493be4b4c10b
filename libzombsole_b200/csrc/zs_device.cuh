// zs_device.cuh — device-side data model and helpers of the batched zombsole simulator (sm_100a).
//
// Execution model: ONE WARP PER ENVIRONMENT, ZS_WPC environments per CTA.  The world of an
// environment lives in shared memory while its warp works on it:
//   * an occupancy grid, one byte per map cell (0 empty, 1..250 mobile slot+1, 253/255 box or wall,
//     254 dead-body decoration) — the shared-memory-staged stand-in for the reference's
//     position-keyed dict World.things (zombsole/core.py:15) and World.decoration (core.py:16);
//   * the mobile things as structure-of-arrays (x, y, life, dict-order stamp, meta).
// Phases that are parallel in the reference's semantics (every actor decides against the
// pre-step world; observation cells are independent) are spread over the 32 lanes; the phases
// that the reference defines sequentially (shuffle, execute_actions) are run by lane 0 on
// shared memory.  All per-thing and per-cell loops are lane-strided so the same code serves 13
// things on bridge and 101 on the maze.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/zs_b200.h"

#define ZS_WPC 4              // warps (= environments) per CTA
#define ZS_MIN_CTAS 7         // 7 CTAs x 4 warps = 28 envs resident per SM: 4,096 envs fit the 148 SMs in one wave
#define ZS_FULL 0xffffffffu

// occupancy-grid byte codes
#define G_EMPTY 0
#define G_MAX_SLOT 250        // 1..250: mobile slot + 1
#define G_STATIC_DMG 253      // box/wall present, life != MAX_LIFE (observation takes the per-cell path)
#define G_DEAD 254            // DeadBody decoration and no thing
#define G_STATIC 255          // pristine box/wall present

// decided action types (the 2-tuples things return from next_step, core.py:87-90)
#define D_IDLE 0
#define D_MOVE 1
#define D_ATTACK 2
#define D_HEAL 3
#define D_WANDER 4            // zombie with no humans: destination drawn in dict order (things.py:101-103)

struct ZsParams {
    // ---- configuration
    int32_t N;
    uint32_t env_base;        // low 32 bits of the global index of env 0
    uint32_t key0, key1;      // Philox key = seed
    int32_t rules, P, A, Z, M, Mp, Ap, S, Sp, W, H, cells, cells_pad, dead_words;
    int32_t initial_zombies, minimum_zombies;
    int32_t obs_scope, obs_enc, sw, obs_count, obs_C;
    int32_t obs_per_agent, max_steps, auto_reset, n_discrete;
    int32_t n_ps, n_zs;
    int64_t obs_elems;
    uint8_t agent_weapons[ZS_MAX_AGENTS];
    int32_t agent_obs_ids[ZS_MAX_AGENTS];
    // ---- map tables (device, read-only)
    const int16_t* cell_static;    // [cells] static index or -1
    const uint16_t* static_cell;   // [Sp] cell of static i
    const int16_t* static_max;     // [Sp] MAX_LIFE of static i (0 in the padding)
    const uint8_t* static_label;   // [Sp]
    const uint8_t* tmpl_grid;      // [cells_pad] G_STATIC on box/wall cells, else 0
    const int32_t* tmpl_obs;       // [2][cells] world-scope observation of the pristine static layer
    const uint32_t* objective_bits;// [dead_words]
    const uint16_t* ps_cells;      // [n_ps] player spawn cells, file order
    const uint16_t* zs_cells;      // [n_zs]
    // ---- state (device, caller-owned buffer; see ZsLayout)
    int16_t* X; int16_t* Y; int16_t* LIFE; int32_t* STAMP; uint8_t* META;
    int16_t* PREV; int16_t* SLIFE; uint32_t* DEAD; int32_t* SCAL;
    unsigned long long* stats;     // [4]
    // ---- shared-memory carve-up (bytes from the warp's base)
    int32_t off_dead, off_tx, off_ty, off_tl, off_ts, off_tm, off_dtype, off_da, off_db, off_act, off_draws,
        off_cand, off_list, off_prev, off_acts, off_sl, off_cq, off_ats, off_scal, off_rk, off_sor, off_zb;
    int32_t draws_cap, cand_cap;
    int32_t smem_per_warp;
};

struct ZsIO {
    const int32_t* actions;   // [n_steps, N, A(,3)] or NULL (synthetic)
    int32_t fmt;
    int32_t* obs;             // [obs_slots, N, obs_elems] or NULL
    int32_t obs_slots;
    double* reward;           // [n_steps, N(,A)] or NULL
    uint8_t* terminated;      // [n_steps, N] or NULL
    uint8_t* truncated;
    uint8_t* agent_mask;      // [n_steps, N, A] or NULL
    int32_t* draws;           // [n_steps, N] or NULL
    const uint8_t* env_mask;  // reset only
    int32_t n_steps;
    int64_t first_step;
    int32_t force_auto_reset;
};

// weapon tables indexed by weapon code (zombsole/weapons.py:18-25); range2 = floor(max_range^2)
__constant__ int16_t c_range2[16] = {0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 36, 100, 9, 0};
__constant__ int16_t c_dmg_lo[16] = {0, 5, 0, 0, 0, 0, 0, 0, 0, 0, 5, 75, 10, 25, 75, 0};
__constant__ int16_t c_dmg_n[16] = {1, 6, 1, 1, 1, 1, 1, 1, 1, 1, 6, 26, 41, 51, 26, 1};  // hi - lo + 1

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1;
        c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The environment a warp is working on.
struct Env {
    // shared memory views
    uint8_t* grid; uint32_t* dead;
    int16_t* tx; int16_t* ty; int16_t* tl; int32_t* ts; uint8_t* tm;
    uint8_t* dtype; int16_t* da; int16_t* db;
    unsigned long long* act; uint32_t* draws; uint16_t* cand; uint16_t* list; int16_t* prev; int32_t* acts;
    int16_t* sl;     // static lives, staged for the whole launch
    uint2* cq;       // per step: {x | y << 16 (sentinel when not in the world), dict-order stamp}
    int32_t* ats;    // per step: closest-player key of a zombie / heal_closest agent
    uint8_t* rk;     // per step: dict-order rank of a slot (255 when not in the world)
    uint8_t* sor;    // per step: slot of a rank
    uint32_t* zb;    // per step: closest-zombie key of a player slot
    int32_t* scal;   // 8 ints: hand-off of the scalars to/from the out-of-line (re)initialisation
    // warp-uniform registers
    int32_t t, episode, deaths, zd, stampctr, flags, prev_zd, ep_steps;
    int32_t env; uint32_t env_global;
    int32_t lane;
};
// Env::flags: bit0 is state (ZS_S_FLAGS); the others live for one launch only
#define FL_FRESH 1       // first step of a world: boxes/walls with life <= 0 are still present
#define FL_DMG 2         // some box/wall has life != MAX_LIFE (else the observation needs no static patches)
#define FL_SL_DIRTY 4    // static lives changed during this launch: write them back


__device__ __forceinline__ uint32_t draw_at(const ZsParams& p, const Env& e, uint32_t t_word, int k) {
    uint32_t o[4];
    philox4x32_10(e.env_global, (uint32_t)e.episode, t_word, (uint32_t)(k >> 2), p.key0, p.key1, o);
    int w = k & 3;
    return w == 0 ? o[0] : w == 1 ? o[1] : w == 2 ? o[2] : o[3];
}
__device__ __forceinline__ int below(uint32_t u, int n) { return (int)__umulhi(u, (uint32_t)n); }

__device__ __forceinline__ bool g_is_thing(int g) { return g != G_EMPTY && g != G_DEAD; }
__device__ __forceinline__ bool g_is_static(int g) { return g == G_STATIC || g == G_STATIC_DMG; }

// World.things.get((x, y)) as a grid byte; positions outside the map hold nothing
__device__ __forceinline__ int grid_at(const ZsParams& p, const Env& e, int x, int y) {
    if ((unsigned)x >= (unsigned)p.W || (unsigned)y >= (unsigned)p.H) return G_EMPTY;
    return e.grid[y * p.W + x];
}
__device__ __forceinline__ int dist2(int x1, int y1, int x2, int y2) {
    int dx = x1 - x2, dy = y1 - y2;
    return dx * dx + dy * dy;
}
__device__ __forceinline__ bool dead_bit(const Env& e, int c) { return (e.dead[c >> 5] >> (c & 31)) & 1u; }
__device__ __forceinline__ bool objective_bit(const ZsParams& p, int c) {
    return (__ldg(p.objective_bits + (c >> 5)) >> (c & 31)) & 1u;
}
// adjacent_positions order (zombsole/utils.py:34-44): (0,+1), (0,-1), (+1,0), (-1,0)
__device__ __forceinline__ int adj_dx(int a) { return a == 2 ? 1 : a == 3 ? -1 : 0; }
__device__ __forceinline__ int adj_dy(int a) { return a == 0 ? 1 : a == 1 ? -1 : 0; }

__device__ __forceinline__ int floordiv100(int a) { return a >= 0 ? a / 100 : -((-a + 99) / 100); }
__device__ __forceinline__ int max_life_of_label(int label) { return label == ZS_LABEL_BOX ? 10 : 200; }
