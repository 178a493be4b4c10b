// zs_obs.cuh — observation encoder (zombsole/gym/observation.py:36-173), lanes over cells.
//
// World scope: the occupancy grid is compared, four cells (one 32-bit word) at a time, with the
// map's pristine template; where they agree (almost everywhere) the precomputed observation of
// the static layer is forwarded as one 128-bit load + one 128-bit streaming store per lane, so a
// warp writes 512 contiguous bytes per instruction.  Cells that differ (mobile things, dead
// bodies, damaged or destroyed boxes/walls) take the per-cell path.
#pragma once
#include "zs_device.cuh"

struct CellInfo { int label, life, weapon, agent; };

__device__ __forceinline__ CellInfo cell_info(const ZsParams& p, const Env& e, int c, int g) {
    CellInfo ci;
    ci.life = 0; ci.weapon = 0; ci.agent = -1;
    if (g == G_EMPTY) ci.label = objective_bit(p, c) ? ZS_LABEL_OBJECTIVE : 0;
    else if (g == G_DEAD) ci.label = ZS_LABEL_DEAD_BODY;
    else if (g > G_MAX_SLOT) {
        const int i = __ldg(p.cell_static + c);
        ci.label = __ldg(p.static_label + i);
        ci.life = e.slife[i];
    } else {
        const int s = g - 1;
        ci.life = e.tl[s];
        ci.weapon = e.tm[s] & 15;
        ci.label = s < p.P ? ZS_LABEL_PLAYER : s < p.P + p.A ? ZS_LABEL_AGENT : ZS_LABEL_ZOMBIE;
        ci.agent = s - p.P;
    }
    return ci;
}
// encode_position_simple (observation.py:36-55)
__device__ __forceinline__ int encode_simple(const CellInfo& ci) {
    if (ci.label == 0) return 0;
    const int adj = ci.life < 100 ? ci.life : 100;
    return 256 * ci.label + 16 * ci.weapon + floordiv100(15 * adj);
}
// encode_position_as_channels (observation.py:57-81): thing code of an agent is 8 + int(agent_id)
__device__ __forceinline__ int channel_label(const ZsParams& p, const CellInfo& ci) {
    return ci.label == ZS_LABEL_AGENT ? 8 + p.agent_obs_ids[ci.agent] : ci.label;
}

__device__ __forceinline__ void encode_world(const ZsParams& p, const Env& e, int32_t* __restrict__ obs) {
    const int lane = e.lane;
    const int cells = p.cells;
    if ((cells & 3) == 0) {
        const int n4 = cells >> 2;
        const uint32_t* g32 = (const uint32_t*)e.grid;
        const uint32_t* t32 = (const uint32_t*)p.tmpl_grid;
        const uint4* to4 = (const uint4*)p.tmpl_obs;
        uint4* o4 = (uint4*)obs;
        if (p.obs_enc == ZS_OBS_SIMPLE) {
#pragma unroll 2
            for (int i = lane; i < n4; i += 32) {
                const uint32_t g = g32[i];
                uint4 v;
                if (g == __ldg(t32 + i)) v = __ldg(to4 + i);
                else {
                    v.x = encode_simple(cell_info(p, e, 4 * i, g & 255));
                    v.y = encode_simple(cell_info(p, e, 4 * i + 1, (g >> 8) & 255));
                    v.z = encode_simple(cell_info(p, e, 4 * i + 2, (g >> 16) & 255));
                    v.w = encode_simple(cell_info(p, e, 4 * i + 3, g >> 24));
                }
                __stcs(o4 + i, v);
            }
        } else {
            for (int i = lane; i < n4; i += 32) {
                const uint32_t g = g32[i];
                uint4 v0, v1, v2 = make_uint4(0, 0, 0, 0);
                if (g == __ldg(t32 + i)) { v0 = __ldg(to4 + i); v1 = __ldg(to4 + n4 + i); }
                else {
                    CellInfo a = cell_info(p, e, 4 * i, g & 255), b = cell_info(p, e, 4 * i + 1, (g >> 8) & 255);
                    CellInfo c = cell_info(p, e, 4 * i + 2, (g >> 16) & 255), d = cell_info(p, e, 4 * i + 3, g >> 24);
                    v0 = make_uint4(channel_label(p, a), channel_label(p, b), channel_label(p, c), channel_label(p, d));
                    v1 = make_uint4(a.life, b.life, c.life, d.life);
                    v2 = make_uint4(a.weapon, b.weapon, c.weapon, d.weapon);
                }
                __stcs(o4 + i, v0); __stcs(o4 + n4 + i, v1); __stcs(o4 + 2 * n4 + i, v2);
            }
        }
        return;
    }
    // maps whose cell count is not a multiple of 4 (rows of an env are then not 16-byte aligned)
    for (int c = lane; c < cells; c += 32) {
        const CellInfo ci = cell_info(p, e, c, e.grid[c]);
        if (p.obs_enc == ZS_OBS_SIMPLE) __stcs(obs + c, encode_simple(ci));
        else {
            __stcs(obs + c, channel_label(p, ci));
            __stcs(obs + cells + c, ci.life);
            __stcs(obs + 2 * cells + c, ci.weapon);
        }
    }
}

// surroundings window (observation.py:99-119): rows = y, columns = x, centred on the agent's
// (possibly stale, if dead) position; out-of-bounds cells are a fresh Wall (observation.py:43-44,64-65)
__device__ __forceinline__ void encode_surroundings(const ZsParams& p, const Env& e, int32_t* __restrict__ obs) {
    const int lane = e.lane;
    const int w = p.sw, half = p.sw >> 1, ww = p.sw * p.sw;
    for (int a = 0; a < p.obs_count; ++a) {
        const int ax = e.tx[p.P + a] - half, ay = e.ty[p.P + a] - half;
        int32_t* o = obs + (size_t)a * p.obs_C * ww;
        for (int i = lane; i < ww; i += 32) {
            const int r = i / w, c = i - r * w;
            const int x = ax + c, y = ay + r;
            CellInfo ci;
            if ((unsigned)x >= (unsigned)p.W || (unsigned)y >= (unsigned)p.H) {
                ci.label = ZS_LABEL_WALL; ci.life = 200; ci.weapon = 0; ci.agent = -1;
            } else {
                const int cell = y * p.W + x;
                ci = cell_info(p, e, cell, e.grid[cell]);
            }
            if (p.obs_enc == ZS_OBS_SIMPLE) __stcs(o + i, encode_simple(ci));
            else {
                __stcs(o + i, channel_label(p, ci));
                __stcs(o + ww + i, ci.life);
                __stcs(o + 2 * ww + i, ci.weapon);
            }
        }
    }
}

__device__ __forceinline__ void encode_obs(const ZsParams& p, const Env& e, int32_t* obs) {
    if (p.obs_scope == ZS_OBS_WORLD) encode_world(p, e, obs);
    else encode_surroundings(p, e, obs);
}
