// zs_world.cuh — the world transition, rules, rewards and world (re)initialisation, one warp per env.
// Reference line numbers are relative to the reference tree (jvstinian/libzombsole v0.13.2).
#pragma once
#include "zs_device.cuh"

// ---------------------------------------------------------------- state <-> shared memory
__device__ __forceinline__ void env_bind(const ZsParams& p, Env& e, unsigned char* base, int env, int lane) {
    e.grid = base;
    e.dead = (uint32_t*)(base + p.off_dead);
    e.tx = (int16_t*)(base + p.off_tx); e.ty = (int16_t*)(base + p.off_ty); e.tl = (int16_t*)(base + p.off_tl);
    e.ts = (int32_t*)(base + p.off_ts); e.tm = base + p.off_tm;
    e.dtype = base + p.off_dtype; e.da = (int16_t*)(base + p.off_da); e.db = (int16_t*)(base + p.off_db);
    e.act = (unsigned long long*)(base + p.off_act); e.draws = (uint32_t*)(base + p.off_draws);
    e.cand = (uint16_t*)(base + p.off_cand); e.list = (uint16_t*)(base + p.off_list);
    e.prev = (int16_t*)(base + p.off_prev); e.acts = (int32_t*)(base + p.off_acts);
    e.sl = (int16_t*)(base + p.off_sl); e.cq = (uint2*)(base + p.off_cq); e.ats = (int32_t*)(base + p.off_ats);
    e.scal = (int32_t*)(base + p.off_scal);
    e.rk = base + p.off_rk; e.sor = base + p.off_sor; e.zb = (uint32_t*)(base + p.off_zb);
    e.env = env; e.env_global = p.env_base + (uint32_t)env; e.lane = lane;
}

__device__ __forceinline__ void scalars_from_lane(Env& e, int sc) {
    e.t = __shfl_sync(ZS_FULL, sc, ZS_S_T); e.episode = __shfl_sync(ZS_FULL, sc, ZS_S_EPISODE);
    e.deaths = __shfl_sync(ZS_FULL, sc, ZS_S_DEATHS); e.zd = __shfl_sync(ZS_FULL, sc, ZS_S_ZOMBIE_DEATHS);
    e.stampctr = __shfl_sync(ZS_FULL, sc, ZS_S_STAMP_COUNTER); e.flags = __shfl_sync(ZS_FULL, sc, ZS_S_FLAGS);
    e.prev_zd = __shfl_sync(ZS_FULL, sc, ZS_S_PREV_ZOMBIE_DEATHS); e.ep_steps = __shfl_sync(ZS_FULL, sc, ZS_S_EPISODE_STEPS);
}
__device__ __forceinline__ int scalar_of_lane(const Env& e) {
    const int l = e.lane;
    return l == ZS_S_T ? e.t : l == ZS_S_EPISODE ? e.episode : l == ZS_S_DEATHS ? e.deaths
         : l == ZS_S_ZOMBIE_DEATHS ? e.zd : l == ZS_S_STAMP_COUNTER ? e.stampctr
         : l == ZS_S_FLAGS ? e.flags : l == ZS_S_PREV_ZOMBIE_DEATHS ? e.prev_zd : e.ep_steps;
}
// hand-off through shared memory around the out-of-line functions (keeps Env in registers)
__device__ __forceinline__ void scalars_to_smem(Env& e) {
    if (e.lane < 8) e.scal[e.lane] = scalar_of_lane(e);
    __syncwarp();
}
__device__ __forceinline__ void scalars_from_smem(Env& e) {
    __syncwarp();
    scalars_from_lane(e, e.lane < 8 ? e.scal[e.lane] : 0);
}

__device__ __forceinline__ int16_t half_of(const uint4& v, int q) {
    uint32_t w = q < 2 ? v.x : q < 4 ? v.y : q < 6 ? v.z : v.w;
    return (int16_t)((q & 1) ? (w >> 16) : (w & 0xffffu));
}

__device__ __forceinline__ void load_state(const ZsParams& p, Env& e) {
    const size_t row = (size_t)e.env * p.Mp;
#pragma unroll 1
    for (int s = e.lane; s < p.Mp; s += 32) {
        e.tx[s] = p.X[row + s]; e.ty[s] = p.Y[row + s]; e.tl[s] = p.LIFE[row + s];
        e.ts[s] = p.STAMP[row + s]; e.tm[s] = p.META[row + s];
    }
#pragma unroll 1
    for (int w = e.lane; w < p.dead_words; w += 32) e.dead[w] = p.DEAD[(size_t)e.env * p.dead_words + w];
#pragma unroll 1
    for (int a = e.lane; a < p.Ap; a += 32) e.prev[a] = p.PREV[(size_t)e.env * p.Ap + a];
    // static lives: staged in shared memory for the launch; note whether any differs from its MAX_LIFE
    const uint4* sl4 = (const uint4*)(p.SLIFE + (size_t)e.env * p.Sp);
    const uint4* mx4 = (const uint4*)p.static_max;
    bool dmg = false;
#pragma unroll 1
    for (int i = e.lane; i < (p.Sp >> 3); i += 32) {
        const uint4 a = sl4[i];
        const uint4 m = __ldg(mx4 + i);
        ((uint4*)e.sl)[i] = a;
        dmg |= a.x != m.x || a.y != m.y || a.z != m.z || a.w != m.w;
    }
    scalars_from_lane(e, e.lane < 8 ? p.SCAL[(size_t)e.env * 8 + e.lane] : 0);
    e.flags = (e.flags & FL_FRESH) | (__any_sync(ZS_FULL, dmg) ? FL_DMG : 0);
    __syncwarp();
}

__device__ __forceinline__ void store_state(const ZsParams& p, Env& e) {
    __syncwarp();
    const size_t row = (size_t)e.env * p.Mp;
#pragma unroll 1
    for (int s = e.lane; s < p.Mp; s += 32) {
        p.X[row + s] = e.tx[s]; p.Y[row + s] = e.ty[s]; p.LIFE[row + s] = e.tl[s];
        p.STAMP[row + s] = e.ts[s]; p.META[row + s] = e.tm[s];
    }
#pragma unroll 1
    for (int w = e.lane; w < p.dead_words; w += 32) p.DEAD[(size_t)e.env * p.dead_words + w] = e.dead[w];
#pragma unroll 1
    for (int a = e.lane; a < p.Ap; a += 32) p.PREV[(size_t)e.env * p.Ap + a] = e.prev[a];
    if (e.flags & FL_SL_DIRTY) {
        uint4* sl4 = (uint4*)(p.SLIFE + (size_t)e.env * p.Sp);
#pragma unroll 1
        for (int i = e.lane; i < (p.Sp >> 3); i += 32) sl4[i] = ((const uint4*)e.sl)[i];
    }
    const int keep = e.flags;
    e.flags &= FL_FRESH;
    if (e.lane < 8) p.SCAL[(size_t)e.env * 8 + e.lane] = scalar_of_lane(e);
    e.flags = keep;
}

// Rebuild the occupancy grid from the compact state.  FL_FRESH = first step after a world init:
// boxes/walls whose life is already <= 0 are still in World.things (game.py:154-155) until the
// first clean_dead_things.
__device__ __noinline__ void build_grid(const ZsParams& p, unsigned char* base, int env, int lane, int flags) {
    Env e;
    env_bind(p, e, base, env, lane);
    const bool fresh = flags & FL_FRESH;
    const uint4* tg = (const uint4*)p.tmpl_grid;
    uint4* g4 = (uint4*)e.grid;
#pragma unroll 1
    for (int i = lane; i < (p.cells_pad >> 4); i += 32) g4[i] = __ldg(tg + i);
    __syncwarp();
    if (flags & FL_DMG) {
        const uint4* mx4 = (const uint4*)p.static_max;
#pragma unroll 1
        for (int i = lane; i < (p.Sp >> 3); i += 32) {
            const uint4 a = ((const uint4*)e.sl)[i];
            const uint4 m = __ldg(mx4 + i);
            if (a.x != m.x || a.y != m.y || a.z != m.z || a.w != m.w) {
#pragma unroll 1
                for (int q = 0; q < 8; ++q) {
                    const int life = e.sl[i * 8 + q], mx = __ldg(p.static_max + i * 8 + q);
                    if (life != mx) e.grid[__ldg(p.static_cell + i * 8 + q)] = (life <= 0 && !fresh) ? G_EMPTY : G_STATIC_DMG;
                }
            }
        }
        __syncwarp();
    }
#pragma unroll 1
    for (int w = lane; w < p.dead_words; w += 32) {
        uint32_t bits = e.dead[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int c = w * 32 + b;
            if (e.grid[c] == G_EMPTY) e.grid[c] = G_DEAD;
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int s = lane; s < p.M; s += 32)
        if (e.tm[s] & 0x80) e.grid[e.ty[s] * p.W + e.tx[s]] = (uint8_t)(s + 1);
    __syncwarp();
}

// ---------------------------------------------------------------- decide phase
// target id: mobile slot s -> s, static i -> M + i
__device__ __forceinline__ int target_of_cell(const ZsParams& p, int g, int cell) {
    return g <= G_MAX_SLOT ? g - 1 : p.M + (int)__ldg(p.cell_static + cell);
}

// A decided action, packed so that the sequential execute loop touches as little as possible:
//   bits 0-7 actor | 8-10 kind | 11-26 a | 27-42 b | 43-49 range^2 | 50-56 lo | 57-62 n
// kind: X_NOP (an action that can no longer have an effect but still takes part in the shuffle),
// X_MOVE (a, b = destination, already known to be in bounds and one step away), X_ATTACK_M /
// X_HEAL_M (a = mobile target slot; range is checked against current positions at execute time),
// X_ATTACK_S / X_HEAL_S (a = static index, already known to be in range: neither end can move
// before the actor acts).  lo/n: the draw is lo + randbelow(n).
#define X_NOP 0
#define X_MOVE 1
#define X_ATTACK_M 2
#define X_HEAL_M 3
#define X_ATTACK_S 4
#define X_HEAL_S 5
__device__ __forceinline__ unsigned long long pack_action(int actor, int kind, int a, int b, int r2, int lo, int n) {
    return (unsigned long long)(uint32_t)actor | ((unsigned long long)(uint32_t)kind << 8) |
           ((unsigned long long)(uint16_t)(int16_t)a << 11) | ((unsigned long long)(uint16_t)(int16_t)b << 27) |
           ((unsigned long long)(uint32_t)r2 << 43) | ((unsigned long long)(uint32_t)lo << 50) |
           ((unsigned long long)(uint32_t)n << 57);
}

// ---------------------------------------------------------------- World.step (core.py:72-78)
// Returns the number of draws consumed so far in this step's draw cell.
__device__ __forceinline__ int world_step(const ZsParams& p, Env& e) {
    const int lane = e.lane;
    const int NP = p.P + p.A;
    e.t += 1;
    const uint32_t t_word = (uint32_t)(e.t + 1);

    // ---- per-step tables: packed position + stamp of every slot, which players need their closest zombie
    bool humans = false;
    unsigned needz0 = 0, needz1 = 0;  // players (slots < NP <= 64) that look for the closest zombie
#pragma unroll 1
    for (int s0 = 0; s0 < p.Mp; s0 += 32) {
        const int s = s0 + lane;
        bool live = false, needz = false;
        if (s < p.Mp) {
            live = (e.tm[s] & 0x80) != 0;  // padding slots are never in the world
            e.cq[s] = make_uint2((uint32_t)(uint16_t)e.tx[s] | ((uint32_t)(uint16_t)e.ty[s] << 16),
                                 live ? (uint32_t)e.ts[s] : 0x7fffffffu);
            if (live && s < NP) {
                humans = true;
                // terminators always (terminator.py:10-14), agents for attack_closest (agent.py:41-47)
                needz = s < p.P || e.acts[3 * (s - p.P)] == ZS_ACT_ATTACK_CLOSEST;
            }
        }
        if (s0 < NP) {
            const unsigned m = __ballot_sync(ZS_FULL, needz);
            if (s0 == 0) needz0 = m; else needz1 = m;
        }
        if (s < NP) e.zb[s] = 0xffffffffu;
    }
    const bool has_humans = __any_sync(ZS_FULL, humans);
    __syncwarp();
    // ---- dict-order rank of every thing in the world = number of things with a smaller stamp.  It breaks
    // distance ties (sorted() is stable, utils.py:23-31) and orders the actors (core.py:83-90).
#pragma unroll 1
    for (int s0 = 0; s0 < p.Mp; s0 += 32) {
        const int s = s0 + lane;
        const int st = s < p.Mp ? (int)e.cq[s].y : 0x7fffffff;
        int r = 0;
#pragma unroll 4
        for (int j = 0; j < p.Mp; ++j) r += (int)e.cq[j].y < st;
        if (s < p.Mp) {
            const bool live = st != 0x7fffffff;
            e.rk[s] = live ? (uint8_t)r : (uint8_t)255;
            if (live) e.sor[r] = (uint8_t)s;
        }
    }
    __syncwarp();
    // ---- closest(self, others) (utils.py:23-31) for everybody from ONE pass over the (thing, player) distances:
    // a zombie (things.py:73-82) or a heal_closest agent (agent.py:79-86) takes the minimum over the players in
    // its own lane; a player's closest zombie (terminator.py:10-14, agent.py:41-47) is the minimum of the same
    // distances across the zombie lanes (redux.sync).  Key = (d^2 << 8) | dict rank: ties go to the earlier thing.
#pragma unroll 1
    for (int s0 = 0; s0 < p.M; s0 += 32) {
        const int s = s0 + lane;
        const bool live = s < p.M && (e.tm[s] & 0x80);
        const bool zombie = s >= NP;
        const int x = live ? e.tx[s] : 0, y = live ? e.ty[s] : 0;
        const uint32_t myrank = live ? e.rk[s] : 255u;
        const bool wantp = live && (zombie ? has_humans : (s >= p.P && e.acts[3 * (s - p.P)] == ZS_ACT_HEAL_CLOSEST));
        uint32_t bestp = 0xffffffffu;
#pragma unroll 1
        for (int q = 0; q < NP; ++q) {
            const uint2 c = e.cq[q];
            if (c.y == 0x7fffffffu) continue;  // player q is not in the world (warp-uniform)
            const int dx = x - (int)(int16_t)(c.x & 0xffffu), dy = y - (int)(int16_t)(c.x >> 16);
            const uint32_t d = (uint32_t)(dx * dx + dy * dy) << 8;
            if (wantp && q != s) bestp = min(bestp, d | (uint32_t)e.rk[q]);
            if (((q < 32 ? needz0 : needz1) >> (q & 31)) & 1u) {
                const uint32_t m = __reduce_min_sync(ZS_FULL, (live && zombie) ? (d | myrank) : 0xffffffffu);
                if (lane == 0 && m < e.zb[q]) e.zb[q] = m;
            }
        }
        if (s < p.Mp) e.ats[s] = (int)bestp;
    }
    __syncwarp();

    // ---- get_actions (core.py:80-101): every actor decides against the pre-step world
    bool any_wander = false;
    int n_idle = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < p.M; s0 += 32) {
        const int s = s0 + lane;
        const bool live = s < p.M && (e.tm[s] & 0x80);
        const int x = live ? e.tx[s] : 0, y = live ? e.ty[s] : 0;
        const bool zombie = s >= NP, agent = !zombie && s >= p.P;
        int at = ZS_ACT_NONE, adx = 0, ady = 0;
        if (live && agent) {
            at = e.acts[3 * (s - p.P)]; adx = e.acts[3 * (s - p.P) + 1]; ady = e.acts[3 * (s - p.P) + 2];
            if (at == ZS_ACT_ABSENT) { at = ZS_ACT_HEAL; adx = 0; ady = 0; }  // multiagent_env.py:129-131
        }
        uint32_t key = 0xffffffffu;
        if (live) key = (zombie || at == ZS_ACT_HEAL_CLOSEST) ? (uint32_t)e.ats[s] : e.zb[s];
        const int tg = key == 0xffffffffu ? -1 : (int)e.sor[key & 255u];
        const int d2 = (int)(key >> 8);
        int type = D_IDLE, a = 0, b = 0;
        if (live) {
            const int gx = tg >= 0 ? e.tx[tg] : 0, gy = tg >= 0 ? e.ty[tg] : 0;
            // the four adjacent cells (utils.py:34-44): what is on them and how far they are from the target
            unsigned freemask = 0, gs[4];
            int dd[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {  // no bounds check: cells outside the map hold nothing (utils.py:47-52)
                gs[d] = grid_at(p, e, x + adj_dx(d), y + adj_dy(d));
                dd[d] = dist2(gx, gy, x + adj_dx(d), y + adj_dy(d));
                if (!g_is_thing(gs[d])) freemask |= 1u << d;
            }
            if (zombie) {  // Zombie.next_step (things.py:70-105)
                if (!has_humans) {
                    if (freemask) { type = D_WANDER; a = (int)freemask; any_wander = true; }
                } else if (d2 <= 2) { type = D_ATTACK; a = tg; }  // distance < 1.5 (things.py:83)
                else {
                    // free cells: closest(target, positions), first minimum in adjacency order; boxed in: the
                    // first Box/Wall among the adjacent cells stably sorted by distance to the target (things.py:88-99)
                    int bdir = -1, bdist = 0x7fffffff;
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        const bool cand = freemask ? ((freemask >> d) & 1u) : g_is_static(gs[d]);
                        if (cand && dd[d] < bdist) { bdir = d; bdist = dd[d]; }
                    }
                    if (bdir >= 0) {
                        const int cx = x + adj_dx(bdir), cy = y + adj_dy(bdir);
                        if (freemask) { type = D_MOVE; a = cx; b = cy; }
                        else { type = D_ATTACK; a = p.M + (int)__ldg(p.cell_static + cy * p.W + cx); }
                    }
                }
            } else if (!agent) {  // Terminator.next_step (players/terminator.py:9-37)
                if (tg < 0) { type = D_HEAL; a = s; }
                else if (d2 > c_range2[e.tm[s] & 15]) {
                    int bdir = 0, bdist = 0x7fffffff, g = 0;
#pragma unroll
                    for (int d = 0; d < 4; ++d)  // closest(target, adjacent_positions(self)): out-of-bounds cells included
                        if (dd[d] < bdist) { bdir = d; bdist = dd[d]; g = (int)gs[d]; }
                    const int bx = x + adj_dx(bdir), by = y + adj_dy(bdir);
                    if (g_is_thing(g)) {
                        type = (g <= G_MAX_SLOT && (g - 1) < NP) ? D_HEAL : D_ATTACK;
                        a = target_of_cell(p, g, by * p.W + bx);
                    } else { type = D_MOVE; a = bx; b = by; }
                } else { type = D_ATTACK; a = tg; }
            } else {  // Agent.next_step (players/agent.py:28-96)
                if (at == ZS_ACT_MOVE) { type = D_MOVE; a = x + adx; b = y + ady; }
                else if (at == ZS_ACT_ATTACK_CLOSEST) { if (tg >= 0) { type = D_ATTACK; a = tg; } }
                else if (at == ZS_ACT_HEAL_CLOSEST) { type = D_HEAL; a = tg >= 0 ? tg : s; }
                else if (at == ZS_ACT_ATTACK || at == ZS_ACT_HEAL) {
                    if (at == ZS_ACT_HEAL && adx == 0 && ady == 0) { type = D_HEAL; a = s; }
                    else {
                        const int g = grid_at(p, e, x + adx, y + ady);
                        // attack: any thing; heal: Player / Box / Wall only (agent.py:69-75)
                        const bool ok = at == ZS_ACT_ATTACK ? g_is_thing(g)
                                                            : (g_is_static(g) || (g_is_thing(g) && (g - 1) < NP));
                        if (ok) { type = at == ZS_ACT_ATTACK ? D_ATTACK : D_HEAL; a = target_of_cell(p, g, (y + ady) * p.W + (x + adx)); }
                    }
                }
            }
            n_idle += type == D_IDLE;
        }
        if (s < p.Mp) { e.dtype[s] = (uint8_t)type; e.da[s] = (int16_t)a; e.db[s] = (int16_t)b; }
    }
    n_idle = __reduce_add_sync(ZS_FULL, n_idle);
    __syncwarp();
    int nd = 0;
    if (__any_sync(ZS_FULL, any_wander)) {
        // wandering zombies draw random.choice(positions) in dict (= stamp) order (things.py:101-103)
        int mine = 0;
#pragma unroll 1
        for (int s = lane; s < p.M; s += 32) {
            if (e.dtype[s] != D_WANDER) continue;
            int rank = 0;
            for (int j = NP; j < p.M; ++j) rank += (e.dtype[j] == D_WANDER && e.ts[j] < e.ts[s]);
            unsigned fm = (unsigned)e.da[s];
            int pick = below(draw_at(p, e, t_word, rank), __popc(fm));
            int d = 0;
            for (int q = 0; q < 4; ++q) if ((fm >> q) & 1u) { if (pick == 0) { d = q; break; } --pick; }
            e.da[s] = (int16_t)(e.tx[s] + adj_dx(d)); e.db[s] = (int16_t)(e.ty[s] + adj_dy(d));
            ++mine;
        }
        nd = __reduce_add_sync(ZS_FULL, mine);
        __syncwarp();
#pragma unroll 1
        for (int s = lane; s < p.M; s += 32) if (e.dtype[s] == D_WANDER) e.dtype[s] = D_MOVE;
        __syncwarp();
    }
    // actions list in actor (dict) order (core.py:83-90): the position of an acting thing is its dict rank minus
    // the idle things before it (idle things are rare: usually the rank is the position).
    // Everything that cannot change before the actor acts is resolved here, in parallel.
    int cnt = 0, n_ah = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < p.M; s0 += 32) {
        const int s = s0 + lane;
        const int type = (s < p.M && (e.tm[s] & 0x80)) ? e.dtype[s] : D_IDLE;
        int pos = s < p.M ? e.rk[s] : 0;
        if (n_idle) {
            const int mine = pos;
#pragma unroll 1
            for (int j = 0; j < p.M; ++j) pos -= ((e.tm[j] & 0x80) && e.dtype[j] == D_IDLE && e.rk[j] < mine);
        }
        if (type == D_IDLE) continue;
        const int a = e.da[s], b = e.db[s], x = e.tx[s], y = e.ty[s];
        int kind = X_NOP, r2 = 0, dlo = 0, dn = 1;
        if (type == D_MOVE) {  // in bounds and at most one step (core.py:149-153); occupancy is checked when it runs
            if ((unsigned)a < (unsigned)p.W && (unsigned)b < (unsigned)p.H && dist2(x, y, a, b) <= 1) kind = X_MOVE;
        } else {
            const bool is_static = a >= p.M;
            int mx = 100;
            if (type == D_ATTACK) { const int w = e.tm[s] & 15; r2 = c_range2[w]; dlo = c_dmg_lo[w]; dn = c_dmg_n[w]; }
            else {  // heal: randint(MAX_LIFE // 10, MAX_LIFE // 4) of the target's class, range 3 (core.py:194-198)
                if (is_static) mx = max_life_of_label(__ldg(p.static_label + (a - p.M)));
                r2 = 9; dlo = mx / 10; dn = mx / 4 - mx / 10 + 1;
            }
            if (is_static) {
                const int cell = __ldg(p.static_cell + (a - p.M));
                const int gy = cell / p.W, gx = cell - gy * p.W;
                if (dist2(x, y, gx, gy) <= r2) kind = type == D_ATTACK ? X_ATTACK_S : X_HEAL_S;
            } else kind = type == D_ATTACK ? X_ATTACK_M : X_HEAL_M;
        }
        e.act[pos] = pack_action(s, kind, kind >= X_ATTACK_S ? a - p.M : a, b, r2, dlo, dn);
        ++cnt;
        n_ah += kind >= X_ATTACK_M;
    }
    const int L = __reduce_add_sync(ZS_FULL, cnt);
    n_ah = __reduce_add_sync(ZS_FULL, n_ah);
    // ---- draws of this step, generated 4 per lane (counter-based: any k is available directly)
    const int n_need = nd + (L > 1 ? L - 1 : 0) + n_ah;
#pragma unroll 1
    for (int blk = lane; blk * 4 < n_need; blk += 32) {
        uint32_t o[4];
        philox4x32_10(e.env_global, (uint32_t)e.episode, t_word, (uint32_t)blk, p.key0, p.key1, o);
        *(uint4*)(e.draws + 4 * blk) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncwarp();
    // Fisher-Yates partner of every iteration (random.shuffle: for i = L-1 .. 1: j = randbelow(i + 1))
#pragma unroll 1
    for (int i = 1 + lane; i < L; i += 32) e.dtype[i] = (uint8_t)below(e.draws[nd + (L - 1 - i)], i + 1);
    __syncwarp();

    // ---- random.shuffle + execute_actions: order-dependent by definition, run by lane 0
    int k = nd + (L > 1 ? L - 1 : 0);
    if (lane == 0) {
#pragma unroll 1
        for (int i = L - 1; i >= 1; --i) {
            const int j = e.dtype[i];
            const unsigned long long tmp = e.act[i]; e.act[i] = e.act[j]; e.act[j] = tmp;
        }
        int n_touched = 0, stamp = e.stampctr, fl = e.flags, deaths = e.deaths;
#pragma unroll 1
        for (int i = 0; i < L; ++i) {
            const unsigned long long pk = e.act[i];
            const int actor = (int)(pk & 0xff), kind = (int)((pk >> 8) & 7);
            const int a = (int16_t)(pk >> 11), b = (int16_t)(pk >> 27);
            if (kind == X_NOP) continue;
            if (kind == X_MOVE) {  // World.thing_move (core.py:140-166)
                const int c = b * p.W + a;
                if (!g_is_thing(e.grid[c])) {
                    const int old = e.ty[actor] * p.W + e.tx[actor];
                    e.grid[old] = dead_bit(e, old) ? G_DEAD : G_EMPTY;
                    e.grid[c] = (uint8_t)(actor + 1);
                    e.tx[actor] = (int16_t)a; e.ty[actor] = (int16_t)b;
                    e.ts[actor] = stamp++;  // things[dest] = thing; del things[old]: goes last
                }
                continue;
            }
            const int r2 = (int)((pk >> 43) & 127), dlo = (int)((pk >> 50) & 127), dn = (int)((pk >> 57) & 63);
            if (kind <= X_HEAL_M) {  // mobile target: distance between CURRENT positions (core.py:176,194)
                if (dist2(e.tx[actor], e.ty[actor], e.tx[a], e.ty[a]) > r2) continue;
                const int amount = dlo + below(e.draws[k++], dn);
                if (kind == X_ATTACK_M) e.tl[a] = (int16_t)(e.tl[a] - amount);
                else { const int nl = e.tl[a] + amount; e.tl[a] = (int16_t)(nl < 100 ? nl : 100); }
            } else {
                const int amount = dlo + below(e.draws[k++], dn);
                if (kind == X_ATTACK_S) {
                    e.sl[a] = (int16_t)(e.sl[a] - amount);
                    e.list[n_touched++] = (uint16_t)a;
                } else {
                    const int mx = dlo == 1 ? 10 : 200;  // MAX_LIFE // 10 is 1 for a Box, 20 for a Wall
                    const int nl = e.sl[a] + amount;
                    e.sl[a] = (int16_t)(nl < mx ? nl : mx);
                }
                e.grid[__ldg(p.static_cell + a)] = G_STATIC_DMG;
                fl |= FL_DMG | FL_SL_DIRTY;
            }
        }
        // clean_dead_things for boxes/walls hit this step (core.py:121-138); the full scan below
        // covers them on the first step of a world
        if (!(fl & FL_FRESH)) {
#pragma unroll 1
            for (int i = 0; i < n_touched; ++i) {
                const int si = e.list[i];
                const int cell = __ldg(p.static_cell + si);
                if (e.sl[si] <= 0 && g_is_static(e.grid[cell])) { e.grid[cell] = G_EMPTY; deaths++; }
            }
        }
        e.stampctr = stamp; e.flags = fl; e.deaths = deaths;
    }
    k = __shfl_sync(ZS_FULL, k, 0);
    e.stampctr = __shfl_sync(ZS_FULL, e.stampctr, 0);
    e.deaths = __shfl_sync(ZS_FULL, e.deaths, 0);
    e.flags = __shfl_sync(ZS_FULL, e.flags, 0);
    __syncwarp();

    // ---- clean_dead_things (core.py:121-138)
    int nd_all = 0, nd_z = 0;
    if (e.flags & FL_FRESH) {  // first step of this world: every box/wall with life <= 0 leaves now
        if (e.flags & FL_DMG) {
#pragma unroll 1
            for (int i = lane; i < p.S; i += 32) {
                if (e.sl[i] <= 0) {
                    const int cell = __ldg(p.static_cell + i);
                    if (g_is_static(e.grid[cell])) { e.grid[cell] = G_EMPTY; ++nd_all; }
                }
            }
        }
        e.flags &= ~FL_FRESH;
    }
#pragma unroll 1
    for (int s = lane; s < p.M; s += 32) {
        if ((e.tm[s] & 0x80) && e.tl[s] <= 0) {
            const int c = e.ty[s] * p.W + e.tx[s];
            e.grid[c] = G_DEAD;                       // DeadBody overwrites any decoration (core.py:30-31,126-128)
            atomicOr(&e.dead[c >> 5], 1u << (c & 31));
            e.tm[s] &= 0x7f;
            ++nd_all;
            nd_z += s >= NP;
        }
    }
    e.deaths += __reduce_add_sync(ZS_FULL, nd_all);
    e.zd += __reduce_add_sync(ZS_FULL, nd_z);
    __syncwarp();
    return k;
}

// ---------------------------------------------------------------- World.spawn_in_random (core.py:40-66)
// Places the `count` slots listed in e.list[0..count) on shuffled free cells of `spawn` (or of the
// whole map, x-major, when the map has no such spawn cells).  Only the first `count` Fisher-Yates
// iterations decide placements (spawns.pop() takes from the end); the rest of the shuffle only
// advances the draw counter.  Returns the new draw index.  Cold path: out of line.
__device__ __noinline__ int spawn_in_random(const ZsParams& p, unsigned char* base, int env, int lane, int episode,
                                            uint32_t t_word, int k, int count, int which, int stamp0) {
    Env e;
    env_bind(p, e, base, env, lane);
    e.episode = episode;
    const uint16_t* spawn = which ? p.zs_cells : p.ps_cells;
    const int n_spawn = which ? p.n_zs : p.n_ps;
    const int n_src = n_spawn > 0 ? n_spawn : p.cells;
    int n = 0;
#pragma unroll 1
    for (int b0 = 0; b0 < n_src; b0 += 32) {
        const int i = b0 + lane;
        bool ok = false;
        int c = 0;
        if (i < n_src) {
            if (n_spawn > 0) c = __ldg(spawn + i);
            else { int x = i / p.H; int y = i - x * p.H; c = y * p.W + x; }
            ok = !g_is_thing(e.grid[c]);
        }
        const unsigned m = __ballot_sync(ZS_FULL, ok);
        if (ok) e.cand[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)c;
        n += __popc(m);
    }
    __syncwarp();
    const int placed = count < n ? count : n;
#pragma unroll 1
    for (int it = lane; it < placed; it += 32) {
        const int i = n - 1 - it;
        e.draws[it] = i >= 1 ? (uint32_t)below(draw_at(p, e, t_word, k + it), i + 1) : 0u;
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll 1
        for (int it = 0; it < placed; ++it) {
            const int i = n - 1 - it;
            if (i >= 1) { const int j = (int)e.draws[it]; uint16_t tmp = e.cand[i]; e.cand[i] = e.cand[j]; e.cand[j] = tmp; }
            const int c = e.cand[i];
            const int s = e.list[it];
            const int y = c / p.W;
            e.tx[s] = (int16_t)(c - y * p.W); e.ty[s] = (int16_t)y;
            e.tm[s] |= 0x80;
            e.ts[s] = stamp0 + it;
            e.grid[c] = (uint8_t)(s + 1);
        }
        e.scal[ZS_S_STAMP_COUNTER] = stamp0 + placed;
    }
    __syncwarp();
    return k + (n > 1 ? n - 1 : 0);
}

// Game.spawn_zombies (game.py:189-194): `count` Zombie() constructions (life draws, things.py:62)
// followed by spawn_in_random on the zombie spawn cells; free zombie slots are taken in ascending order.
// The new stamp counter is left in e.scal[ZS_S_STAMP_COUNTER].
__device__ __noinline__ int spawn_zombies(const ZsParams& p, unsigned char* base, int env, int lane, int episode,
                                          uint32_t t_word, int k, int count, int stamp0) {
    Env e;
    env_bind(p, e, base, env, lane);
    e.episode = episode;
    const int NP = p.P + p.A;
    int n = 0;
#pragma unroll 1
    for (int b0 = NP; b0 < p.M; b0 += 32) {
        const int s = b0 + lane;
        const bool free_slot = s < p.M && !(e.tm[s] & 0x80);
        const unsigned m = __ballot_sync(ZS_FULL, free_slot);
        const int pos = n + __popc(m & ((1u << lane) - 1u));
        if (free_slot && pos < count) e.list[pos] = (uint16_t)s;
        n += __popc(m);
    }
    const int made = count < n ? count : n;
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < made; i += 32) {
        const int s = e.list[i];
        e.tl[s] = (int16_t)(50 + below(draw_at(p, e, t_word, k + i), 51));
        e.tm[s] = ZS_WEAPON_CLAWS;
    }
    __syncwarp();
    return spawn_in_random(p, base, env, lane, episode, t_word, k + count, made, 1, stamp0);
}

// Game.__initialize_world__ (game.py:151-169) + reward_tracker.reset (reward.py:26-28).  Out of line and
// with its own binding of the env's shared memory, so the hot loop's registers stay registers: the new
// scalars are left in e.scal (read back with scalars_from_smem).  `flags_in`: the launch-lifetime static
// damage flags survive a world init (the damage itself does, game.py:154-155).  Returns the draws consumed.
__device__ __noinline__ int initialize_world(const ZsParams& p, unsigned char* base, int env, int lane, int episode, int flags_in) {
    Env e;
    env_bind(p, e, base, env, lane);
    const int NP = p.P + p.A;
    const int flags = (flags_in & (FL_DMG | FL_SL_DIRTY)) | FL_FRESH;
    e.episode = episode;
#pragma unroll 1
    for (int w = lane; w < p.dead_words; w += 32) e.dead[w] = 0;
    int k = 0;
#pragma unroll 1
    for (int s = lane; s < p.Mp; s += 32) {
        int w = 0;
        if (s < p.P) w = ZS_WEAPON_SHOTGUN;            // terminator.py:41-42
        else if (s < NP) w = p.agent_weapons[s - p.P];
        else w = ZS_WEAPON_CLAWS;
        e.tm[s] = (uint8_t)(w == ZS_WEAPON_RANDOM ? 15 : w);
        if (s < NP) e.tl[s] = 100;
    }
    __syncwarp();
    // agent_weapon="random": one random.choice per agent, in agent order (weapons.py:43)
#pragma unroll 1
    for (int a = 0; a < p.A; ++a) {
        if (p.agent_weapons[a] == ZS_WEAPON_RANDOM) {
            const int pick = below(draw_at(p, e, 0u, k), 5);
            ++k;
            if (lane == 0) e.tm[p.P + a] = (uint8_t)(pick == 0 ? ZS_WEAPON_KNIFE : pick == 1 ? ZS_WEAPON_AXE
                                                     : pick == 2 ? ZS_WEAPON_GUN : pick == 3 ? ZS_WEAPON_RIFLE : ZS_WEAPON_SHOTGUN);
        }
    }
    __syncwarp();
    build_grid(p, base, env, lane, flags);  // every slot is out of the world here: statics (all present) only
#pragma unroll 1
    for (int s = lane; s < p.P; s += 32) e.list[s] = (uint16_t)s;
    if (lane == 0) e.scal[ZS_S_STAMP_COUNTER] = 0;
    __syncwarp();
    k = spawn_in_random(p, base, env, lane, episode, 0u, k, p.P, 0, 0);
#pragma unroll 1
    for (int a = lane; a < p.A; a += 32) e.list[a] = (uint16_t)(p.P + a);
    __syncwarp();
    k = spawn_in_random(p, base, env, lane, episode, 0u, k, p.A, 0, e.scal[ZS_S_STAMP_COUNTER]);
    k = spawn_zombies(p, base, env, lane, episode, 0u, k, p.initial_zombies, e.scal[ZS_S_STAMP_COUNTER]);
#pragma unroll 1
    for (int a = lane; a < p.A; a += 32) e.prev[a] = e.tl[p.P + a];
    if (lane == 0) {
        e.scal[ZS_S_T] = -1; e.scal[ZS_S_EPISODE] = episode; e.scal[ZS_S_DEATHS] = 0; e.scal[ZS_S_ZOMBIE_DEATHS] = 0;
        e.scal[ZS_S_FLAGS] = flags; e.scal[ZS_S_PREV_ZOMBIE_DEATHS] = 0; e.scal[ZS_S_EPISODE_STEPS] = 0;
    }
    __syncwarp();
    return k;
}

// ---------------------------------------------------------------- rules (zombsole/rules/*.py)
__device__ __forceinline__ void rules_eval(const ZsParams& p, Env& e, bool& ended, bool& won, bool& agents_alive) {
    const int lane = e.lane;
    const int NP = p.P + p.A;
    int alive = 0, ag = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < NP; s0 += 32) {
        const int s = s0 + lane;
        const bool al = s < NP && e.tl[s] > 0;
        alive += __popc(__ballot_sync(ZS_FULL, al));
        ag += __popc(__ballot_sync(ZS_FULL, al && s >= p.P));
    }
    agents_alive = ag > 0;                    // rules/rules.py:13-18
    const bool players_alive = alive > 0;     // rules/rules.py:6-11
    won = players_alive;
    if (p.rules == ZS_RULES_EXTERMINATION) {  // extermination.py:12-26
        bool z = false;
#pragma unroll 1
        for (int s = NP + lane; s < p.M; s += 32) z |= (e.tm[s] & 0x80) && e.tl[s] > 0;
        ended = !players_alive || !__any_sync(ZS_FULL, z);
    } else if (p.rules == ZS_RULES_SURVIVAL) {  // survival.py:5-7
        ended = !players_alive;
    } else if (p.rules == ZS_RULES_SAFEHOUSE) {  // safehouse.py:10-32
        bool out = false;
#pragma unroll 1
        for (int s = lane; s < NP; s += 32)
            out |= e.tl[s] > 0 && !objective_bit(p, e.ty[s] * p.W + e.tx[s]);
        ended = players_alive ? !__any_sync(ZS_FULL, out) : true;
    } else {  // evacuation.py:13-57: at least half the team alive and the living form one 4-connected cluster
        const bool half = 2 * alive >= NP;  // len(alive) >= len(all) / 2.0
        won = half;
        ended = true;
        if (half) {
            int together = 0;
            if (lane == 0) {
                unsigned long long seen = 0, pending = 0;
                int first = 0;
                while (e.tl[first] <= 0) ++first;
                pending = 1ull << first;
                while (pending) {
                    const int s = __ffsll((long long)pending) - 1;
                    pending &= pending - 1;
                    seen |= 1ull << s;
                    ++together;
                    const int x = e.tx[s], y = e.ty[s];
#pragma unroll 1
                    for (int d = 0; d < 4; ++d) {
                        const int g = grid_at(p, e, x + adj_dx(d), y + adj_dy(d));
                        if (g >= 1 && g <= NP && e.tl[g - 1] > 0 && !((seen | pending) >> (g - 1) & 1ull)) pending |= 1ull << (g - 1);
                    }
                }
            }
            together = __shfl_sync(ZS_FULL, together, 0);
            ended = together == alive;
        }
    }
}

__device__ __forceinline__ double total_reward(int zombie_deaths, int life_sum) {
    // reward.py:37-41 / 90-92: int + float, the division first
    return __dadd_rn(__int2double_rn(zombie_deaths), __ddiv_rn(__int2double_rn(life_sum), 100.0));
}
