/*
 * zs_oracle.c — TEST INFRASTRUCTURE.  CPU restatement of the reference's
 * reset/step/observation/reward path (jvstinian/libzombsole v0.13.2, pure Python).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  It is the checker, never the product: nothing under
 * libzombsole_b200/ links or loads it.
 *
 * PARITY PINNING: this file is checked bit-for-bit (slot positions, lives, dict order,
 * static lives, decorations, counters, observation tensors, float64 reward bits, flags and
 * the number of random draws) against traces of the UNMODIFIED reference run in the build
 * container under the injected draw stream (oracle/ref_harness.py), both live
 * (tests/test_oracle_vs_reference.py, needs /root/reference) and from the committed
 * committed npz fixtures under tests/golden/ (tests/test_oracle_golden.py).
 *
 * The restatement deliberately mirrors the reference's data model rather than the CUDA
 * kernels': World.things is an insertion-ordered dict keyed by position, kept here as an
 * explicit order list plus a position->thing grid; a successful move re-inserts the mover
 * at the end of the order (zombsole/core.py:158-159).  Every function cites the reference
 * lines it follows (paths relative to the reference tree).
 *
 * Randomness: CPython's random.shuffle/randint/choice all reduce to Random._randbelow(n)
 * (Lib/random.py, CPython 3.12.3).  Draw k of cell (env, episode, t_word) is
 * (Philox4x32-10(key=seed, ctr=(env, episode, t_word, k>>2))[k&3] * n) >> 32.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/zs_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define ZSO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Philox */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

ZSO_EXPORT void zso_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------ types */
enum { T_BOX = 0, T_WALL, T_ZOMBIE, T_TERMINATOR, T_AGENT, T_SNIPER, T_TROLL, T_HAMSTER, T_RANDOMAN };
enum { A_MOVE = 1, A_ATTACK = 2, A_HEAL = 3 };

typedef struct Thing {
    int16_t x, y;
    int32_t life;
    uint8_t type;     /* T_* */
    uint8_t weapon;   /* ZS_WEAPON_* code, 0 for statics */
    uint8_t in_world; /* present in World.things */
    int32_t agent_index; /* for T_AGENT */
} Thing;

typedef struct Action {
    int32_t actor;
    int32_t type;       /* A_* */
    int32_t target;     /* thing id for attack/heal */
    int32_t dx, dy;     /* destination for move */
} Action;

typedef struct Env {
    Thing* things;      /* [S + M]: statics first (file order), then slots */
    int32_t* order;     /* World.things insertion order (thing ids) */
    int32_t n_order;
    int32_t* grid;      /* cell -> thing id, -1 = none (World.things keyed by position) */
    uint8_t* deco;      /* cell -> 0 | ZS_LABEL_DEAD_BODY | ZS_LABEL_OBJECTIVE (World.decoration) */
    int32_t t, deaths, zombie_deaths, episode, episode_steps;
    int32_t* tracker_life; /* reward tracker agents_life */
    int32_t tracker_zd;
    /* draw stream */
    uint32_t cell[3];
    int32_t k;
    uint32_t cache[4];
    int32_t cache_block;
    int32_t last_draws;  /* draws consumed by the last world init */
    int32_t step_draws;  /* draws consumed by the last step (incl. minimum-zombie respawn) */
    Action* actions;    /* scratch [M] */
    int32_t* scratch;   /* scratch [max(cells, M)] */
} Env;

typedef struct ZsoHandle {
    ZsConfig cfg;
    int32_t W, H, cells, S, M, A, P, Z;
    int16_t* static_xy; uint8_t* static_label;
    int32_t n_ps, n_zs, n_obj;
    int16_t* ps_xy; int16_t* zs_xy;
    uint8_t* objective; /* cell -> 1 if objective */
    uint32_t key[2];
    int32_t obs_C, obs_H, obs_W, obs_count;
    int64_t obs_elems;
    Env* envs;
    int64_t stats[4];
} ZsoHandle;

static int weapon_range2(int w) {
    /* max_range squared, floored: 1.5 -> 2, 3 -> 9, 6 -> 36, 10 -> 100 (weapons.py:18-25);
       distance() is sqrt(dx^2+dy^2) (utils.py:13-20) so "d > range" <=> d2 > floor(range^2) */
    switch (w) {
        case ZS_WEAPON_CLAWS: case ZS_WEAPON_KNIFE: case ZS_WEAPON_AXE: return 2;
        case ZS_WEAPON_GUN: return 36;
        case ZS_WEAPON_RIFLE: return 100;
        case ZS_WEAPON_SHOTGUN: return 9;
    }
    return 0;
}
static void weapon_damage(int w, int* lo, int* hi) {
    switch (w) {
        case ZS_WEAPON_CLAWS: case ZS_WEAPON_KNIFE: *lo = 5; *hi = 10; return;
        case ZS_WEAPON_AXE: case ZS_WEAPON_SHOTGUN: *lo = 75; *hi = 100; return;
        case ZS_WEAPON_GUN: *lo = 10; *hi = 50; return;
        case ZS_WEAPON_RIFLE: *lo = 25; *hi = 75; return;
    }
    *lo = 0; *hi = 0;
}
static int max_life(int type) {
    /* things.py:12,47,57,109 */
    switch (type) { case T_BOX: return 10; case T_WALL: return 200; default: return 100; }
}
static int is_player(int type) { return type == T_TERMINATOR || type >= T_AGENT; }
static int is_fighter(int type) { return type >= T_ZOMBIE; }

/* ------------------------------------------------------------------ draws */
static void draws_begin(Env* e, uint32_t env_index, uint32_t episode, uint32_t t_word) {
    e->cell[0] = env_index; e->cell[1] = episode; e->cell[2] = t_word;
    e->k = 0; e->cache_block = -1;
}
static uint32_t randbelow(const ZsoHandle* h, Env* e, uint32_t n) {
    int block = e->k >> 2;
    if (block != e->cache_block) {
        uint32_t ctr[4] = { e->cell[0], e->cell[1], e->cell[2], (uint32_t)block };
        philox4x32_10(ctr, h->key, e->cache);
        e->cache_block = block;
    }
    uint32_t u = e->cache[e->k & 3];
    e->k++;
    return (uint32_t)(((uint64_t)u * n) >> 32);
}
/* random.randint(lo, hi) = lo + _randbelow(hi - lo + 1) */
static int randint(const ZsoHandle* h, Env* e, int lo, int hi) { return lo + (int)randbelow(h, e, (uint32_t)(hi - lo + 1)); }

/* ------------------------------------------------------------------ world helpers */
static int in_bounds(const ZsoHandle* h, int x, int y) {
    /* core.py:204-208 */
    return x >= 0 && x < h->W && y >= 0 && y < h->H;
}
static int thing_at(const ZsoHandle* h, const Env* e, int x, int y) {
    /* World.things.get(position): positions outside the map hold nothing */
    if (!in_bounds(h, x, y)) return -1;
    return e->grid[y * h->W + x];
}
static int dist2(int x1, int y1, int x2, int y2) {
    int dx = x1 - x2, dy = y1 - y2;
    return dx * dx + dy * dy;
}
static void order_append(Env* e, int id) { e->order[e->n_order++] = id; }
static void order_remove(Env* e, int id) {
    int i = 0;
    while (i < e->n_order && e->order[i] != id) ++i;
    for (; i + 1 < e->n_order; ++i) e->order[i] = e->order[i + 1];
    e->n_order--;
}
/* World.spawn_thing for a non-decoration (core.py:24-38) */
static void world_insert(const ZsoHandle* h, Env* e, int id) {
    Thing* t = &e->things[id];
    e->grid[t->y * h->W + t->x] = id;
    t->in_world = 1;
    order_append(e, id);
}

static const int ADJ_DX[4] = { 0, 0, 1, -1 }; /* utils.py:34-44 */
static const int ADJ_DY[4] = { 1, -1, 0, 0 };

/* closest(something, others): sorted() is stable => first minimum in list order (utils.py:23-31) */
static int closest_of_type(const ZsoHandle* h, const Env* e, int self, int want_player, int want_zombie, int exclude_self) {
    (void)h;
    const Thing* me = &e->things[self];
    int best = -1, best_d = 0;
    for (int i = 0; i < e->n_order; ++i) {
        int id = e->order[i];
        const Thing* t = &e->things[id];
        if (exclude_self && id == self) continue;
        if (!((want_player && is_player(t->type)) || (want_zombie && t->type == T_ZOMBIE))) continue;
        int d = dist2(me->x, me->y, t->x, t->y);
        if (best < 0 || d < best_d) { best = id; best_d = d; }
    }
    return best;
}

/* ------------------------------------------------------------------ decide phase */
/* Zombie.next_step (things.py:70-105) */
static int zombie_next_step(const ZsoHandle* h, Env* e, int self, Action* out) {
    const Thing* me = &e->things[self];
    int px[4], py[4], npos = 0;
    for (int a = 0; a < 4; ++a) { /* possible_moves: no bounds check (utils.py:47-52) */
        int x = me->x + ADJ_DX[a], y = me->y + ADJ_DY[a];
        if (thing_at(h, e, x, y) < 0) { px[npos] = x; py[npos] = y; ++npos; }
    }
    int target = closest_of_type(h, e, self, 1, 0, 0);
    out->actor = self;
    if (target >= 0) {
        const Thing* tg = &e->things[target];
        if (dist2(me->x, me->y, tg->x, tg->y) <= 2) { /* distance < 1.5 (things.py:83) */
            out->type = A_ATTACK; out->target = target; return 1;
        }
        if (npos) {
            int best = 0, best_d = dist2(tg->x, tg->y, px[0], py[0]);
            for (int i = 1; i < npos; ++i) {
                int d = dist2(tg->x, tg->y, px[i], py[i]);
                if (d < best_d) { best = i; best_d = d; }
            }
            out->type = A_MOVE; out->dx = px[best]; out->dy = py[best]; return 1;
        }
        /* blocked: adjacent cells stably sorted by distance to the target; first Box/Wall (things.py:93-99) */
        int idx[4] = { 0, 1, 2, 3 }, key[4];
        for (int a = 0; a < 4; ++a) key[a] = dist2(tg->x, tg->y, me->x + ADJ_DX[a], me->y + ADJ_DY[a]);
        for (int i = 1; i < 4; ++i) { /* stable insertion sort */
            int v = idx[i], j = i;
            while (j > 0 && key[idx[j - 1]] > key[v]) { idx[j] = idx[j - 1]; --j; }
            idx[j] = v;
        }
        for (int i = 0; i < 4; ++i) {
            int id = thing_at(h, e, me->x + ADJ_DX[idx[i]], me->y + ADJ_DY[idx[i]]);
            if (id >= 0 && (e->things[id].type == T_BOX || e->things[id].type == T_WALL)) {
                out->type = A_ATTACK; out->target = id; return 1;
            }
        }
        return 0;
    }
    if (npos) { /* no humans: wander, random.choice(positions) (things.py:101-103) */
        int i = (int)randbelow(h, e, (uint32_t)npos);
        out->type = A_MOVE; out->dx = px[i]; out->dy = py[i]; return 1;
    }
    return 0;
}

/* Terminator.next_step (players/terminator.py:9-37) */
static int terminator_next_step(const ZsoHandle* h, Env* e, int self, Action* out) {
    const Thing* me = &e->things[self];
    out->actor = self;
    int target = closest_of_type(h, e, self, 0, 1, 0);
    if (target < 0) { out->type = A_HEAL; out->target = self; return 1; }
    const Thing* tg = &e->things[target];
    if (dist2(me->x, me->y, tg->x, tg->y) > weapon_range2(me->weapon)) {
        int best = 0, best_d = dist2(tg->x, tg->y, me->x + ADJ_DX[0], me->y + ADJ_DY[0]);
        for (int a = 1; a < 4; ++a) {
            int d = dist2(tg->x, tg->y, me->x + ADJ_DX[a], me->y + ADJ_DY[a]);
            if (d < best_d) { best = a; best_d = d; }
        }
        int bx = me->x + ADJ_DX[best], by = me->y + ADJ_DY[best];
        int ob = thing_at(h, e, bx, by);
        if (ob >= 0) {
            out->type = is_player(e->things[ob].type) ? A_HEAL : A_ATTACK;
            out->target = ob; return 1;
        }
        out->type = A_MOVE; out->dx = bx; out->dy = by; return 1;
    }
    out->type = A_ATTACK; out->target = target; return 1;
}

/* Sniper.next_step (players/sniper.py:9-19): shoot the closest zombie wherever it is, idle without zombies */
static int sniper_next_step(const ZsoHandle* h, Env* e, int self, Action* out) {
    out->actor = self;
    int target = closest_of_type(h, e, self, 0, 1, 0);
    if (target < 0) return 0;
    out->type = A_ATTACK; out->target = target; return 1;
}
/* Troll.next_step (players/troll.py:10-12) */
static int troll_next_step(int self, Action* out) {
    out->actor = self; out->type = A_HEAL; out->target = self; return 1;
}
/* Hamster.next_step (players/hamster.py:10-14): random.choice(possible_moves) */
static int hamster_next_step(const ZsoHandle* h, Env* e, int self, Action* out) {
    const Thing* me = &e->things[self];
    int px[4], py[4], npos = 0;
    for (int a = 0; a < 4; ++a) {
        int x = me->x + ADJ_DX[a], y = me->y + ADJ_DY[a];
        if (thing_at(h, e, x, y) < 0) { px[npos] = x; py[npos] = y; ++npos; }
    }
    if (!npos) return 0;
    int i = (int)randbelow(h, e, (uint32_t)npos);
    out->actor = self; out->type = A_MOVE; out->dx = px[i]; out->dy = py[i]; return 1;
}

/* Agent.next_step (players/agent.py:28-96) */
static int agent_next_step(const ZsoHandle* h, Env* e, int self, const int32_t act[3], Action* out) {
    const Thing* me = &e->things[self];
    out->actor = self;
    int type = act[0], dx = act[1], dy = act[2];
    if (type == ZS_ACT_ABSENT) { type = ZS_ACT_HEAL; dx = 0; dy = 0; } /* multiagent_env.py:129-131 */
    switch (type) {
        case ZS_ACT_MOVE:
            out->type = A_MOVE; out->dx = me->x + dx; out->dy = me->y + dy; return 1;
        case ZS_ACT_ATTACK_CLOSEST: {
            int target = closest_of_type(h, e, self, 0, 1, 0);
            if (target < 0) return 0;
            out->type = A_ATTACK; out->target = target; return 1;
        }
        case ZS_ACT_ATTACK: {
            int target = thing_at(h, e, me->x + dx, me->y + dy);
            if (target < 0) return 0;
            out->type = A_ATTACK; out->target = target; return 1;
        }
        case ZS_ACT_HEAL: {
            if (dx == 0 && dy == 0) { out->type = A_HEAL; out->target = self; return 1; }
            int target = thing_at(h, e, me->x + dx, me->y + dy);
            if (target < 0) return 0;
            int tt = e->things[target].type;
            if (!(is_player(tt) || tt == T_BOX || tt == T_WALL)) return 0;
            out->type = A_HEAL; out->target = target; return 1;
        }
        case ZS_ACT_HEAL_CLOSEST: {
            int target = closest_of_type(h, e, self, 1, 0, 1);
            out->type = A_HEAL; out->target = target >= 0 ? target : self; return 1;
        }
        default: return 0;
    }
}

/* ------------------------------------------------------------------ execute phase */
/* World.thing_move (core.py:140-166) */
static void thing_move(const ZsoHandle* h, Env* e, int actor, int x, int y) {
    Thing* t = &e->things[actor];
    if (!in_bounds(h, x, y)) return;
    if (e->grid[y * h->W + x] >= 0) return;
    if (dist2(t->x, t->y, x, y) > 1) return; /* distance > 1 */
    e->grid[y * h->W + x] = actor;
    e->grid[t->y * h->W + t->x] = -1;
    t->x = (int16_t)x; t->y = (int16_t)y;
    order_remove(e, actor); /* things[dest] = thing; del things[old] => goes to the end */
    order_append(e, actor);
}
/* World.thing_attack (core.py:168-184) */
static void thing_attack(const ZsoHandle* h, Env* e, int actor, int target) {
    Thing* a = &e->things[actor]; Thing* g = &e->things[target];
    if (dist2(a->x, a->y, g->x, g->y) > weapon_range2(a->weapon)) return;
    int lo, hi; weapon_damage(a->weapon, &lo, &hi);
    g->life -= randint(h, e, lo, hi);
}
/* World.thing_heal (core.py:186-202) */
static void thing_heal(const ZsoHandle* h, Env* e, int actor, int target) {
    Thing* a = &e->things[actor]; Thing* g = &e->things[target];
    if (dist2(a->x, a->y, g->x, g->y) > 9) return; /* HEALING_RANGE = 3 */
    int mx = max_life(g->type);
    int heal = randint(h, e, mx / 10, mx / 4);
    int nl = g->life + heal;
    g->life = nl < mx ? nl : mx;
}

/* World.clean_dead_things (core.py:121-138) */
static void clean_dead_things(const ZsoHandle* h, Env* e) {
    int n = 0;
    int32_t* dead = e->scratch;
    for (int i = 0; i < e->n_order; ++i) if (e->things[e->order[i]].life <= 0) dead[n++] = e->order[i];
    for (int i = 0; i < n; ++i) {
        Thing* t = &e->things[dead[i]];
        int cell = t->y * h->W + t->x;
        if (is_fighter(t->type)) e->deco[cell] = ZS_LABEL_DEAD_BODY; /* overwrites any decoration (core.py:30-31) */
        e->grid[cell] = -1;
        t->in_world = 0;
        order_remove(e, dead[i]);
        e->deaths++;
        if (t->type == T_ZOMBIE) e->zombie_deaths++;
    }
}

/* ------------------------------------------------------------------ spawning */
/* World.spawn_in_random (core.py:40-66).  ids: things to place in order; returns how many were placed. */
static int spawn_in_random(const ZsoHandle* h, Env* e, const int32_t* ids, int n_ids, const int16_t* spawn_xy, int n_spawn) {
    int32_t* cand = e->scratch; /* cell indices as y*W+x */
    int n = 0;
    if (n_spawn == 0) { /* "if not possible_positions": every cell, x-major */
        for (int x = 0; x < h->W; ++x) for (int y = 0; y < h->H; ++y)
            if (e->grid[y * h->W + x] < 0) cand[n++] = y * h->W + x;
    } else {
        for (int i = 0; i < n_spawn; ++i) {
            int x = spawn_xy[2 * i], y = spawn_xy[2 * i + 1];
            if (e->grid[y * h->W + x] < 0) cand[n++] = y * h->W + x;
        }
    }
    /* random.shuffle: for i in reversed(range(1, len)): j = randbelow(i + 1); swap */
    for (int i = n - 1; i >= 1; --i) {
        int j = (int)randbelow(h, e, (uint32_t)(i + 1));
        int32_t tmp = cand[i]; cand[i] = cand[j]; cand[j] = tmp;
    }
    int placed = 0;
    for (int i = 0; i < n_ids; ++i) {
        if (n == 0) break; /* fail_if_cant=False: return; for players the reference raises */
        int cell = cand[--n]; /* spawns.pop() */
        Thing* t = &e->things[ids[i]];
        t->x = (int16_t)(cell % h->W); t->y = (int16_t)(cell / h->W);
        world_insert(h, e, ids[i]);
        ++placed;
    }
    return placed;
}

/* Game.spawn_zombies (game.py:189-194): lives are drawn before the shuffle (things.py:62) */
static void spawn_zombies(const ZsoHandle* h, Env* e, int count) {
    int32_t ids[ZS_MAX_SLOTS];
    int n = 0;
    int base = h->S + h->P + h->A;
    for (int s = 0; s < h->Z && n < count; ++s) if (!e->things[base + s].in_world) ids[n++] = base + s;
    for (int i = 0; i < n; ++i) {
        Thing* z = &e->things[ids[i]];
        z->life = randint(h, e, 50, 100);
        z->type = T_ZOMBIE; z->weapon = ZS_WEAPON_CLAWS;
    }
    spawn_in_random(h, e, ids, n, h->zs_xy, h->n_zs);
}

/* Game.__initialize_world__ (game.py:151-169) + reward_tracker.reset (reward.py:26-28) */
static void initialize_world(const ZsoHandle* h, Env* e, int env_local, int episode) {
    e->episode = episode;
    draws_begin(e, (uint32_t)(h->cfg.env_index_base + env_local), (uint32_t)episode, 0);
    e->t = -1; e->deaths = 0; e->zombie_deaths = 0; e->episode_steps = 0;
    e->n_order = 0;
    for (int c = 0; c < h->cells; ++c) { e->grid[c] = -1; e->deco[c] = h->objective[c] ? ZS_LABEL_OBJECTIVE : 0; }
    for (int i = 0; i < h->S; ++i) world_insert(h, e, i); /* same objects: life persists (game.py:154-155) */
    int32_t ids[ZS_MAX_SLOTS];
    /* bots (create_player, game.py:157-159): terminator carries a Shotgun (terminator.py:41-42), sniper a Rifle
       (sniper.py:23-24); troll, hamster and randoman are created without a weapon and Player.__init__ draws
       random.choice([Gun, Shotgun, Rifle, Knife, Axe]) (things.py:115-116) */
    for (int b = 0; b < h->P; ++b) {
        Thing* t = &e->things[h->S + b];
        t->life = 100; t->in_world = 0;
        switch (h->cfg.bot_kinds[b]) {
            case ZS_KIND_SNIPER: t->type = T_SNIPER; t->weapon = ZS_WEAPON_RIFLE; break;
            case ZS_KIND_TROLL: case ZS_KIND_HAMSTER: case ZS_KIND_RANDOMAN: {
                static const uint8_t choices[5] = { ZS_WEAPON_GUN, ZS_WEAPON_SHOTGUN, ZS_WEAPON_RIFLE, ZS_WEAPON_KNIFE, ZS_WEAPON_AXE };
                t->type = h->cfg.bot_kinds[b] == ZS_KIND_TROLL ? T_TROLL : h->cfg.bot_kinds[b] == ZS_KIND_HAMSTER ? T_HAMSTER : T_RANDOMAN;
                t->weapon = choices[randbelow(h, e, 5)];
                break;
            }
            default: t->type = T_TERMINATOR; t->weapon = ZS_WEAPON_SHOTGUN; break;
        }
        ids[b] = h->S + b;
    }
    /* agents (create_agent -> WeaponFactory, weapons.py:28-45); "random" draws here */
    for (int a = 0; a < h->A; ++a) {
        Thing* t = &e->things[h->S + h->P + a];
        t->type = T_AGENT; t->life = 100; t->in_world = 0; t->agent_index = a;
        int w = h->cfg.agent_weapons[a];
        if (w == ZS_WEAPON_RANDOM) {
            static const uint8_t choices[5] = { ZS_WEAPON_KNIFE, ZS_WEAPON_AXE, ZS_WEAPON_GUN, ZS_WEAPON_RIFLE, ZS_WEAPON_SHOTGUN };
            w = choices[randbelow(h, e, 5)];
        }
        t->weapon = (uint8_t)w;
    }
    for (int z = 0; z < h->Z; ++z) e->things[h->S + h->P + h->A + z].in_world = 0;
    spawn_in_random(h, e, ids, h->P, h->ps_xy, h->n_ps);
    for (int a = 0; a < h->A; ++a) ids[a] = h->S + h->P + a;
    spawn_in_random(h, e, ids, h->A, h->ps_xy, h->n_ps);
    spawn_zombies(h, e, h->cfg.initial_zombies);
    for (int a = 0; a < h->A; ++a) e->tracker_life[a] = e->things[h->S + h->P + a].life;
    e->tracker_zd = 0;
    e->last_draws = e->k;
}

/* ------------------------------------------------------------------ observation */
static int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }

static void cell_codes(const ZsoHandle* h, const Env* e, int x, int y, int* label, int* life, int* weapon, int* agent_index) {
    /* observation.py:36-81: things over decorations; out of bounds = a fresh Wall */
    *agent_index = -1;
    if (!in_bounds(h, x, y)) { *label = ZS_LABEL_WALL; *life = 200; *weapon = 0; return; }
    int cell = y * h->W + x;
    int id = e->grid[cell];
    if (id >= 0) {
        const Thing* t = &e->things[id];
        *life = t->life; *weapon = t->weapon;
        switch (t->type) {
            case T_BOX: *label = ZS_LABEL_BOX; break;
            case T_WALL: *label = ZS_LABEL_WALL; break;
            case T_ZOMBIE: *label = ZS_LABEL_ZOMBIE; break;
            case T_AGENT: *label = ZS_LABEL_AGENT; *agent_index = t->agent_index; break;
            default: *label = ZS_LABEL_PLAYER; break;
        }
        return;
    }
    *label = e->deco[cell]; *life = 0; *weapon = 0;
}
static int32_t encode_simple(const ZsoHandle* h, const Env* e, int x, int y) {
    int label, life, weapon, ai;
    cell_codes(h, e, x, y, &label, &life, &weapon, &ai);
    if (label == 0) return 0;
    int adj = life < 100 ? life : 100;
    return 256 * label + 16 * weapon + floordiv(15 * adj, 100); /* observation.py:47-53 */
}
static void encode_channels(const ZsoHandle* h, const Env* e, int x, int y, int32_t out[3]) {
    int label, life, weapon, ai;
    cell_codes(h, e, x, y, &label, &life, &weapon, &ai);
    if (label == ZS_LABEL_AGENT) label = 8 + h->cfg.agent_obs_ids[ai]; /* observation.py:73-74 */
    out[0] = label; out[1] = life; out[2] = weapon;
}
static void encode_window(const ZsoHandle* h, const Env* e, int x0, int y0, int w, int hh, int32_t* obs) {
    /* [C, hh, w] with rows = y, columns = x (observation.py:83-119) */
    int plane = w * hh;
    for (int r = 0; r < hh; ++r) for (int c = 0; c < w; ++c) {
        if (h->cfg.obs_encoding == ZS_OBS_SIMPLE) obs[r * w + c] = encode_simple(h, e, x0 + c, y0 + r);
        else {
            int32_t v[3]; encode_channels(h, e, x0 + c, y0 + r, v);
            obs[r * w + c] = v[0]; obs[plane + r * w + c] = v[1]; obs[2 * plane + r * w + c] = v[2];
        }
    }
}
static void encode_obs(const ZsoHandle* h, const Env* e, int32_t* obs) {
    if (h->cfg.obs_scope == ZS_OBS_WORLD) { encode_window(h, e, 0, 0, h->W, h->H, obs); return; }
    int w = h->cfg.surroundings_width, half = w / 2;
    int per = h->obs_C * w * w;
    for (int a = 0; a < h->obs_count; ++a) {
        const Thing* ag = &e->things[h->S + h->P + a]; /* agent position, stale if dead (multiagent_env.py:88-97) */
        encode_window(h, e, ag->x - half, ag->y - half, w, w, obs + (int64_t)a * per);
    }
}

/* ------------------------------------------------------------------ rules */
static int players_alive(const ZsoHandle* h, const Env* e) { /* rules/rules.py:6-11 */
    for (int i = 0; i < h->P + h->A; ++i) if (e->things[h->S + i].life > 0) return 1;
    return 0;
}
static int agents_alive(const ZsoHandle* h, const Env* e) { /* rules/rules.py:13-18 */
    for (int a = 0; a < h->A; ++a) if (e->things[h->S + h->P + a].life > 0) return 1;
    return 0;
}
static int half_team_alive(const ZsoHandle* h, const Env* e) { /* evacuation.py:39-42 */
    int alive = 0, team = h->P + h->A;
    for (int i = 0; i < team; ++i) if (e->things[h->S + i].life > 0) ++alive;
    return (double)alive >= (double)team / 2.0;
}
static int alive_players_together(const ZsoHandle* h, const Env* e) { /* evacuation.py:18-37 */
    int team = h->P + h->A, ids[ZS_MAX_BOTS + ZS_MAX_AGENTS], n = 0;
    for (int i = 0; i < team; ++i) if (e->things[h->S + i].life > 0) ids[n++] = h->S + i;
    /* players_by_pos = dict(position -> player): for duplicate positions the later player wins */
    int seen[ZS_MAX_BOTS + ZS_MAX_AGENTS] = { 0 }, stack[ZS_MAX_BOTS + ZS_MAX_AGENTS * 8], sp = 0, together = 0;
    int stack_cap = (int)(sizeof(stack) / sizeof(stack[0]));
    stack[sp++] = 0;
    while (sp > 0) {
        int p = stack[--sp];
        if (!seen[p]) { seen[p] = 1; ++together; }
        const Thing* t = &e->things[ids[p]];
        for (int a = 0; a < 4; ++a) {
            int x = t->x + ADJ_DX[a], y = t->y + ADJ_DY[a], hit = -1;
            for (int q = 0; q < n; ++q) if (e->things[ids[q]].x == x && e->things[ids[q]].y == y) hit = q;
            if (hit >= 0 && !seen[hit] && sp < stack_cap) stack[sp++] = hit;
        }
    }
    return together == n;
}
static int game_ended(const ZsoHandle* h, const Env* e) {
    switch (h->cfg.rules) {
        case ZS_RULES_EXTERMINATION: { /* extermination.py:12-21 */
            int zombies = 0;
            for (int i = 0; i < e->n_order; ++i) {
                const Thing* t = &e->things[e->order[i]];
                if (t->type == T_ZOMBIE && t->life > 0) { zombies = 1; break; }
            }
            return !players_alive(h, e) || !zombies;
        }
        case ZS_RULES_SURVIVAL: return !players_alive(h, e); /* survival.py:5-7 */
        case ZS_RULES_EVACUATION: return half_team_alive(h, e) ? alive_players_together(h, e) : 1; /* evacuation.py:44-49 */
        default: { /* safehouse.py:10-26 */
            if (!players_alive(h, e)) return 1;
            for (int i = 0; i < h->P + h->A; ++i) {
                const Thing* t = &e->things[h->S + i];
                if (t->life > 0 && !(in_bounds(h, t->x, t->y) && h->objective[t->y * h->W + t->x])) return 0;
            }
            return 1;
        }
    }
}
static int game_won(const ZsoHandle* h, const Env* e) {
    if (h->cfg.rules == ZS_RULES_EVACUATION) return half_team_alive(h, e); /* evacuation.py:51-56 */
    return players_alive(h, e);
}

/* ------------------------------------------------------------------ step */
static const int32_t DISCRETE_ACTIONS[7][3] = {
    /* gym_env.py:328-351, multiagent_env.py:259-285 */
    { ZS_ACT_MOVE, 0, 1 }, { ZS_ACT_MOVE, -1, 0 }, { ZS_ACT_MOVE, 0, -1 }, { ZS_ACT_MOVE, 1, 0 },
    { ZS_ACT_ATTACK_CLOSEST, 0, 0 }, { ZS_ACT_HEAL, 0, 0 }, { ZS_ACT_HEAL_CLOSEST, 0, 0 },
};

/* RandoMan.next_step (players/randoman.py:9-21): random.choice(('move', 'attack', 'heal')); attack / heal take
   random.choice(list(things.values())) — ANY thing in World.things, boxes and walls included, in dict order; move
   changes one coordinate (random.choice((0, 1))) by random.choice((-1, 1)), drawn in that order */
static int randoman_next_step(const ZsoHandle* h, Env* e, int self, Action* out) {
    const Thing* me = &e->things[self];
    out->actor = self;
    int action = (int)randbelow(h, e, 3);
    if (action != 0) {
        out->type = action == 1 ? A_ATTACK : A_HEAL;
        out->target = e->order[randbelow(h, e, (uint32_t)e->n_order)];
    } else {
        int axis = (int)randbelow(h, e, 2);
        int sign = randbelow(h, e, 2) ? 1 : -1;
        out->type = A_MOVE;
        out->dx = me->x + (axis == 0 ? sign : 0);
        out->dy = me->y + (axis == 1 ? sign : 0);
    }
    return 1;
}

static void env_step(ZsoHandle* h, Env* e, int env_local, const int32_t* actions, int fmt, int32_t* obs,
                     double* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* agent_mask, int force_auto_reset) {
    const int A = h->A;
    int32_t acts[ZS_MAX_AGENTS][3];
    for (int a = 0; a < A; ++a) {
        if (fmt == ZS_ACTIONS_DISCRETE) {
            int id = actions[a];
            int n = h->cfg.obs_per_agent ? 7 : 6;
            if (id >= 0 && id < n) memcpy(acts[a], DISCRETE_ACTIONS[id], sizeof(acts[a]));
            else if (id < 0 && h->cfg.obs_per_agent) { acts[a][0] = ZS_ACT_ABSENT; acts[a][1] = acts[a][2] = 0; }
            else { acts[a][0] = ZS_ACT_NONE; acts[a][1] = acts[a][2] = 0; }
        } else memcpy(acts[a], actions + 3 * a, sizeof(acts[a]));
    }
    uint8_t alive_before[ZS_MAX_AGENTS];
    for (int a = 0; a < A; ++a) alive_before[a] = e->things[h->S + h->P + a].life > 0;

    /* World.step (core.py:72-78) */
    e->t += 1;
    draws_begin(e, (uint32_t)(h->cfg.env_index_base + env_local), (uint32_t)e->episode, (uint32_t)(e->t + 1));
    /* get_actions (core.py:80-101): actors in dict order */
    int n_act = 0;
    int n_actors = 0;
    int32_t* actors = e->scratch;
    for (int i = 0; i < e->n_order; ++i) if (is_fighter(e->things[e->order[i]].type)) actors[n_actors++] = e->order[i];
    for (int i = 0; i < n_actors; ++i) {
        int id = actors[i];
        Action act; memset(&act, 0, sizeof(act));
        int ok;
        switch (e->things[id].type) {
            case T_ZOMBIE: ok = zombie_next_step(h, e, id, &act); break;
            case T_TERMINATOR: ok = terminator_next_step(h, e, id, &act); break;
            case T_SNIPER: ok = sniper_next_step(h, e, id, &act); break;
            case T_TROLL: ok = troll_next_step(id, &act); break;
            case T_HAMSTER: ok = hamster_next_step(h, e, id, &act); break;
            case T_RANDOMAN: ok = randoman_next_step(h, e, id, &act); break;
            default: ok = agent_next_step(h, e, id, acts[e->things[id].agent_index], &act); break;
        }
        if (ok) e->actions[n_act++] = act;
    }
    /* random.shuffle(actions) (core.py:76) */
    for (int i = n_act - 1; i >= 1; --i) {
        int j = (int)randbelow(h, e, (uint32_t)(i + 1));
        Action tmp = e->actions[i]; e->actions[i] = e->actions[j]; e->actions[j] = tmp;
    }
    /* execute_actions (core.py:103-119) */
    for (int i = 0; i < n_act; ++i) {
        const Action* a = &e->actions[i];
        if (a->type == A_MOVE) thing_move(h, e, a->actor, a->dx, a->dy);
        else if (a->type == A_ATTACK) thing_attack(h, e, a->actor, a->target);
        else thing_heal(h, e, a->actor, a->target);
    }
    clean_dead_things(h, e);
    e->episode_steps += 1;

    /* reward tracker update (reward.py:30-41, 77-92) */
    double rew[ZS_MAX_AGENTS];
    if (!h->cfg.obs_per_agent) {
        int sum_prev = 0, sum_new = 0;
        for (int a = 0; a < A; ++a) sum_prev += e->tracker_life[a];
        double prev = (double)e->tracker_zd + (double)sum_prev / 100.0;
        for (int a = 0; a < A; ++a) { e->tracker_life[a] = e->things[h->S + h->P + a].life; sum_new += e->tracker_life[a]; }
        e->tracker_zd = e->zombie_deaths;
        double cur = (double)e->tracker_zd + (double)sum_new / 100.0;
        rew[0] = cur - prev;
    } else {
        for (int a = 0; a < A; ++a) {
            double prev = (double)e->tracker_zd + (double)e->tracker_life[a] / 100.0;
            e->tracker_life[a] = e->things[h->S + h->P + a].life;
            double cur = (double)e->zombie_deaths + (double)e->tracker_life[a] / 100.0;
            rew[a] = cur - prev;
        }
        e->tracker_zd = e->zombie_deaths;
    }
    /* Game.spawn_zombies_to_maintain_minimum (game.py:196-201) */
    if (h->cfg.minimum_zombies > 0) {
        int zombies = 0;
        for (int i = 0; i < e->n_order; ++i) if (e->things[e->order[i]].type == T_ZOMBIE) ++zombies;
        if (zombies < h->cfg.minimum_zombies) spawn_zombies(h, e, h->cfg.minimum_zombies - zombies);
    }
    e->step_draws = e->k;

    /* observation before the rules (gym_env.py:126; multiagent_env.py:163 — same world state) */
    int done = 0, trunc = 0;
    double end_reward = 0.0;
    if (game_ended(h, e)) { done = 1; end_reward = game_won(h, e) ? 10.0 : -10.0; }
    else if (!agents_alive(h, e)) { trunc = 1; end_reward = -10.0; }
    if (!h->cfg.obs_per_agent) {
        if (done || trunc) rew[0] += end_reward; /* gym_env.py:130-141 */
        reward[0] = rew[0];
    } else {
        for (int a = 0; a < A; ++a) { /* multiagent_env.py:156-162 */
            if (!alive_before[a]) { reward[a] = 0.0; continue; }
            reward[a] = e->things[h->S + h->P + a].life > 0 ? rew[a] + end_reward : rew[a];
        }
    }
    if (h->cfg.max_episode_steps > 0 && e->episode_steps >= h->cfg.max_episode_steps) trunc = 1;
    *terminated = (uint8_t)done; *truncated = (uint8_t)trunc;
    if (agent_mask) for (int a = 0; a < A; ++a) agent_mask[a] = alive_before[a];

    if ((done || trunc) && (h->cfg.auto_reset || force_auto_reset)) {
#ifdef _OPENMP
#pragma omp atomic
#endif
        h->stats[0] += 1;
        if (done && end_reward > 0) {
#ifdef _OPENMP
#pragma omp atomic
#endif
            h->stats[1] += 1;
        }
#ifdef _OPENMP
#pragma omp atomic
#endif
        h->stats[2] += e->episode_steps;
#ifdef _OPENMP
#pragma omp atomic
#endif
        h->stats[3] += e->zombie_deaths;
        initialize_world(h, e, env_local, e->episode + 1);
    }
    if (obs) encode_obs(h, e, obs);
}

/* ------------------------------------------------------------------ API */
ZSO_EXPORT ZsoHandle* zso_create(const ZsConfig* cfg, const ZsMap* map) {
    if (cfg->abi_version != ZS_ABI_VERSION) return NULL;
    ZsoHandle* h = (ZsoHandle*)calloc(1, sizeof(ZsoHandle));
    h->cfg = *cfg;
    h->W = map->width; h->H = map->height; h->cells = h->W * h->H;
    h->S = map->n_statics; h->P = cfg->n_bots; h->A = cfg->n_agents;
    h->Z = cfg->initial_zombies > cfg->minimum_zombies ? cfg->initial_zombies : cfg->minimum_zombies;
    h->M = h->P + h->A + h->Z;
    h->static_xy = (int16_t*)malloc(sizeof(int16_t) * 2 * (h->S + 1));
    h->static_label = (uint8_t*)malloc(h->S + 1);
    memcpy(h->static_xy, map->static_xy, sizeof(int16_t) * 2 * h->S);
    memcpy(h->static_label, map->static_label, h->S);
    h->n_ps = map->n_player_spawns; h->n_zs = map->n_zombie_spawns; h->n_obj = map->n_objectives;
    h->ps_xy = (int16_t*)malloc(sizeof(int16_t) * 2 * (h->n_ps + 1));
    h->zs_xy = (int16_t*)malloc(sizeof(int16_t) * 2 * (h->n_zs + 1));
    memcpy(h->ps_xy, map->player_spawn_xy, sizeof(int16_t) * 2 * h->n_ps);
    memcpy(h->zs_xy, map->zombie_spawn_xy, sizeof(int16_t) * 2 * h->n_zs);
    h->objective = (uint8_t*)calloc(h->cells, 1);
    for (int i = 0; i < h->n_obj; ++i) h->objective[map->objective_xy[2 * i + 1] * h->W + map->objective_xy[2 * i]] = 1;
    h->key[0] = (uint32_t)cfg->seed; h->key[1] = (uint32_t)(cfg->seed >> 32);
    if (cfg->obs_scope == ZS_OBS_WORLD) { h->obs_H = h->H; h->obs_W = h->W; }
    else { h->obs_H = h->obs_W = cfg->surroundings_width; }
    h->obs_C = cfg->obs_encoding == ZS_OBS_CHANNELS ? 3 : 1;
    h->obs_count = cfg->obs_per_agent ? h->A : 1;
    h->obs_elems = (int64_t)h->obs_count * h->obs_C * h->obs_H * h->obs_W;
    int N = cfg->num_envs;
    h->envs = (Env*)calloc(N, sizeof(Env));
    int scratch_n = h->cells > h->M ? h->cells : h->M;
    for (int i = 0; i < N; ++i) {
        Env* e = &h->envs[i];
        e->things = (Thing*)calloc(h->S + h->M, sizeof(Thing));
        e->order = (int32_t*)malloc(sizeof(int32_t) * (h->S + h->M));
        e->grid = (int32_t*)malloc(sizeof(int32_t) * h->cells);
        e->deco = (uint8_t*)malloc(h->cells);
        e->tracker_life = (int32_t*)calloc(h->A + 1, sizeof(int32_t));
        e->actions = (Action*)malloc(sizeof(Action) * (h->M + 1));
        e->scratch = (int32_t*)malloc(sizeof(int32_t) * (scratch_n + 1));
        for (int s = 0; s < h->S; ++s) { /* Map.from_file builds the Box/Wall objects once (game.py:76-79) */
            Thing* t = &e->things[s];
            t->x = h->static_xy[2 * s]; t->y = h->static_xy[2 * s + 1];
            t->type = h->static_label[s] == ZS_LABEL_BOX ? T_BOX : T_WALL;
            t->life = max_life(t->type);
        }
        initialize_world(h, e, i, 0); /* Game.__init__ -> __initialize_world__ (game.py:138) */
    }
    return h;
}

ZSO_EXPORT void zso_destroy(ZsoHandle* h) {
    if (!h) return;
    for (int i = 0; i < h->cfg.num_envs; ++i) {
        Env* e = &h->envs[i];
        free(e->things); free(e->order); free(e->grid); free(e->deco); free(e->tracker_life); free(e->actions); free(e->scratch);
    }
    free(h->envs); free(h->static_xy); free(h->static_label); free(h->ps_xy); free(h->zs_xy); free(h->objective);
    free(h);
}

ZSO_EXPORT int64_t zso_obs_elems(const ZsoHandle* h) { return h->obs_elems; }
ZSO_EXPORT int32_t zso_n_slots(const ZsoHandle* h) { return h->M; }

ZSO_EXPORT void zso_reset(ZsoHandle* h, const uint8_t* mask, int32_t* obs) {
    int N = h->cfg.num_envs;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < N; ++i) {
        if (mask && !mask[i]) continue;
        Env* e = &h->envs[i];
        initialize_world(h, e, i, e->episode + 1);
        if (obs) encode_obs(h, e, obs + (int64_t)i * h->obs_elems);
    }
}

ZSO_EXPORT void zso_encode_obs(ZsoHandle* h, int32_t* obs) {
    int N = h->cfg.num_envs;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < N; ++i) encode_obs(h, &h->envs[i], obs + (int64_t)i * h->obs_elems);
}

ZSO_EXPORT void zso_step(ZsoHandle* h, const int32_t* actions, int32_t fmt, int32_t* obs, double* reward,
                         uint8_t* terminated, uint8_t* truncated, uint8_t* agent_mask, int32_t* draws) {
    int N = h->cfg.num_envs, A = h->A;
    int astride = fmt == ZS_ACTIONS_DISCRETE ? A : 3 * A;
    int rstride = h->cfg.obs_per_agent ? A : 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int i = 0; i < N; ++i) {
        Env* e = &h->envs[i];
        env_step(h, e, i, actions + (int64_t)i * astride, fmt, obs ? obs + (int64_t)i * h->obs_elems : NULL,
                 reward + (int64_t)i * rstride, terminated + i, truncated + i, agent_mask ? agent_mask + (int64_t)i * A : NULL, 0);
        if (draws) draws[i] = e->step_draws;
    }
}

/* uniform discrete ids from the synthetic action stream (libzombsole_b200/philox.py:synthetic_actions) */
static int32_t synthetic_action(const ZsoHandle* h, int64_t env_global, int64_t step_index, int agent, int n_actions) {
    uint32_t key[2] = { h->key[0], h->key[1] ^ 0xAC710115u };
    uint32_t ctr[4] = { (uint32_t)env_global, (uint32_t)step_index, 0u, (uint32_t)(agent >> 2) };
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return (int32_t)(((uint64_t)out[agent & 3] * (uint32_t)n_actions) >> 32);
}
ZSO_EXPORT void zso_synthetic_actions(const ZsoHandle* h, int64_t step_index, int32_t* actions) {
    int n = h->cfg.obs_per_agent ? 7 : 6;
    for (int i = 0; i < h->cfg.num_envs; ++i)
        for (int a = 0; a < h->A; ++a)
            actions[(int64_t)i * h->A + a] = synthetic_action(h, h->cfg.env_index_base + i, step_index, a, n);
}

/* K synthetic steps with same-step auto-reset; obs written to a single [N, obs_elems] buffer (or NULL).
 * Used by bench.py's cpu_baseline / --impl reference legs. */
ZSO_EXPORT void zso_rollout_synthetic(ZsoHandle* h, int32_t n_steps, int64_t first_step_index, int32_t* obs,
                                      double* reward, uint8_t* terminated, uint8_t* truncated) {
    int N = h->cfg.num_envs, A = h->A;
    int n = h->cfg.obs_per_agent ? 7 : 6;
    int rstride = h->cfg.obs_per_agent ? A : 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int i = 0; i < N; ++i) {
        Env* e = &h->envs[i];
        int32_t acts[ZS_MAX_AGENTS];
        for (int s = 0; s < n_steps; ++s) {
            for (int a = 0; a < A; ++a) acts[a] = synthetic_action(h, h->cfg.env_index_base + i, first_step_index + s, a, n);
            env_step(h, e, i, acts, ZS_ACTIONS_DISCRETE, obs ? obs + (int64_t)i * h->obs_elems : NULL,
                     reward + ((int64_t)s * N + i) * rstride, terminated + (int64_t)s * N + i, truncated + (int64_t)s * N + i, NULL, 1);
        }
    }
}

ZSO_EXPORT void zso_stats(ZsoHandle* h, int64_t* out, int32_t reset) {
    memcpy(out, h->stats, sizeof(h->stats));
    if (reset) memset(h->stats, 0, sizeof(h->stats));
}

/* Slot-indexed state export of one env (same shapes as oracle/ref_harness.py dumps). */
ZSO_EXPORT void zso_export(const ZsoHandle* h, int32_t env, int16_t* x, int16_t* y, int16_t* life, uint8_t* in_world,
                           uint8_t* weapon, int16_t* order, int16_t* static_life, uint8_t* static_present,
                           uint8_t* dead_body_cells, int32_t* counters) {
    const Env* e = &h->envs[env];
    for (int s = 0; s < h->M; ++s) {
        const Thing* t = &e->things[h->S + s];
        x[s] = t->x; y[s] = t->y; life[s] = (int16_t)t->life; in_world[s] = t->in_world; weapon[s] = t->weapon;
        order[s] = -1;
    }
    int n = 0;
    for (int i = 0; i < e->n_order; ++i) if (e->order[i] >= h->S) order[n++] = (int16_t)(e->order[i] - h->S);
    for (int s = 0; s < h->S; ++s) { static_life[s] = (int16_t)e->things[s].life; static_present[s] = e->things[s].in_world; }
    for (int c = 0; c < h->cells; ++c) dead_body_cells[c] = e->deco[c] == ZS_LABEL_DEAD_BODY;
    counters[0] = e->t; counters[1] = e->deaths; counters[2] = e->zombie_deaths;
    counters[3] = e->episode; counters[4] = e->episode_steps; counters[5] = e->last_draws; counters[6] = e->step_draws;
}

/* Overwrite lives (tests mirror the reference tests that poke thing.life, tests/test_game.py:54,82,106). */
ZSO_EXPORT void zso_set_life(ZsoHandle* h, int32_t env, int32_t slot, int32_t life) { h->envs[env].things[h->S + slot].life = life; }
ZSO_EXPORT void zso_set_static_life(ZsoHandle* h, int32_t env, int32_t index, int32_t life) { h->envs[env].things[index].life = life; }

ZSO_EXPORT int32_t zso_set_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}
