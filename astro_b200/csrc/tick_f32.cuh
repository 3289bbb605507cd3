// tick_f32.cuh — the production tick kernel (included by astro_b200.cu inside its anonymous
// namespace, after TickParams / recreate_from_pool / warp_stats).
//
// One warp owns one 32-game tile.  Two phases:
//
//  A. BULLETS, warp-cooperative.  Each lane stages its game's ships and planets in shared
//     memory, the per-game bullet counts are prefix-summed across the warp, and the tile's
//     bullets are processed as ONE flat list, 32 per iteration, whatever game they belong to —
//     a warp with bullet counts (0, 17, 3, ...) runs ceil(sum/32) iterations, not max = 17.
//     Per bullet: squared distances to both ships and four planet slots (dead slots hold a
//     far-away sentinel) reduced with min, the advance, and ONE band test deciding whether the
//     fp32 comparisons provably equal the reference's float64 results; only a bullet inside a
//     band (|d2 - R2| <= 1e-6 R2, or |x'| within 4e-6 of the arena bound — ~1e-5 of bullets),
//     and every ship hit, takes the float64 path.  Survivors are compacted in place with a
//     warp ballot: rank inside the game's segment = popc(ballot & segment mask), so the
//     reference order (stable) is kept.  Bullets are game-major in HBM (a game's pool is one
//     contiguous row), so a window of 32 consecutive list items reads a few contiguous runs.
//  B. SHIPS AND PLANETS, thread-per-game from the staged copies: direction, gravity, collisions,
//     terminal logic, spawn, integration, reset — all coalesced 128-bit accesses.
#pragma once

// Dead planet slots sit at kFar: the squared distance overflows to +inf, so a dead slot is never
// a minimum, never inside a band, and its gravity G*M/inf is exactly 0 — no predication on np.
constexpr float kFar = 1.0e20f;
constexpr int kTickWarps = kTickThreads / 32;
#ifndef ASTRO_PREFETCH_PLANETS
#define ASTRO_PREFETCH_PLANETS 4
#endif

constexpr int kStageWindows = 8;   // bullets staged per round: 8 windows x 32 = 256 (4 KB per warp)

struct TileScratch {               // per warp
    float4 bul[kStageWindows * 32];  // the tile's bullets, staged by cp.async (flat list order)
    float4 sxy[32];                // OLD ship0.xy, ship1.xy          } what the bullet loop reads,
    float4 pxy[2][32];             // OLD planet0.xy planet1.xy / 2,3 } addressed by game
    float4 svel[32];               // OLD ship velocities   } for the newborn bullets
    float4 dir[32];                // sin/cos of both bearings }
    uint32_t cinfo[32];            // k-th non-empty game: game | first list index << 5
    uint32_t outn[32];             // survivors written so far
    uint32_t hits[32];             // bits 0-1: ship hits found by the bullet loop; bits 8+: np
    uint16_t ref[kStageWindows * 32];  // staged item -> game | slot << 5 ; 0xFFFF = none
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Exact (reference float64) evaluation of one bullet against the OLD ship / planet positions:
// despawn flags, per-ship hit bits, advance + cull.  Taken by ~1e-5 of bullets.
struct ExactResult {
    float x, y;
    unsigned flags;  // bit 0 keep, bits 1-2 ship hits
};
__device__ __forceinline__ ExactResult bullet_exact(float bx, float by, float bdx, float bdy, float4 sxy, float4 p01,
                                                    float4 p23, int n_ships, int np, const Consts& c) {
    unsigned ship_hits = 0;
    if (collide_exact((double)sxy.x, (double)sxy.y, (double)bx, (double)by, c.r2_sb)) ship_hits |= 1u;
    if (n_ships > 1 && collide_exact((double)sxy.z, (double)sxy.w, (double)bx, (double)by, c.r2_sb)) ship_hits |= 2u;
    bool gone = ship_hits != 0;
    gone |= collide_exact((double)p01.x, (double)p01.y, (double)bx, (double)by, c.r2_pb);  // np >= 1
    if (np > 1) gone |= collide_exact((double)p01.z, (double)p01.w, (double)bx, (double)by, c.r2_pb);
    if (np > 2) gone |= collide_exact((double)p23.x, (double)p23.y, (double)bx, (double)by, c.r2_pb);
    if (np > 3) gone |= collide_exact((double)p23.z, (double)p23.w, (double)bx, (double)by, c.r2_pb);
    double v0 = __dadd_rn((double)bdx, c.zero_dt), v1 = __dadd_rn((double)bdy, c.zero_dt);
    double e0 = __dadd_rn((double)bx, __dmul_rn(c.dt, v0)), e1 = __dadd_rn((double)by, __dmul_rn(c.dt, v1));
    ExactResult r;
    r.x = (float)e0;
    r.y = (float)e1;
    r.flags = ((in_arena(e0, e1) && !gone) ? 1u : 0u) | (ship_hits << 1);
    return r;
}

__device__ __forceinline__ float dist2(float ax, float ay, float bx, float by) {
    float d0 = __fsub_rn(ax, bx), d1 = __fsub_rn(ay, by);
    return __fmaf_rn(d1, d1, __fmul_rn(d0, d0));
}

// One bullet against its game's staged frame.  Returns keep; b.x/b.y advanced; ship_hits
// receives bits 0/1 when the bullet touches ship 0/1 (always decided in float64).
// Arena test: keep iff (|x'| <= 1) or (|y'| <= 1)  <=>  min(|x'|, |y'|) <= 1, so only the
// smaller magnitude can sit in the uncertainty band.
template <int S>
__device__ __forceinline__ bool bullet_step(Body4<float>& b, float4 sxy, float4 p01, float4 p23, const uint32_t* np_of,
                                            unsigned gi, const Consts& c, unsigned& ship_hits) {  // np_of[g] >> 8 = np
    float ds = dist2(sxy.x, sxy.y, b.x, b.y);
    if (S == 2) ds = fminf(ds, dist2(sxy.z, sxy.w, b.x, b.y));
    float dp = fminf(fminf(dist2(p01.x, p01.y, b.x, b.y), dist2(p01.z, p01.w, b.x, b.y)),
                     fminf(dist2(p23.x, p23.y, b.x, b.y), dist2(p23.z, p23.w, b.x, b.y)));
    float x0 = __fmaf_rn(c.dt_f, b.dx, b.x), x1 = __fmaf_rn(c.dt_f, b.dy, b.y);
    float mn = fminf(fabsf(x0), fabsf(x1));
    bool sure = (ds >= c.r2f_sb * 1.000001f) & (fabsf(dp - c.r2f_pb) > c.r2f_pb * 1e-6f) & (fabsf(mn - 1.0f) > 4e-6f);
    bool keep = (mn <= 1.0f) & (dp >= c.r2f_pb);
    if (__builtin_expect(!sure, 0)) {
        ExactResult r = bullet_exact(b.x, b.y, b.dx, b.dy, sxy, p01, p23, S, (int)(np_of[gi] >> 8), c);
        x0 = r.x;
        x1 = r.y;
        keep = r.flags & 1u;
        ship_hits = r.flags >> 1;
    }
    b.x = x0;
    b.y = x1;
    return keep;
}

// Flat bullet list of a tile: item i of the list -> (game, slot).  `starts` = bit r set when a
// non-empty game's first bullet is item base + r; c0 = non-empty games that start before base.
struct ItemRef {
    unsigned game, excl;
    bool valid;
};
__device__ __forceinline__ ItemRef map_item(const TileScratch& t, unsigned base, unsigned total, unsigned my_excl,
                                            bool my_nonempty, unsigned lane, unsigned& c0) {
    const unsigned full = 0xffffffffu;
    unsigned rel = my_excl - base;  // wraps to a huge value when the game starts before base
    unsigned starts = __reduce_or_sync(full, (my_nonempty && rel < 32u) ? (1u << rel) : 0u);
    unsigned le = full >> (31u - lane);
    unsigned idx = c0 + __popc(starts & le);  // >= 1 for a valid item
    c0 += __popc(starts);
    unsigned ci = t.cinfo[(idx - 1u) & 31u];  // (stale only for invalid items)
    ItemRef r;
    r.valid = base + lane < total;
    r.game = ci & 31u;
    r.excl = ci >> 5;
    return r;
}

#ifndef ASTRO_TICK_MIN_BLOCKS
#define ASTRO_TICK_MIN_BLOCKS 26  /* shared memory admits 26 one-warp CTAs per SM: 72 registers */
#endif
template <int S, bool STATS>
__global__ void __launch_bounds__(kTickThreads, ASTRO_TICK_MIN_BLOCKS) tick_f32_kernel(const __grid_constant__ TickParams p) {
    using B4 = Body4<float>;
    const unsigned full = 0xffffffffu;
    __shared__ TileScratch s_tiles[kTickWarps];
    const int g = blockIdx.x * kTickThreads + threadIdx.x;
    if (g >= p.n_games) return;  // whole warps: n_games % 32 == 0
    const unsigned lane = threadIdx.x & 31u;
    const Consts& c = p.c;
    TileScratch& t = s_tiles[threadIdx.x >> 5];
    const size_t tile = (size_t)(g >> 5);
    B4* ships = reinterpret_cast<B4*>(p.ships) + tile * (S * 32) + lane;
    float* ship_b = reinterpret_cast<float*>(p.ship_b) + tile * (S * 32) + lane;
    B4* planets = reinterpret_cast<B4*>(p.planets) + tile * (ASTRO_MAX_PLANETS * 32) + lane;
    B4* tile_bullets = reinterpret_cast<B4*>(p.bullets) + tile * 32 * (size_t)p.K;
    const unsigned K = (unsigned)p.K;  // 32-bit slot arithmetic: a tile's pool is <= 32 * 1023 slots

    // ================= 1. this lane's game: loads (independent except planets <- meta) ========
    const uint32_t meta = p.meta[g];
    // The planet slots to load depend on meta (np).  Warm L2 with the tile's planet rows meanwhile:
    // the dependent loads below then take an L2 round trip instead of an HBM one (measured:
    // 93.4 -> 91.4 us per 1M-game tick; sectors without a live lane are the price).
#pragma unroll
    for (int j = 0; j < ASTRO_PREFETCH_PLANETS; j++)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&planets[j * 32]));
    float4 shv[S];
    float sb[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        shv[s] = *reinterpret_cast<const float4*>(&ships[s * 32]);
        sb[s] = ship_b[s * 32];
    }
    int ctl[S];
    if (p.actions) {
        if (S == 2) {
            uint16_t a = reinterpret_cast<const uint16_t*>(p.actions)[g];
            ctl[0] = a & 0xff;
            ctl[S - 1] = a >> 8;
        } else {
            ctl[0] = p.actions[g];
        }
    } else {
        uint32_t h0 = game_key(p.seed, p.first_game + (uint32_t)g);
#pragma unroll
        for (int s = 0; s < S; s++) ctl[s] = action_from_key(h0, p.step, (uint32_t)s);
    }
    const bool active = !ASTRO_META_FINISHED(meta);
    const int nb = active ? (int)ASTRO_META_NB(meta) : 0;
    const int np = active ? (int)ASTRO_META_NP(meta) : 0;
    const uint32_t tick = ASTRO_META_TICK(meta);
    float4 plv[ASTRO_MAX_PLANETS];
#pragma unroll
    for (int j = 0; j < ASTRO_MAX_PLANETS; j++) {
        plv[j] = make_float4(kFar, kFar, 0.f, 0.f);
        if (j < np) plv[j] = *reinterpret_cast<const float4*>(&planets[j * 32]);
    }

    // ================= 2. flat bullet list of the tile; stage it with cp.async =================
    unsigned incl = (unsigned)nb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned v = __shfl_up_sync(full, incl, d);
        if ((int)lane >= d) incl += v;
    }
    const unsigned my_excl = incl - (unsigned)nb;
    const unsigned total = __shfl_sync(full, incl, 31);
    const bool nonempty = nb > 0;
    t.outn[lane] = 0u;
    t.hits[lane] = (unsigned)np << 8;
    const unsigned ne = __ballot_sync(full, nonempty);
    const unsigned lt_mask = (1u << lane) - 1u;
    if (nonempty) t.cinfo[__popc(ne & lt_mask)] = lane | (my_excl << 5);
    __syncwarp();
    unsigned c0 = 0;
    // One window = 32 consecutive list items; a round = up to kStageWindows windows, all requested
    // at once (16-byte cp.async each), so the whole tile's bullet traffic is in flight together.
    const unsigned start_key = nonempty ? my_excl : 0x80000000u;  // empty games never "start"
    const unsigned le_mask = full >> (31u - lane);
    const unsigned bul_s = (unsigned)__cvta_generic_to_shared(&t.bul[lane]);
    auto stage_round = [&](unsigned round_base) {
        const unsigned left = total - round_base;
        const unsigned n_win = left >= (unsigned)kStageWindows * 32u ? (unsigned)kStageWindows : (left + 31u) >> 5;
#pragma unroll 1
        for (unsigned w = 0; w < n_win; w++) {
            const unsigned base = round_base + w * 32u;
            const unsigned rel = start_key - base;  // >= 32 unless this lane's game starts in the window
            const unsigned starts = __reduce_or_sync(full, rel < 32u ? (1u << rel) : 0u);
            const unsigned idx = c0 + __popc(starts & le_mask);  // >= 1 for a valid item
            c0 += __popc(starts);
            const unsigned ci = t.cinfo[(idx - 1u) & 31u];  // (stale only for invalid items)
            const unsigned item = base + lane;
            const bool valid = item < total;
            const unsigned game = ci & 31u, slot = item - (ci >> 5);
            if (valid)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bul_s + w * 512u),
                             "l"(tile_bullets + (game * K + slot))
                             : "memory");
            t.ref[w * 32u + lane] = valid ? (uint16_t)(game | (slot << 5)) : (uint16_t)0xFFFFu;
        }
        cp_async_commit();
    };
    stage_round(0u);

    // ================= 3. ships and planets while the bullets fly ===============================
    // OLD positions are staged for the bullet loop (addressed by game); the new ship / planet state
    // goes straight to HBM.  (A game that ends is re-created below; without auto-reset its
    // ships and planets are left in this post-step state: the state of a finished game is
    // unspecified, the reference has none.)
    t.sxy[lane] = make_float4(shv[0].x, shv[0].y, shv[S - 1].x, shv[S - 1].y);
    t.svel[lane] = make_float4(shv[0].z, shv[0].w, shv[S - 1].z, shv[S - 1].w);
    t.pxy[0][lane] = make_float4(plv[0].x, plv[0].y, plv[1].x, plv[1].y);
    t.pxy[1][lane] = make_float4(plv[2].x, plv[2].y, plv[3].x, plv[3].y);
    unsigned hits = 0;
    if (active) {
        B4 sh[S];
#pragma unroll
        for (int s = 0; s < S; s++) { sh[s].x = shv[s].x; sh[s].y = shv[s].y; sh[s].dx = shv[s].z; sh[s].dy = shv[s].w; }
        B4 pl[ASTRO_MAX_PLANETS];
#pragma unroll
        for (int j = 0; j < ASTRO_MAX_PLANETS; j++) { pl[j].x = plv[j].x; pl[j].y = plv[j].y; pl[j].dx = plv[j].z; pl[j].dy = plv[j].w; }
        float dirs[4] = {0.f, 0.f, 0.f, 0.f};
        // direction, gravity, ship-planet and ship-ship collisions on the old state
#pragma unroll
        for (int s = 0; s < S; s++) {
            float d0, d1;
            np_sincos_f32(sb[s], d0, d1);
            dirs[2 * s] = d0;
            dirs[2 * s + 1] = d1;
            float g0 = 0.f, g1 = 0.f, dmin = 3.0e38f;
#pragma unroll
            for (int j = 0; j < ASTRO_MAX_PLANETS; j++) {
                float q0 = __fsub_rn(pl[j].x, sh[s].x), q1 = __fsub_rn(pl[j].y, sh[s].y);
                float d2 = __fmaf_rn(q1, q1, __fmul_rn(q0, q0));
                float fj = __fdividef(c.gm_f, fmaxf(1e-12f, d2));  // dead slot: G*M / inf = 0
                g0 = __fmaf_rn(fj, q0, g0);
                g1 = __fmaf_rn(fj, q1, g1);
                dmin = fminf(dmin, d2);
            }
            bool h = dmin < c.r2f_sp;
            if (__builtin_expect(fabsf(dmin - c.r2f_sp) <= c.r2f_sp * 1e-6f, 0)) {
                h = false;
#pragma unroll
                for (int j = 0; j < ASTRO_MAX_PLANETS; j++)
                    if (j < np)
                        h |= collide_exact((double)sh[s].x, (double)sh[s].y, (double)pl[j].x, (double)pl[j].y, c.r2_sp);
            }
            hits |= h ? (1u << s) : 0u;
            float th = (ctl[s] & 1) ? c.thrust_f : 0.f;
            B4 o = sh[s];
            advance_body(o, __fmaf_rn(th, d0, g0), __fmaf_rn(th, d1, g1), c);  // core.py:283-288
            ships[s * 32] = o;
            ship_b[s * 32] = __fmaf_rn(c.db_unit_f, (float)((ctl[s] >> 1) - 1), sb[s]);
        }
        t.dir[lane] = make_float4(dirs[0], dirs[1], dirs[2], dirs[3]);
        if (S == 2) {
            if (collide(sh[0].x, sh[0].y, sh[S - 1].x, sh[S - 1].y, c.r2_ss, c.r2f_ss)) hits |= 3u;
        }
        // planets (core.py:289-294): pair forces are antisymmetric, the clamped self term is zero
        float q0[ASTRO_MAX_PLANETS], q1[ASTRO_MAX_PLANETS];
#pragma unroll
        for (int i = 0; i < ASTRO_MAX_PLANETS; i++) { q0[i] = 0.f; q1[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < ASTRO_MAX_PLANETS; i++) {
#pragma unroll
            for (int j = i + 1; j < ASTRO_MAX_PLANETS; j++) {
                float e0 = __fsub_rn(pl[j].x, pl[i].x), e1 = __fsub_rn(pl[j].y, pl[i].y);
                float d2 = __fmaf_rn(e1, e1, __fmul_rn(e0, e0));
                float fj = __fdividef(c.gm_f, fmaxf(1e-12f, d2));  // dead: 0, or e = 0
                q0[i] = __fmaf_rn(fj, e0, q0[i]);
                q1[i] = __fmaf_rn(fj, e1, q1[i]);
                q0[j] = __fmaf_rn(-fj, e0, q0[j]);
                q1[j] = __fmaf_rn(-fj, e1, q1[j]);
            }
        }
#pragma unroll
        for (int i = 0; i < ASTRO_MAX_PLANETS; i++) {
            if (i < np) {
                advance_body(pl[i], q0[i], q1[i], c);
                planets[i * 32] = pl[i];
            }
        }
    }

    // ================= 4. the bullet loop, from shared memory =======================================
    for (unsigned round_base = 0; round_base < total; round_base += (unsigned)kStageWindows * 32u) {
        if (round_base) stage_round(round_base);  // (tiles with more than 256 bullets: rare)
        cp_async_wait_all();
        __syncwarp();
        const unsigned left = total - round_base;
        const unsigned n_win = left >= (unsigned)kStageWindows * 32u ? (unsigned)kStageWindows : (left + 31u) >> 5;
#pragma unroll 1
        for (unsigned w = 0; w < n_win; w++) {
            const unsigned ref = t.ref[w * 32u + lane];
            const bool valid = ref != 0xFFFFu;
            const unsigned gi = ref & 31u, slot = (ref >> 5) & 1023u;
            const float4 bv = t.bul[w * 32u + lane];
            B4 b;
            b.x = bv.x; b.y = bv.y; b.dx = bv.z; b.dy = bv.w;
            bool keep = false;
            unsigned sh_hits = 0;
            if (valid) keep = bullet_step<S>(b, t.sxy[gi], t.pxy[0][gi], t.pxy[1][gi], t.hits, gi, c, sh_hits);
            if (sh_hits) atomicOr(&t.hits[gi], sh_hits);
            // stable in-place compaction inside each game's segment of the window
            const unsigned kb = __ballot_sync(full, keep);
            const unsigned seg_lo = slot < lane ? lane - slot : 0u;  // first lane of this game's segment
            const unsigned rank = __popc(kb & (lt_mask & (full << seg_lo)));
            const unsigned ob = t.outn[gi];
            const unsigned g_next = __shfl_down_sync(full, valid ? gi : 32u, 1);
            const bool last = valid && (lane == 31u || g_next != gi);
            __syncwarp();
            if (keep) tile_bullets[gi * K + ob + rank] = b;
            if (last) t.outn[gi] = ob + rank + (keep ? 1u : 0u);
            __syncwarp();
        }
    }

    // ================= 5. terminal logic, spawn, bookkeeping ==========================================
    uint32_t ev = 0;
    int m_out = 0, spawned = 0;
    float rw[S];
#pragma unroll
    for (int s = 0; s < S; s++) rw[s] = 0.0f;
    if (!active) {
        ev = ASTRO_EV_SKIPPED;
    } else {
        int m = (int)t.outn[lane];
        hits |= t.hits[lane] & 3u;
        const bool timeout = tick >= (uint32_t)p.timeout_tick;
        if (hits) {  // core.py:253-255
            ev = hits;  // ASTRO_EV_HIT0 | ASTRO_EV_HIT1 are bits 0 and 1
#pragma unroll
            for (int s = 0; s < S; s++) rw[s] = ((hits >> s) & 1u) ? -1.0f : 1.0f;
        } else if (timeout) {  // core.py:257-260
            ev = ASTRO_EV_TIMEOUT;
#pragma unroll
            for (int s = 0; s < S; s++) rw[s] = c.reward_timeout;
        } else {
            const bool fire = tick < (uint32_t)p.n_sched_ticks && ((p.fire_bits[tick >> 5] >> (tick & 31)) & 1u);
            if (fire) {  // core.py:267-280, from the OLD ship state
                ev |= ASTRO_EV_FIRED;
                B4* row = tile_bullets + lane * K;
                const float4 dv = t.dir[lane], oxy = t.sxy[lane], ov = t.svel[lane];
#pragma unroll
                for (int s = 0; s < S; s++) {
                    const float d0 = s == 0 ? dv.x : dv.z, d1 = s == 0 ? dv.y : dv.w;
                    const float sx = s == 0 ? oxy.x : oxy.z, sy = s == 0 ? oxy.y : oxy.w;
                    const float vx = s == 0 ? ov.x : ov.z, vy = s == 0 ? ov.y : ov.w;
                    // fp32 products as in the reference; the sums and the advance in fp32 too,
                    // unless the newborn lands within the band of the arena bound
                    float o0 = __fmul_rn(c.off_f, d0), o1 = __fmul_rn(c.off_f, d1);
                    float w0 = __fmul_rn(c.spd_f, d0), w1 = __fmul_rn(c.spd_f, d1);
                    B4 o;
                    o.dx = __fadd_rn(vx, w0);
                    o.dy = __fadd_rn(vy, w1);
                    o.x = __fmaf_rn(c.dt_f, o.dx, __fadd_rn(sx, o0));
                    o.y = __fmaf_rn(c.dt_f, o.dy, __fadd_rn(sy, o1));
                    float mn = fminf(fabsf(o.x), fabsf(o.y));
                    bool keep = mn <= 1.0f;
                    if (__builtin_expect(fabsf(mn - 1.0f) <= 8e-6f, 0)) {
                        Body4<double> nbl;
                        nbl.x = __dadd_rn((double)sx, (double)o0);
                        nbl.y = __dadd_rn((double)sy, (double)o1);
                        nbl.dx = __dadd_rn((double)vx, (double)w0);
                        nbl.dy = __dadd_rn((double)vy, (double)w1);
                        keep = advance_bullet(nbl, c);
                        o.x = (float)nbl.x; o.y = (float)nbl.y; o.dx = (float)nbl.dx; o.dy = (float)nbl.dy;
                    }
                    if (keep) {
                        if (m < (int)K) {
                            row[m] = o;
                            m++;
                        } else {
                            ev |= ASTRO_EV_OVERFLOW;
                        }
                    }
                }
                spawned = S;
            }
            p.meta[g] = ASTRO_META_PACK(m, np, 0, tick + 1);
            m_out = m;
        }
        if (ev & ASTRO_EV_DONE_MASK) {
            if ((p.flags & ASTRO_TICK_AUTO_RESET) && p.pool_size > 0)
                recreate_from_pool<float, S>(p, g, p.step + 1u, ships, ship_b, planets);
            else
                p.meta[g] = ASTRO_META_PACK(0, np, 1, tick);
        }
    }
    if (p.reward) {
        if (S == 2) reinterpret_cast<float2*>(p.reward)[g] = make_float2(rw[0], rw[S - 1]);
        else p.reward[g] = rw[0];
    }
    if (p.events) p.events[g] = (uint8_t)ev;
    if (p.done) p.done[g] = (uint8_t)((ev & (ASTRO_EV_DONE_MASK | ASTRO_EV_SKIPPED)) ? 1 : 0);

    if (STATS) {
        // warp totals -> this warp's private slot row in HBM (no block barrier, no contention);
        // astro_stats() folds the rows
        unsigned mine = warp_totals((int)lane, S, ev, active, spawned, np, nb, m_out);
        unsigned* slot = p.stat_slots + ((size_t)(g >> 5) * 16u + lane);
        if (lane < ASTRO_N_STATS && mine) atomicAdd(slot, mine);  // RED: fire and forget
    }
}
