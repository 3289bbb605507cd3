// tick_f32_pipe.cuh — persistent, software-pipelined form of the production tick kernel
// (same arithmetic and helpers as tick_f32.cuh; included after it).
//
// Each warp walks tiles  w, w + W, w + 2W, ...  (W = warps in the grid).  A tile's two dependent
// HBM round trips (meta -> planets/bullets) are taken off the critical path:
//
//   iteration k (tile k):
//     T0  meta[k+2] is requested (registers); ships / planets / bearings / controls of tile k
//         are already in shared memory (requested during iteration k-1) -> registers
//     T1  warp scan of the bullet counts, the tile's bullets are requested with cp.async; then the
//         rows of tile k+1 (np[k+1] known, raw buffer free) are requested as a second group
//     T2  ship / planet physics of tile k, new state stored to HBM; OLD positions parked in smem
//     T3  wait for the bullets; bullet loop of tile k from shared memory
//     T4  terminal logic, spawn, bookkeeping of tile k
//
// so every request has a whole phase of arithmetic between issue and use.  Warps are independent:
// no block barrier anywhere.
#pragma once

struct RawRows {                   // NEXT tile's rows, filled by cp.async one iteration ahead
    float4 ship[2][32];
    float4 planet[4][32];
    float sb[2][32];
    uint32_t ctl[16];              // 32 x 2 control bytes (or 32 x 1)
};
struct NoRows {};

template <bool ROWS>
struct PipeScratch {               // per warp: 7,552 bytes, + 3,392 with ROWS
    float4 bul[kStageWindows * 32];  // the tile's bullets, staged by cp.async (flat list order)
    float4 sxy[32];                // OLD ship0.xy, ship1.xy          } what the bullet loop reads,
    float4 pxy[2][32];             // OLD planet0.xy planet1.xy / 2,3 } addressed by game
    float4 svel[32];               // OLD ship velocities   } for the newborn bullets
    float4 dir[32];                // sin/cos of both bearings }
    uint32_t cinfo[32];            // k-th non-empty game: game | first list index << 5
    uint32_t outn[32];             // survivors written so far
    uint32_t hits[32];             // bits 0-1: ship hits found by the bullet loop; bits 8+: np
    uint16_t ref[kStageWindows * 32];  // staged item -> game | slot << 5 ; 0xFFFF = none
    typename std::conditional<ROWS, RawRows, NoRows>::type raw;
};

__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}

#ifndef ASTRO_PIPE_THREADS
#define ASTRO_PIPE_THREADS 128
#endif
constexpr int kPipeThreads = ASTRO_PIPE_THREADS;
constexpr int kPipeWarps = kPipeThreads / 32;

// ROWS = true : the full pipeline above (5 CTAs of 4 warps per SM: 11 KB of shared memory per warp).
// ROWS = false: only the meta word travels ahead (one register); the rows are loaded at T0 with
//               np already known — one exposed round trip per tile instead of two — and 7 CTAs fit.
template <int S, bool STATS, bool ROWS>
__global__ void __launch_bounds__(kPipeThreads, ROWS ? 5 : 7) tick_f32_pipe_kernel(const __grid_constant__ TickParams p) {
    using B4 = Body4<float>;
    const unsigned full = 0xffffffffu;
    extern __shared__ float4 s_pipe_raw[];
    PipeScratch<ROWS>& t = reinterpret_cast<PipeScratch<ROWS>*>(s_pipe_raw)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const Consts& c = p.c;
    const unsigned K = (unsigned)p.K;  // 32-bit slot arithmetic: a tile's pool is <= 32 * 1023 slots
    const int n_tiles = p.n_games >> 5;
    const int W = (int)gridDim.x * kPipeWarps;
    int tile = (int)blockIdx.x * kPipeWarps + (int)(threadIdx.x >> 5);
    if (tile >= n_tiles) return;

    // request the rows of one tile into the raw buffer (its meta word already known)
    auto request_rows = [&](int tl, uint32_t meta_w) {
        if constexpr (ROWS) {
        const bool act = !ASTRO_META_FINISHED(meta_w);
        const int npn = act ? (int)ASTRO_META_NP(meta_w) : 0;
        const B4* sh_g = reinterpret_cast<const B4*>(p.ships) + (size_t)tl * (S * 32) + lane;
        const float* sb_g = reinterpret_cast<const float*>(p.ship_b) + (size_t)tl * (S * 32) + lane;
        const B4* pl_g = reinterpret_cast<const B4*>(p.planets) + (size_t)tl * (ASTRO_MAX_PLANETS * 32) + lane;
#pragma unroll
        for (int s = 0; s < S; s++) {
            cp_async16(&t.raw.ship[s][lane], &sh_g[s * 32]);
            cp_async4(&t.raw.sb[s][lane], &sb_g[s * 32]);
        }
#pragma unroll
        for (int j = 0; j < ASTRO_MAX_PLANETS; j++)
            if (j < npn) cp_async16(&t.raw.planet[j][lane], &pl_g[j * 32]);
        if (p.actions && lane < (unsigned)(8 * S))
            cp_async4(&t.raw.ctl[lane], p.actions + (size_t)tl * (32 * S) + lane * 4u);
        cp_async_commit();
        }
    };

    uint32_t meta = p.meta[tile * 32 + (int)lane];
    request_rows(tile, meta);
    uint32_t meta_next = 0;  // meta words travel two tiles ahead, rows one tile ahead
    if (tile + W < n_tiles) meta_next = p.meta[(tile + W) * 32 + (int)lane];

    for (; tile < n_tiles; tile += W) {
        const int g = tile * 32 + (int)lane;
        B4* ships = reinterpret_cast<B4*>(p.ships) + (size_t)tile * (S * 32) + lane;
        float* ship_b = reinterpret_cast<float*>(p.ship_b) + (size_t)tile * (S * 32) + lane;
        B4* planets = reinterpret_cast<B4*>(p.planets) + (size_t)tile * (ASTRO_MAX_PLANETS * 32) + lane;
        B4* tile_bullets = reinterpret_cast<B4*>(p.bullets) + (size_t)tile * 32 * (size_t)p.K;

        // ================= T0: meta[k+2] in flight; this tile's rows -> registers ===================
        const int tile_next = tile + W;
        uint32_t meta_next2 = 0;
        if (tile_next + W < n_tiles) meta_next2 = p.meta[(tile_next + W) * 32 + (int)lane];
        const bool active = !ASTRO_META_FINISHED(meta);
        const int nb = active ? (int)ASTRO_META_NB(meta) : 0;
        const int np = active ? (int)ASTRO_META_NP(meta) : 0;
        const uint32_t tick = ASTRO_META_TICK(meta);
        // fire flag of this tick (core.py:267), requested early: it is consumed at T4
        const bool fire = tick < (uint32_t)p.n_sched_ticks && ((p.fire_bits[tick >> 5] >> (tick & 31)) & 1u);
        float4 shv[S];
        float sb[S];
        float4 plv[ASTRO_MAX_PLANETS];
        int ctl[S];
        if constexpr (ROWS) {
            cp_async_wait_all();
            __syncwarp();
#pragma unroll
            for (int s = 0; s < S; s++) {
                shv[s] = t.raw.ship[s][lane];
                sb[s] = t.raw.sb[s][lane];
            }
#pragma unroll
            for (int j = 0; j < ASTRO_MAX_PLANETS; j++) {
                plv[j] = make_float4(kFar, kFar, 0.f, 0.f);
                if (j < np) plv[j] = t.raw.planet[j][lane];
            }
            if (p.actions) {
                if (S == 2) {
                    unsigned a = reinterpret_cast<const uint16_t*>(t.raw.ctl)[lane];
                    ctl[0] = a & 0xff;
                    ctl[S - 1] = a >> 8;
                } else {
                    ctl[0] = reinterpret_cast<const uint8_t*>(t.raw.ctl)[lane];
                }
            }
            __syncwarp();  // the raw buffer is drained
        } else {
            // np is already known: ships, bearings, controls, live planet slots and (T1) the bullets
            // all leave in one batch — one exposed round trip
#pragma unroll
            for (int s = 0; s < S; s++) {
                shv[s] = *reinterpret_cast<const float4*>(&ships[s * 32]);
                sb[s] = ship_b[s * 32];
            }
#pragma unroll
            for (int j = 0; j < ASTRO_MAX_PLANETS; j++) {
                plv[j] = make_float4(kFar, kFar, 0.f, 0.f);
                if (j < np) plv[j] = *reinterpret_cast<const float4*>(&planets[j * 32]);
            }
            if (p.actions) {
                if (S == 2) {
                    unsigned a = reinterpret_cast<const uint16_t*>(p.actions)[g];
                    ctl[0] = a & 0xff;
                    ctl[S - 1] = a >> 8;
                } else {
                    ctl[0] = p.actions[g];
                }
            }
        }
        if (!p.actions) {
            uint32_t h0 = game_key(p.seed, p.first_game + (uint32_t)g);
#pragma unroll
            for (int s = 0; s < S; s++) ctl[s] = action_from_key(h0, p.step, (uint32_t)s);
        }

        // ================= T1: flat bullet list of the tile; stage it with cp.async ==================
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
        if (nonempty) t.cinfo[__popc(ne & lt_mask)] = lane | (my_excl << 5);
        __syncwarp();
        unsigned c0 = 0;
        auto stage_round = [&](unsigned round_base) {
#pragma unroll 1
            for (unsigned w = 0; w < (unsigned)kStageWindows; w++) {
                const unsigned base = round_base + w * 32u;
                if (base >= total) break;
                unsigned rel = my_excl - base;  // huge when the game starts before base
                unsigned starts = __reduce_or_sync(full, (nonempty && rel < 32u) ? (1u << rel) : 0u);
                unsigned idx = c0 + __popc(starts & (full >> (31u - lane)));  // >= 1 for a valid item
                c0 += __popc(starts);
                const unsigned ci = t.cinfo[(idx - 1u) & 31u];  // (stale only for invalid items)
                const bool valid = base + lane < total;
                const unsigned game = ci & 31u, slot = base + lane - (ci >> 5);
                if (valid) cp_async16(&t.bul[w * 32u + lane], &tile_bullets[game * K + slot]);
                t.ref[w * 32u + lane] = valid ? (uint16_t)(game | (slot << 5)) : (uint16_t)0xFFFFu;
            }
            cp_async_commit();
        };
        stage_round(0u);
        // ... and, committed AFTER the bullets so that T3 can wait for the bullets alone, the rows of
        // tile k+1: they stay in flight through T2, T3 and T4
        const bool rows_requested = ROWS && tile_next < n_tiles;
        if (rows_requested) request_rows(tile_next, meta_next);

        // ================= T2: ships and planets while the bullets fly ===============================
        // (the new state goes straight to HBM; the state of a game that ends without auto-reset is
        //  unspecified — the reference has none)
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

        // ================= T3: the bullet loop, from shared memory; next tile's rows in flight =========
        if (rows_requested) cp_async_wait_but_one();  // this tile's bullets; the rows group stays pending
        else cp_async_wait_all();
        __syncwarp();
        for (unsigned round_base = 0; round_base < total; round_base += (unsigned)kStageWindows * 32u) {
            if (round_base) {  // tiles with more than 256 bullets: rare
                stage_round(round_base);
                cp_async_wait_all();
                __syncwarp();
            }
#pragma unroll 1
            for (unsigned w = 0; w < (unsigned)kStageWindows; w++) {
                const unsigned base = round_base + w * 32u;
                if (base >= total) break;
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

        // ================= T4: terminal logic, spawn, bookkeeping ========================================
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
                if (fire) {  // core.py:267-280, from the OLD ship state
                    ev |= ASTRO_EV_FIRED;
                    B4* row = tile_bullets + lane * K;
                    const float4 dv = t.dir[lane], oxy = t.sxy[lane], ov = t.svel[lane];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const float d0 = s == 0 ? dv.x : dv.z, d1 = s == 0 ? dv.y : dv.w;
                        const float sx = s == 0 ? oxy.x : oxy.z, sy = s == 0 ? oxy.y : oxy.w;
                        const float vx = s == 0 ? ov.x : ov.z, vy = s == 0 ? ov.y : ov.w;
                        // fp32 products as in the reference; sums and advance in fp32 too, unless the
                        // newborn lands within the band of the arena bound
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
            // warp totals -> this tile's private slot row in HBM (no block barrier, no contention)
            unsigned mine = warp_totals((int)lane, S, ev, active, spawned, np, nb, m_out);
            unsigned* slot = p.stat_slots + ((size_t)tile * 16u + lane);
            if (lane < ASTRO_N_STATS && mine) atomicAdd(slot, mine);  // RED: fire and forget
        }
        meta = meta_next;
        meta_next = meta_next2;
    }
}
