/*
 * astro_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, float64, CPU restatement of the reference game tick
 *   /root/reference/astro/core.py:215-303 (step) with its helpers
 *   core.py:138-153 (_gravity), :156-168 (_mask), :171-197 (_update_bodies),
 *   :200-212 (_collisions), util.py:87-92 (direction), :145-148 (wrap_unit_square),
 * and of the observation extraction
 *   /root/reference/astro/rl.py:43-99 (get_features, to_batch), util.py:125-132 (norm_angle).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the timed CPU baseline.  The
 * product (astro_b200/, include/) never links, imports or calls it.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file bit-for-bit against
 * vectors produced by the unmodified reference (tests/golden/make_golden.py): 39 full
 * trajectories, 70 hand-built edge cases, the reference tests' own known answers and
 * 42k float32 sin/cos bit patterns.
 *
 * The arithmetic lives in numpy 2.3.5 (un-vendored, requirements.txt:6 unpinned): every
 * operation below is one IEEE-754 binary64 operation per numpy ufunc call, in numpy's
 * evaluation order; build with -ffp-contract=off so the compiler cannot fuse them.  The
 * one non-trivial third-party algorithm is numpy's SIMD float32 sin/cos
 * (numpy/_core/src/umath/loops_trigonometric.dispatch.*: Cody-Waite 3-constant range
 * reduction with FMA, degree-8/9 minimax polynomials) restated in np_sincos_f32 below.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AO_MAXS 2
#define AO_MAXP 4

/* events bitmask (same values as include/astro_b200.h ASTRO_EV_*) */
#define AO_EV_HIT0 1
#define AO_EV_HIT1 2
#define AO_EV_TIMEOUT 4
#define AO_EV_FIRED 8
#define AO_EV_OVERFLOW 16

typedef struct {
    double gravity, dt, max_time, reload_time, bullet_speed, ship_thrust, ship_rspeed,
        ship_radius, planet_mass, planet_radius;
    int32_t solo;
    int32_t reserved;
} ao_config;

/* ---- numpy float32 sin/cos (util.direction, util.py:90-91) --------------------------
 * Valid for |x| <= 71476.0625 (numpy falls back to libm beyond that; bearings in a game
 * stay below 3000*0.08 + 2*pi = 247).                                                  */
static void np_sincos_f32(float x, float* s_out, float* c_out) {
    const float two_over_pi = 0x1.45f306p-1f;
    const float magic = 0x1.8p+23f;
    volatile float qm = fmaf(x, two_over_pi, magic); /* rint(x*2/pi), one rounding */
    float q = qm - magic;
    float r = fmaf(q, -0x1.921fb0p+00f, x);
    r = fmaf(q, -0x1.5110b4p-22f, r);
    r = fmaf(q, -0x1.846988p-48f, r);
    float r2 = r * r;
    float c = fmaf(0x1.98e616p-16f, r2, -0x1.6c06dcp-10f);
    c = fmaf(c, r2, 0x1.55553cp-5f);
    c = fmaf(c, r2, -0x1.000000p-1f);
    c = fmaf(c, r2, 0x1.000000p+0f);
    float s = fmaf(0x1.7d3bbcp-19f, r2, -0x1.a06bbap-13f);
    s = fmaf(s, r2, 0x1.11119ap-7f);
    s = fmaf(s, r2, -0x1.555556p-3f);
    s = fmaf(s, r2, 0.0f); /* numpy's last Horner step adds +0: sin(-0.0) stays -0.0 */
    s = fmaf(s, r, r);
    int iq = (int)q;
    float vs = (iq & 1) ? c : s;
    if (iq & 2) vs = -vs;
    int ic = iq + 1;
    float vc = (ic & 1) ? c : s;
    if (ic & 2) vc = -vc;
    *s_out = vs;
    *c_out = vc;
}

void ao_sincos_f32(const float* x, float* s, float* c, int64_t n) {
    for (int64_t i = 0; i < n; i++) np_sincos_f32(x[i], s + i, c + i);
}

/* numpy float remainder (Python modulo: result takes the sign of the divisor) */
static double np_remainder(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* util.wrap_unit_square, util.py:145-148 */
static double wrap_unit(double x) { return np_remainder(x + 1.0, 2.0) - 1.0; }

/* util.norm_angle, util.py:125-132 */
static double norm_angle(double b) { return np_remainder(b + M_PI, 2.0 * M_PI) - M_PI; }

void ao_wrap_unit_square(const double* x, double* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = wrap_unit(x[i]);
}
void ao_norm_angle(const double* x, double* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = norm_angle(x[i]);
}

/* core._collisions, core.py:200-212: hit_i = any_{j != i} |x_j - x_i|^2 < (r_j + r_i)^2 */
void ao_collisions(const double* x, const double* r, int32_t n, uint8_t* hit) {
    for (int i = 0; i < n; i++) {
        int h = 0;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            double d0 = x[2 * j] - x[2 * i], d1 = x[2 * j + 1] - x[2 * i + 1];
            double rx2 = d0 * d0 + d1 * d1;
            double rr = r[j] + r[i];
            h |= (rx2 < rr * rr);
        }
        hit[i] = (uint8_t)h;
    }
}

/* core._gravity, core.py:138-153: a_i = sum_j (G*M / max(1e-12, |p_j - x_i|^2)) * (p_j - x_i) */
static void gravity(const ao_config* c, int P, const double* planets /*[P][4]*/, double x0, double x1,
                    double* a0, double* a1) {
    double gm = c->gravity * c->planet_mass;
    double s0 = 0, s1 = 0;
    for (int j = 0; j < P; j++) {
        double r0 = planets[4 * j] - x0, r1 = planets[4 * j + 1] - x1;
        double d2 = r0 * r0 + r1 * r1;
        double f = gm / (1e-12 > d2 ? 1e-12 : d2);
        double t0 = f * r0, t1 = f * r1;
        if (j == 0) { s0 = t0; s1 = t1; } else { s0 = s0 + t0; s1 = s1 + t1; }
    }
    *a0 = s0;
    *a1 = s1;
}

/* The same on the arrays core.create returns (core.py:86-135): positions are float32 there, so on a game's FIRST
 * tick numpy evaluates _gravity in float32 (NEP 50: the Python-float constants are weak scalars) — rx, rx**2,
 * the sum, maximum(1e-12, .), the division and the products are float32 operations, one rounding each. */
static void gravity_f32(const ao_config* c, int P, const double* planets /*[P][4], float32 values*/, float x0, float x1,
                        float* a0, float* a1) {
    float gm = (float)(c->gravity * c->planet_mass);
    float s0 = 0, s1 = 0;
    for (int j = 0; j < P; j++) {
        float r0 = (float)planets[4 * j] - x0, r1 = (float)planets[4 * j + 1] - x1;
        float q0 = r0 * r0, q1 = r1 * r1;
        float d2 = q0 + q1;
        float f = gm / (1e-12f > d2 ? 1e-12f : d2);
        float t0 = f * r0, t1 = f * r1;
        if (j == 0) { s0 = t0; s1 = t1; } else { s0 = s0 + t0; s1 = s1 + t1; }
    }
    *a0 = s0;
    *a1 = s1;
}
/* |b - a|^2 of float32 positions, evaluated in float32 (core.py:210 on create()'s arrays); the comparison with the
 * float64 r2 promotes it exactly. */
static double dist2_f32(double ax, double ay, double bx, double by) {
    float d0 = (float)bx - (float)ax, d1 = (float)by - (float)ay;
    float q0 = d0 * d0, q1 = d1 * d1;
    return (double)(q0 + q1);
}

static int in_arena(double x0, double x1) {
    /* core.py:195: (([-1,-1] <= x) & (x <= [1,1])).any(axis=1) */
    return ((-1.0 <= x0) & (x0 <= 1.0)) | ((-1.0 <= x1) & (x1 <= 1.0));
}

/*
 * One game, one tick.  ships [S][5] = x,y,dx,dy,b ; planets [P][4] ; bullets [B][4].
 * Outputs may alias nothing.  bullets_o must hold B+S rows.  Returns 1 when the tick is
 * terminal (reference returns None): outputs other than reward/events are then untouched.
 * bullet_cap < 0 = unbounded (the reference); otherwise bullets beyond the cap are dropped
 * from the END of the reference order (they can only be newborn) and OVERFLOW is flagged —
 * the device pool policy.
 */
/* raw != 0: the state holds the arrays core.create returned (float32 ships, planet positions, no bullets): the
 * reference's first tick then runs _gravity and the squared distances of _collisions in float32, the planets'
 * a * dt too (float32 array times a Python float), and a bullet born on that very tick (reload_time <= dt) is
 * float32 throughout.  Everything else promotes to float64 as on any other tick. */
static int step_impl(const ao_config* c, int S, int P, int B, int bullet_cap, const double* ships,
                const double* planets, const double* bullets, double reload, double t,
                const int64_t* control, double* ships_o, double* planets_o, double* bullets_o,
                int32_t* B_o, double* reload_o, double* t_o, double* reward, int32_t* events, int raw) {
    float dir[AO_MAXS][2];
    double acc[AO_MAXS][2], db[AO_MAXS];
    int ev = 0;

    /* core.py:234-239 */
    for (int i = 0; i < S; i++) {
        np_sincos_f32((float)ships[5 * i + 4], &dir[i][0], &dir[i][1]);
        double g0, g1;
        if (raw) {
            float f0, f1;
            gravity_f32(c, P, planets, (float)ships[5 * i], (float)ships[5 * i + 1], &f0, &f1);
            g0 = (double)f0;
            g1 = (double)f1;
        } else {
            gravity(c, P, planets, ships[5 * i], ships[5 * i + 1], &g0, &g1);
        }
        /* floor mod / floor div of the control code, as numpy % and // on int64 */
        int64_t ctl = control[i];
        int64_t m2 = ((ctl % 2) + 2) % 2;
        int64_t d2 = (ctl - m2) / 2;
        double th = c->ship_thrust * (double)m2;
        acc[i][0] = th * (double)dir[i][0] + g0;
        acc[i][1] = th * (double)dir[i][1] + g1;
        db[i] = (c->dt * c->ship_rspeed) * (double)(d2 - 1);
    }

    /* core.py:241-251 — only ship rows and bullet rows of the mask are ever consumed */
    double rs = c->ship_radius, rp = c->planet_radius;
    double R_ss = (rs + rs) * (rs + rs), R_sp = (rp + rs) * (rp + rs);
    double R_sb = (0.0 + rs) * (0.0 + rs), R_pb = (0.0 + rp) * (0.0 + rp);
    int ship_hit[AO_MAXS] = {0, 0};
    uint8_t* bullet_hit = (uint8_t*)calloc((size_t)(B > 0 ? B : 1), 1);
    for (int i = 0; i < S; i++) {
        double x0 = ships[5 * i], x1 = ships[5 * i + 1];
        for (int j = 0; j < S; j++) {
            if (j == i) continue;
            double d0 = ships[5 * j] - x0, d1 = ships[5 * j + 1] - x1;
            double rx2 = raw ? dist2_f32(x0, x1, ships[5 * j], ships[5 * j + 1]) : d0 * d0 + d1 * d1;
            ship_hit[i] |= (rx2 < R_ss);
        }
        for (int j = 0; j < P; j++) {
            double d0 = planets[4 * j] - x0, d1 = planets[4 * j + 1] - x1;
            double rx2 = raw ? dist2_f32(x0, x1, planets[4 * j], planets[4 * j + 1]) : d0 * d0 + d1 * d1;
            ship_hit[i] |= (rx2 < R_sp);
        }
        for (int j = 0; j < B; j++) {
            double d0 = bullets[4 * j] - x0, d1 = bullets[4 * j + 1] - x1;
            ship_hit[i] |= (d0 * d0 + d1 * d1 < R_sb);
        }
    }
    for (int i = 0; i < B; i++) {
        double x0 = bullets[4 * i], x1 = bullets[4 * i + 1];
        int h = 0;
        for (int j = 0; j < S; j++) {
            double d0 = ships[5 * j] - x0, d1 = ships[5 * j + 1] - x1;
            h |= (d0 * d0 + d1 * d1 < R_sb);
        }
        for (int j = 0; j < P; j++) {
            double d0 = planets[4 * j] - x0, d1 = planets[4 * j + 1] - x1;
            h |= (d0 * d0 + d1 * d1 < R_pb);
        }
        /* bullet-bullet: d2 < (0+0)^2 is never true */
        bullet_hit[i] = (uint8_t)h;
    }

    /* core.py:253-255 — collision terminal first */
    int any = 0;
    for (int i = 0; i < S; i++) any |= ship_hit[i];
    if (any) {
        for (int i = 0; i < S; i++) {
            reward[i] = (double)(1 - 2 * ship_hit[i]);
            if (ship_hit[i]) ev |= (i == 0 ? AO_EV_HIT0 : AO_EV_HIT1);
        }
        *events = ev;
        free(bullet_hit);
        return 1;
    }
    /* core.py:257-260 — timeout */
    if (c->max_time <= t + c->dt) {
        for (int i = 0; i < S; i++) reward[i] = c->solo ? 1.0 : 0.0;
        *events = AO_EV_TIMEOUT;
        free(bullet_hit);
        return 1;
    }

    /* core.py:263-280 — despawn, then spawn from the OLD ship state */
    double next_reload = reload + c->dt;
    int n = 0;
    for (int i = 0; i < B; i++) {
        if (bullet_hit[i]) continue;
        memcpy(bullets_o + 4 * n, bullets + 4 * i, 4 * sizeof(double));
        n++;
    }
    free(bullet_hit);
    if (c->reload_time <= next_reload) {
        float off = (float)(1.001 * c->ship_radius); /* python float -> weak scalar -> float32 */
        float spd = (float)c->bullet_speed;
        for (int i = 0; i < S; i++) {
            float o0 = off * dir[i][0], o1 = off * dir[i][1]; /* float32 products */
            float v0 = spd * dir[i][0], v1 = spd * dir[i][1];
            if (raw) { /* float32 ship state + float32 product: a float32 sum */
                bullets_o[4 * n + 0] = (double)((float)ships[5 * i + 0] + o0);
                bullets_o[4 * n + 1] = (double)((float)ships[5 * i + 1] + o1);
                bullets_o[4 * n + 2] = (double)((float)ships[5 * i + 2] + v0);
                bullets_o[4 * n + 3] = (double)((float)ships[5 * i + 3] + v1);
            } else {
                bullets_o[4 * n + 0] = ships[5 * i + 0] + (double)o0;
                bullets_o[4 * n + 1] = ships[5 * i + 1] + (double)o1;
                bullets_o[4 * n + 2] = ships[5 * i + 2] + (double)v0;
                bullets_o[4 * n + 3] = ships[5 * i + 3] + (double)v1;
            }
            n++;
        }
        next_reload -= c->reload_time;
        ev |= AO_EV_FIRED;
    }

    /* core.py:282-302 — symplectic Euler (core.py:189-197) */
    double dt = c->dt;
    for (int i = 0; i < S; i++) {
        double v0 = ships[5 * i + 2] + acc[i][0] * dt, v1 = ships[5 * i + 3] + acc[i][1] * dt;
        double x0 = ships[5 * i + 0] + dt * v0, x1 = ships[5 * i + 1] + dt * v1;
        ships_o[5 * i + 0] = wrap_unit(x0);
        ships_o[5 * i + 1] = wrap_unit(x1);
        ships_o[5 * i + 2] = v0;
        ships_o[5 * i + 3] = v1;
        ships_o[5 * i + 4] = ships[5 * i + 4] + db[i];
    }
    for (int i = 0; i < P; i++) {
        double g0, g1, v0, v1;
        if (raw) { /* float32 field times the weak scalar dt: a float32 product, then the float64 sum */
            float f0, f1;
            gravity_f32(c, P, planets, (float)planets[4 * i], (float)planets[4 * i + 1], &f0, &f1);
            float p0 = f0 * (float)dt, p1 = f1 * (float)dt;
            v0 = planets[4 * i + 2] + (double)p0;
            v1 = planets[4 * i + 3] + (double)p1;
        } else {
            gravity(c, P, planets, planets[4 * i], planets[4 * i + 1], &g0, &g1);
            v0 = planets[4 * i + 2] + g0 * dt;
            v1 = planets[4 * i + 3] + g1 * dt;
        }
        double x0 = planets[4 * i + 0] + dt * v0, x1 = planets[4 * i + 1] + dt * v1;
        planets_o[4 * i + 0] = wrap_unit(x0);
        planets_o[4 * i + 1] = wrap_unit(x1);
        planets_o[4 * i + 2] = v0;
        planets_o[4 * i + 3] = v1;
    }
    double zero_dt = 0.0 * dt; /* a=0 (core.py:297): 0 * dt */
    int m = 0;
    for (int i = 0; i < n; i++) {
        double v0 = bullets_o[4 * i + 2] + zero_dt, v1 = bullets_o[4 * i + 3] + zero_dt;
        double x0 = bullets_o[4 * i + 0] + dt * v0, x1 = bullets_o[4 * i + 1] + dt * v1;
        if (raw) { /* (only newborn exist on a first tick) float32 arrays and weak scalars: float32 operations */
            float w0 = (float)bullets_o[4 * i + 2] + (float)zero_dt, w1 = (float)bullets_o[4 * i + 3] + (float)zero_dt;
            float m0 = (float)dt * w0, m1 = (float)dt * w1;
            float y0 = (float)bullets_o[4 * i + 0] + m0, y1 = (float)bullets_o[4 * i + 1] + m1;
            v0 = (double)w0; v1 = (double)w1; x0 = (double)y0; x1 = (double)y1;
        }
        if (!in_arena(x0, x1)) continue;
        bullets_o[4 * m + 0] = x0;
        bullets_o[4 * m + 1] = x1;
        bullets_o[4 * m + 2] = v0;
        bullets_o[4 * m + 3] = v1;
        m++;
    }
    if (bullet_cap >= 0 && m > bullet_cap) {
        m = bullet_cap;
        ev |= AO_EV_OVERFLOW;
    }
    *B_o = m;
    *reload_o = next_reload;
    *t_o = t + dt;
    for (int i = 0; i < S; i++) reward[i] = 0.0;
    *events = ev;
    return 0;
}

int ao_step_one(const ao_config* c, int S, int P, int B, int bullet_cap, const double* ships,
                const double* planets, const double* bullets, double reload, double t,
                const int64_t* control, double* ships_o, double* planets_o, double* bullets_o,
                int32_t* B_o, double* reload_o, double* t_o, double* reward, int32_t* events) {
    return step_impl(c, S, P, B, bullet_cap, ships, planets, bullets, reload, t, control, ships_o, planets_o, bullets_o, B_o,
                     reload_o, t_o, reward, events, 0);
}
/* The first tick after core.create, on create()'s own dtypes (see step_impl). */
int ao_step_one_raw(const ao_config* c, int S, int P, int B, int bullet_cap, const double* ships,
                    const double* planets, const double* bullets, double reload, double t,
                    const int64_t* control, double* ships_o, double* planets_o, double* bullets_o,
                    int32_t* B_o, double* reload_o, double* t_o, double* reward, int32_t* events) {
    return step_impl(c, S, P, B, bullet_cap, ships, planets, bullets, reload, t, control, ships_o, planets_o, bullets_o, B_o,
                     reload_o, t_o, reward, events, 1);
}

/*
 * Batch of n independent games in fixed-stride arrays (the host-side image of the device
 * pools): ships [n][S][5], planets [n][AO_MAXP][4], np [n], bullets [n][K][4], nb [n],
 * reload [n], t [n], control [n][S] -> same-shaped outputs, reward [n][S], done [n],
 * events [n].  Games whose alive[i] == 0 are skipped (done=1, reward 0, state copied).
 * Terminal games: state outputs are a copy of the inputs.  K is also the bullet cap.
 */
void ao_step_batch(const ao_config* c, int64_t n, int S, int K, const double* ships, const double* planets,
                   const int32_t* np_, const double* bullets, const int32_t* nb, const double* reload,
                   const double* t, const int64_t* control, const uint8_t* alive, double* ships_o,
                   double* planets_o, double* bullets_o, int32_t* nb_o, double* reload_o, double* t_o,
                   double* reward, uint8_t* done, int32_t* events, int nthreads) {
    (void)nthreads; /* threading is done by the caller over disjoint game slices */
    {
        double* tmp = (double*)malloc(sizeof(double) * 4 * (size_t)(K + AO_MAXS));
        for (int64_t i = 0; i < n; i++) {
            const double* sh = ships + i * S * 5;
            const double* pl = planets + i * AO_MAXP * 4;
            const double* bl = bullets + i * (int64_t)K * 4;
            double* sho = ships_o + i * S * 5;
            double* plo = planets_o + i * AO_MAXP * 4;
            double* blo = bullets_o + i * (int64_t)K * 4;
            int term = 1;
            int32_t ev = 0, bo = nb[i];
            double ro = reload[i], to = t[i];
            for (int s = 0; s < S; s++) reward[i * S + s] = 0.0;
            if (alive == NULL || alive[i]) {
                memcpy(sho, sh, sizeof(double) * S * 5);
                memcpy(plo, pl, sizeof(double) * AO_MAXP * 4);
                term = ao_step_one(c, S, np_[i], nb[i], K, sh, pl, bl, reload[i], t[i], control + i * S, sho,
                                   plo, tmp, &bo, &ro, &to, reward + i * S, &ev);
                if (!term) {
                    memcpy(blo, tmp, sizeof(double) * 4 * (size_t)bo);
                } else {
                    bo = nb[i];
                    memcpy(blo, bl, sizeof(double) * 4 * (size_t)bo);
                }
            } else {
                memcpy(sho, sh, sizeof(double) * S * 5);
                memcpy(plo, pl, sizeof(double) * AO_MAXP * 4);
                memcpy(blo, bl, sizeof(double) * 4 * (size_t)nb[i]);
            }
            nb_o[i] = bo;
            reload_o[i] = ro;
            to = term ? t[i] : to;
            t_o[i] = to;
            done[i] = (uint8_t)term;
            events[i] = ev;
        }
        free(tmp);
    }
}

/*
 * rl.ValueNetwork.get_features (rl.py:43-72) + to_batch padding (rl.py:91-98) for one state
 * seen from ship `me` (core.roll_ships, core.py:306-327): out [n_rows][1+5S+4] float32,
 * rows = planets then bullets, remaining rows filled with -1.  Returns P+B, or -1 if it
 * does not fit in n_rows.
 */
int ao_features(int S, int P, int B, const double* ships, const double* planets, const double* bullets,
                int me, int n_rows, float* out) {
    int D = 1 + 5 * S + 4;
    if (P + B > n_rows) return -1;
    float sf[5 * AO_MAXS];
    for (int k = 0; k < S; k++) {
        const double* sh = ships + 5 * ((k + me) % S); /* np.roll(..., -me) */
        sf[5 * k + 0] = (float)sh[0];
        sf[5 * k + 1] = (float)sh[1];
        sf[5 * k + 2] = (float)sh[2];
        sf[5 * k + 3] = (float)sh[3];
        sf[5 * k + 4] = (float)(norm_angle(sh[4]) / M_PI);
    }
    for (int r = 0; r < n_rows; r++) {
        float* row = out + (int64_t)r * D;
        if (r >= P + B) {
            for (int k = 0; k < D; k++) row[k] = -1.0f;
            continue;
        }
        const double* obj = r < P ? planets + 4 * r : bullets + 4 * (r - P);
        row[0] = r < P ? 0.0f : 1.0f;
        for (int k = 0; k < 5 * S; k++) row[1 + k] = sf[k];
        for (int k = 0; k < 4; k++) row[1 + 5 * S + k] = (float)obj[k];
    }
    return P + B;
}

/*
 * script.ScriptBot.__call__ (script.py:67-91) with _danger (:41-65) and _fly_to (:30-39) for the
 * ship `me` of one game (the bot sees roll_ships(state, me), core.py:306-327).  float64 libm
 * (sqrt, atan2, fmod) like numpy on this platform.  Returns the control code 0..5.
 * Quirk kept: inside _danger the parameter `b` (bearing) is shadowed by the quadratic
 * coefficient (script.py:54), so `rotation` uses that coefficient.
 */
static int fly_to(double target, double my_b, double t, int fwd) {
    double angle = norm_angle(target - my_b);
    if (angle < -t) return 0;
    if (t < angle) return 4;
    return fwd ? 3 : 2;
}
int ao_script_control(const ao_config* c, int S, int P, const double* ships, const double* planets, int me,
                      double avoid_distance, double avoid_threshold) {
    const double* my = ships + 5 * me;
    for (int i = 0; i < P; i++) {
        double x0 = my[0] - planets[4 * i], x1 = my[1] - planets[4 * i + 1];
        double v0 = my[2] - planets[4 * i + 2], v1 = my[3] - planets[4 * i + 3];
        double radius = c->planet_radius + c->ship_radius;
        double speed = sqrt(v0 * v0 + v1 * v1);               /* util.mag(dx) */
        double n0 = v0 / (speed + 1e-12), n1 = v1 / (speed + 1e-12); /* util.norm(dx) */
        double b = 2 * (n0 * x0 + n1 * x1);                    /* 2 * dot(norm(dx), x) */
        double mx = sqrt(x0 * x0 + x1 * x1);
        double ra = radius + avoid_distance;
        double cc = mx * mx - ra * ra;
        double det = b * b - 4 * cc;
        if (0 < det && 0 <= -b + sqrt(det)) {
            double distance = -b - sqrt(det);
            double bear = atan2(x0, x1);                       /* util.bearing(x) */
            double rotation = fabs(norm_angle(bear - b));
            if (distance < (speed / c->ship_thrust + c->ship_rspeed / rotation) * speed)
                return fly_to(bear, my[4], avoid_threshold, 1);
        }
    }
    if (c->solo || S < 2) return 2;
    const double* en = ships + 5 * ((me + 1) % S);
    double e0 = en[0] - my[0], e1 = en[1] - my[1];
    double enemy_distance = sqrt(e0 * e0 + e1 * e1);
    double bullet_time = enemy_distance / c->bullet_speed;
    double f0 = en[0] + bullet_time * (en[2] - my[2]), f1 = en[1] + bullet_time * (en[3] - my[3]);
    return fly_to(atan2(f0 - my[0], f1 - my[1]), my[4], c->ship_radius / enemy_distance, 0);
}
void ao_script_batch(const ao_config* c, int64_t n, int S, const double* ships, const double* planets,
                     const int32_t* np_, double avoid_distance, double avoid_threshold, int64_t* out) {
    for (int64_t i = 0; i < n; i++)
        for (int me = 0; me < S; me++)
            out[i * S + me] = ao_script_control(c, S, np_[i], ships + i * S * 5, planets + i * AO_MAXP * 4, me,
                                                avoid_distance, avoid_threshold);
}

/* ---- counter-based streams (astro_b200/rng.py) --------------------------------------- */
static uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
static inline int action_for(uint32_t seed, uint32_t game, uint32_t step, uint32_t ship) {
    uint32_t h0 = mix32(seed ^ (game * 0x9E3779B1u));
    uint32_t h = mix32(h0 ^ (step * 2u + ship));
    return (int)(((uint64_t)h * 6u) >> 32);
}
static inline uint32_t pool_pick(uint32_t seed, uint32_t game, uint32_t key, uint32_t pool_size) {
    uint32_t h0 = mix32(seed ^ 0xA5A5A5A5u ^ (game * 0x9E3779B1u));
    uint32_t h = mix32(h0 ^ key);
    return (uint32_t)(((uint64_t)h * pool_size) >> 32);
}

/*
 * CPU baseline rollout (bench.py cpu_baseline / --impl reference): n games, n_ticks ticks,
 * counter-stream controls, auto-reset from a pool of M initial states built by create().
 * State is updated in place.  stats[8]: episodes, wins0, wins1, both_lost, timeouts,
 * env_steps, bullets_spawned, overflow;  returns env-steps executed.
 */
int64_t ao_rollout(const ao_config* c, int64_t n, int S, int K, double* ships, double* planets, int32_t* np_,
                   double* bullets, int32_t* nb, double* reload, double* t, uint32_t* episode,
                   int64_t M, const double* pool_ships, const double* pool_planets, const int32_t* pool_np,
                   uint32_t seed, int64_t first_game, uint32_t step0, int32_t n_ticks, int nthreads,
                   int64_t* stats) {
    int64_t tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    (void)nthreads; /* threading is done by the caller over disjoint game slices */
    {
        int64_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double* tmp = (double*)malloc(sizeof(double) * 4 * (size_t)(K + AO_MAXS));
        for (int64_t i = 0; i < n; i++) {
            double* sh = ships + i * S * 5;
            double* pl = planets + i * AO_MAXP * 4;
            double* bl = bullets + i * (int64_t)K * 4;
            uint32_t g = (uint32_t)(first_game + i);
            for (int32_t k = 0; k < n_ticks; k++) {
                int64_t ctl[AO_MAXS];
                for (int s = 0; s < S; s++) ctl[s] = action_for(seed, g, step0 + (uint32_t)k, (uint32_t)s);
                double sho[5 * AO_MAXS], plo[4 * AO_MAXP], rew[AO_MAXS], ro, to;
                int32_t bo, ev;
                int term = ao_step_one(c, S, np_[i], nb[i], K, sh, pl, bl, reload[i], t[i], ctl, sho, plo, tmp,
                                       &bo, &ro, &to, rew, &ev);
                loc[5]++;
                if (!term) {
                    memcpy(sh, sho, sizeof(double) * 5 * S);
                    memcpy(pl, plo, sizeof(double) * 4 * np_[i]);
                    memcpy(bl, tmp, sizeof(double) * 4 * (size_t)bo);
                    nb[i] = bo; reload[i] = ro; t[i] = to;
                    if (ev & AO_EV_FIRED) loc[6] += S;
                    if (ev & AO_EV_OVERFLOW) loc[7]++;
                } else {
                    loc[0]++;
                    if (ev & AO_EV_TIMEOUT) loc[4]++;
                    else if (S == 2 && rew[0] > 0) loc[1]++;
                    else if (S == 2 && rew[1] > 0) loc[2]++;
                    else loc[3]++;
                    if (M > 0) {
                        episode[i]++;
                        /* key = 1 + the rollout step that ended the game (0 = initial fill) */
                        int64_t p = pool_pick(seed, g, step0 + (uint32_t)k + 1u, (uint32_t)M);
                        memcpy(sh, pool_ships + p * S * 5, sizeof(double) * 5 * S);
                        memcpy(pl, pool_planets + p * AO_MAXP * 4, sizeof(double) * 4 * AO_MAXP);
                        np_[i] = pool_np[p]; nb[i] = 0; reload[i] = 0.0; t[i] = 0.0;
                    } else {
                        break;
                    }
                }
            }
        }
        free(tmp);
        for (int k = 0; k < 8; k++) tot[k] += loc[k];
    }
    if (stats) for (int k = 0; k < 8; k++) stats[k] = tot[k];
    return tot[5];
}

int ao_actions(uint32_t seed, int64_t first_game, int64_t n, uint32_t step, int S, int64_t* out) {
    for (int64_t i = 0; i < n; i++)
        for (int s = 0; s < S; s++) out[i * S + s] = action_for(seed, (uint32_t)(first_game + i), step, (uint32_t)s);
    return 0;
}
int ao_pool_pick(uint32_t seed, int64_t first_game, int64_t n, const uint32_t* episode, uint32_t M, int64_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = pool_pick(seed, (uint32_t)(first_game + i), episode[i], M);
    return 0;
}
