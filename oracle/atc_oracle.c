/*
 * atc_oracle.c — CPU restatement of the reference's AtcGym.step()/reset() path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library, and only as the checker / the CPU baseline.
 * The product (atc_reinforcement_learning_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED for 1 aircraft / no wind — checked against golden vectors recorded from the
 * live, unmodified reference (tests/golden/, produced by oracle/make_golden.py) and against the
 * reference's own 8 unit tests (envs/atc/model_test.py).  UNPINNED for the extensions the reference
 * does not contain (several aircraft per env, 3 nm / 1000 ft separation, wind, device spawn RNG):
 * those follow this repo's own spec and are proven to degenerate exactly to the
 * pinned path at 1 aircraft / zero wind.
 *
 * Everything is IEEE double like the reference (Python floats); the observation is cast to float32
 * where the reference casts it (atc_gym.py:270-276) and normalised in float32 (atc_gym.py:187-189).
 * Compile with -ffp-contract=off so no multiply-add is fused.
 *
 * All file:line citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_AC 8
#define NM_TO_FT 6076.0          /* model.py:10 */
#define TIMESTEP_LIMIT 6000      /* atc_gym.py:40 */

enum { TERM_RUNNING = 0, TERM_BELOW_MVA = 1, TERM_LEFT_AIRSPACE = 2, TERM_CAPTURED = 3, TERM_TIMEOUT = 4,
       TERM_SEPARATION = 5 };

typedef struct {
    /* static sector (scenarios.py) */
    int n_mva;
    double *ring_xy;     /* closed rings, 2 doubles per vertex */
    int *ring_off;       /* n_mva + 1 */
    double *height;      /* n_mva */
    double *bounds;      /* n_mva * 4: minx, miny, maxx, maxy (model.py:268) */
    double rwy_x, rwy_y, rwy_h, phi_from, phi_to;
    int n_entry;
    double *entry;       /* n_entry * 3: x, y, phi */
    int *level_off;      /* n_entry + 1 */
    int *levels;
    /* derived (model.py:155-186, atc_gym.py:49-58,88-110) */
    double faf[2], iaf[2], corner1[2], corner2[2], normal[2];
    double tri_h[8], tri_1[8], tri_2[8];   /* closed 4-vertex rings */
    double bbox[4], dmax, faf_mva;
    float nmin[10], nmax[10];
    /* sim parameters (model.py:132-145) */
    double dt;
    int shaping, normalize, discrete, normalize_reset_obs;
    /* wind extension */
    int wind_gx, wind_gy;
    double *wind;        /* gy * gx * 2, knots */
    /* batch */
    int n_env, n_ac;
    uint64_t seed;
    int64_t env_index_base;
    double *x, *y, *h, *phi, *v;      /* [n_env * n_ac] */
    double *last_action;              /* [n_env * n_ac * 3] */
    int32_t *timesteps, *actions_taken, *episodes, *win_ring, *last_ep_len;
    double *ep_return, *last_ep_return;
} Oracle;

/* ------------------------------------------------------------------------------------------ geometry */

/* model.py:318-337 — even-odd ray casting over the CLOSED ring, edges (poly[i-1], poly[i % n]), i = 0..n */
static int ray_tracing(double x, double y, const double *poly, int n)
{
    int inside = 0;
    double xints = 0.0;
    double p1x = poly[0], p1y = poly[1];
    for (int i = 0; i <= n; ++i) {
        double p2x = poly[2 * (i % n)], p2y = poly[2 * (i % n) + 1];
        if (y > fmin(p1y, p2y)) {
            if (y <= fmax(p1y, p2y)) {
                if (x <= fmax(p1x, p2x)) {
                    if (p1y != p2y)
                        xints = (y - p1y) * (p2x - p1x) / (p2y - p1y) + p1x;
                    if (p1x == p2x || x <= xints)
                        inside = !inside;
                }
            }
        }
        p1x = p2x;
        p1y = p2y;
    }
    return inside;
}

/* model.py:282-292 — first polygon in list order whose inclusive bbox and ring contain the point; -1 = outside */
static int find_mva(const Oracle *o, double x, double y)
{
    for (int m = 0; m < o->n_mva; ++m) {
        const double *b = o->bounds + 4 * m;
        if (b[0] <= x && x <= b[2] && b[1] <= y && y <= b[3])
            if (ray_tracing(x, y, o->ring_xy + 2 * o->ring_off[m], o->ring_off[m + 1] - o->ring_off[m]))
                return m;
    }
    return -1;
}

/* Python float modulo (result has the sign of the divisor) — used by model.py:340-342 */
static double pymod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0)))
        r += b;
    return r;
}

/* model.py:340-342 */
static double relative_angle(double a1, double a2)
{
    return pymod(a2 - a1 + 180.0, 360.0) - 180.0;
}

static double radians_(double deg) { return deg * (M_PI / 180.0); }   /* math.radians */

/* model.py:212-231.  The reference's if/elif (model.py:224-229): the elif branch IS evaluated whenever the first
 * condition as a whole is false (a conjunction), so both sides must be tested independently. */
static int inside_corridor_angle_full(const Oracle *o, double x, double y, double phi)
{
    double tr = o->phi_to;
    double a0 = sin(radians_(tr)), a1 = cos(radians_(tr));
    double b0 = sin(radians_(phi)), b1 = cos(radians_(phi));
    /* np.dot of two 2-vectors goes through BLAS ddot, which on the machine the golden vectors were recorded on
     * (OpenBLAS 0.3.30, Haswell kernel) evaluates fma(a1, b1, a0 * b0) — verified on 200 000 headings.  It only
     * matters when |relative angle| < ~2e-8 deg, where acos(dot) is rounding noise. */
    double dot = fma(a1, b1, a0 * b0);
    double beta = 45.0 - acos(dot);
    double min_angle = 45.0 - beta;
    int ok = 0;
    if (ray_tracing(x, y, o->tri_1, 4)) {
        double r = relative_angle(tr, phi);
        if (min_angle <= r && r <= 45.0)
            ok = 1;
    }
    if (!ok && ray_tracing(x, y, o->tri_2, 4)) {
        double r = relative_angle(phi, tr);
        if (min_angle <= r && r <= 45.0)
            ok = 1;
    }
    return ok;
}

/* model.py:188-210 */
static int inside_corridor(const Oracle *o, double x, double y, double h, double phi)
{
    if (!ray_tracing(x, y, o->tri_h, 4))
        return 0;
    double t = fma(y - o->faf[1], o->normal[1], (x - o->faf[0]) * o->normal[0]);   /* BLAS ddot order, see below */
    double px = o->faf[0] + t * o->normal[0];
    double py = o->faf[1] + t * o->normal[1];
    double dx = px - o->rwy_x, dy = py - o->rwy_y;
    /* np.linalg.norm -> sqrt(x.dot(x)), BLAS ddot order */
    double h_max = sqrt(fma(dy, dy, dx * dx)) * tan(3.0 * M_PI / 180.0) * NM_TO_FT + o->rwy_h;
    if (!(h <= h_max))
        return 0;
    return inside_corridor_angle_full(o, x, y, phi);
}

/* ------------------------------------------------------------------------------------------ constants */

static void derive_constants(Oracle *o)
{
    /* model.py:155-186 */
    double sf = sin(radians_(o->phi_from)), cf = cos(radians_(o->phi_from));
    o->phi_to = pymod(o->phi_from + 180.0, 360.0);
    o->normal[0] = cf * 0.0 + sf * 1.0;
    o->normal[1] = -sf * 0.0 + cf * 1.0;
    double faf_dist = 7.4, iaf_dist = 3.0;
    double corner = iaf_dist / cos(radians_(45.0));
    o->faf[0] = o->rwy_x + (cf * 0.0 + sf * faf_dist);
    o->faf[1] = o->rwy_y + (-sf * 0.0 + cf * faf_dist);
    double ix = cf * 0.0 + sf * corner, iy = -sf * 0.0 + cf * corner;
    double s45 = sin(radians_(45.0)), c45 = cos(radians_(45.0));
    double sm = sin(radians_(-45.0)), cm = cos(radians_(-45.0));
    o->corner1[0] = (c45 * ix + s45 * iy) + o->faf[0];
    o->corner1[1] = (-s45 * ix + c45 * iy) + o->faf[1];
    o->corner2[0] = (cm * ix + sm * iy) + o->faf[0];
    o->corner2[1] = (-sm * ix + cm * iy) + o->faf[1];
    o->iaf[0] = o->rwy_x + (cf * 0.0 + sf * (faf_dist + iaf_dist));
    o->iaf[1] = o->rwy_y + (-sf * 0.0 + cf * (faf_dist + iaf_dist));
    const double *P[3][3] = {{o->faf, o->corner1, o->corner2}, {o->faf, o->corner1, o->iaf}, {o->faf, o->corner2, o->iaf}};
    double *R[3] = {o->tri_h, o->tri_1, o->tri_2};
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 4; ++i) {
            R[k][2 * i] = P[k][i % 3][0];
            R[k][2 * i + 1] = P[k][i % 3][1];
        }
    /* per-polygon bounds (model.py:268) and union bbox (model.py:294-306) */
    o->bbox[0] = o->bbox[1] = INFINITY;
    o->bbox[2] = o->bbox[3] = -INFINITY;
    for (int m = 0; m < o->n_mva; ++m) {
        double *b = o->bounds + 4 * m;
        b[0] = b[1] = INFINITY;
        b[2] = b[3] = -INFINITY;
        for (int i = o->ring_off[m]; i < o->ring_off[m + 1]; ++i) {
            b[0] = fmin(b[0], o->ring_xy[2 * i]);
            b[2] = fmax(b[2], o->ring_xy[2 * i]);
            b[1] = fmin(b[1], o->ring_xy[2 * i + 1]);
            b[3] = fmax(b[3], o->ring_xy[2 * i + 1]);
        }
        o->bbox[0] = fmin(o->bbox[0], b[0]);
        o->bbox[1] = fmin(o->bbox[1], b[1]);
        o->bbox[2] = fmax(o->bbox[2], b[2]);
        o->bbox[3] = fmax(o->bbox[3], b[3]);
    }
    /* atc_gym.py:49-58 */
    int m = find_mva(o, o->faf[0], o->faf[1]);
    o->faf_mva = m >= 0 ? o->height[m] : NAN;
    double lx = o->bbox[2] - o->bbox[0], ly = o->bbox[3] - o->bbox[1];
    o->dmax = hypot(lx, ly);
    /* atc_gym.py:88-110 */
    float nmin[10] = {(float)o->bbox[0], (float)o->bbox[1], 0, 0, 100, 0, 0, 0, -180, -180};
    float nmax[10] = {(float)lx, (float)ly, 38000, 360, 200, 38000, 38000, (float)o->dmax, 360, 360};
    memcpy(o->nmin, nmin, sizeof nmin);
    memcpy(o->nmax, nmax, sizeof nmax);
}

/* ------------------------------------------------------------------------------------------ spawn RNG (own spec) */

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* counter = (global env index lo, hi, episode index, call), key = seed */
static void spawn_env(Oracle *o, int e)
{
    int A = o->n_ac, E = o->n_entry;
    uint64_t gid = (uint64_t)(o->env_index_base + e);
    uint32_t ep = (uint32_t)o->episodes[e];
    uint32_t rnd[2 * ORACLE_MAX_AC];
    for (int c = 0; c < (2 * A + 3) / 4; ++c) {
        uint32_t w[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), ep, (uint32_t)c, (uint32_t)o->seed, (uint32_t)(o->seed >> 32), w);
        for (int k = 0; k < 4 && 4 * c + k < 2 * A; ++k)
            rnd[4 * c + k] = w[k];
    }
    uint32_t used = 0;
    for (int a = 0; a < A; ++a) {
        int ent;
        if (E >= A) {       /* without replacement: j-th still-unused entry point */
            int j = (int)(((uint64_t)rnd[2 * a] * (uint64_t)(E - a)) >> 32);
            ent = 0;
            for (int i = 0; i < E; ++i) {
                if (used & (1u << i))
                    continue;
                if (j == 0) { ent = i; break; }
                --j;
            }
            used |= 1u << ent;
        } else {
            ent = (int)(((uint64_t)rnd[2 * a] * (uint64_t)E) >> 32);
        }
        int L = o->level_off[ent + 1] - o->level_off[ent];
        int k = (int)(((uint64_t)rnd[2 * a + 1] * (uint64_t)L) >> 32);
        int i = e * A + a;
        o->x[i] = o->entry[3 * ent];
        o->y[i] = o->entry[3 * ent + 1];
        o->phi[i] = o->entry[3 * ent + 2];
        o->h[i] = (double)(o->levels[o->level_off[ent] + k] * 100);   /* atc_gym.py:348 */
        o->v[i] = 250.0;
    }
}

/* ------------------------------------------------------------------------------------------ observation / reward */

typedef struct { double d_faf, phi_rel_faf, on_gp; } ObsAux;

/* atc_gym.py:262-297 */
static void get_state(const Oracle *o, int i, double mva, float *raw, ObsAux *aux)
{
    double to_x = o->faf[0] - o->x[i], to_y = o->faf[1] - o->y[i];
    double phi_rel_runway = relative_angle(o->phi_to, o->phi[i]);
    aux->d_faf = hypot(to_x, to_y);
    aux->phi_rel_faf = atan2(to_y, to_x) * (180.0 / M_PI);    /* np.degrees */
    aux->on_gp = 318.4 * aux->d_faf + o->faf_mva - 200.0;
    raw[0] = (float)o->x[i];
    raw[1] = (float)o->y[i];
    raw[2] = (float)o->h[i];
    raw[3] = (float)o->phi[i];
    raw[4] = (float)o->v[i];
    raw[5] = (float)(o->h[i] - mva);
    raw[6] = (float)aux->on_gp;
    raw[7] = (float)aux->d_faf;
    raw[8] = (float)aux->phi_rel_faf;
    raw[9] = (float)phi_rel_runway;
}

/* atc_gym.py:187-189 — float32 arithmetic, numpy operation order */
static void normalize_obs(const Oracle *o, const float *raw, float *out)
{
    for (int k = 0; k < 10; ++k) {
        volatile float a = raw[k] - o->nmin[k];
        volatile float hm = 0.5f * o->nmax[k];
        volatile float b = a - hm;
        out[k] = b / hm;
    }
}

/* atc_gym.py:17-19 */
static double sigmoid_distance(double d, double d_max) { return (1.0 - tanh(4.0 * (d / d_max) - 2.0)) / 2.0; }

static double sign_(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : v)); }

/* ------------------------------------------------------------------------------------------ wind (own spec) */

static void wind_at(const Oracle *o, double x, double y, double *wx, double *wy)
{
    int gx = o->wind_gx, gy = o->wind_gy;
    double sx = (double)(gx - 1) / (o->bbox[2] - o->bbox[0]), sy = (double)(gy - 1) / (o->bbox[3] - o->bbox[1]);
    double fx = (x - o->bbox[0]) * sx, fy = (y - o->bbox[1]) * sy;
    fx = fx > 0.0 ? fx : 0.0;
    fy = fy > 0.0 ? fy : 0.0;
    fx = fx < (double)(gx - 1) ? fx : (double)(gx - 1);
    fy = fy < (double)(gy - 1) ? fy : (double)(gy - 1);
    int i0 = (int)fx, j0 = (int)fy;
    if (i0 > gx - 2) i0 = gx - 2;
    if (j0 > gy - 2) j0 = gy - 2;
    double tx = fx - (double)i0, ty = fy - (double)j0;
    const double *w00 = o->wind + 2 * (j0 * gx + i0), *w10 = w00 + 2, *w01 = w00 + 2 * gx, *w11 = w01 + 2;
    double ux = 1.0 - tx, uy = 1.0 - ty;
    *wx = (w00[0] * ux + w10[0] * tx) * uy + (w01[0] * ux + w11[0] * tx) * ty;
    *wy = (w00[1] * ux + w10[1] * tx) * uy + (w01[1] * ux + w11[1] * tx) * ty;
}

/* ------------------------------------------------------------------------------------------ step */

static double tree_sum(const double *v, int n)
{
    /* xor-butterfly order over the next power of two, zero padded — the order a warp-shuffle reduction uses */
    double buf[ORACLE_MAX_AC] = {0};
    int g = 1;
    while (g < n) g <<= 1;
    for (int i = 0; i < n; ++i) buf[i] = v[i];
    for (int s = 1; s < g; s <<= 1)
        for (int i = 0; i < g; i += 2 * s)
            buf[i] = buf[i] + buf[i + s];
    return buf[0];
}

/* one aircraft: actions + move (atc_gym.py:137-143; model.py:60-129).  returns accumulated base reward */
static double aircraft_advance(Oracle *o, int i, const float *act)
{
    const double dt = o->dt;
    double reward = -0.05 * dt;                                  /* atc_gym.py:137 */
    const double disc[3] = {5.0, 50.0, 0.5};                     /* atc_gym.py:84 */
    const double fac_c[3] = {200.0, 38000.0, 360.0}, fac_d[3] = {10.0, 100.0, 1.0}, off[3] = {100.0, 0.0, 0.0};
    double *st[3] = {&o->v[i], &o->h[i], &o->phi[i]};
    const double lo[3] = {-5.0 * dt, -41.0 * dt, -3.0 * dt}, hi[3] = {5.0 * dt, 15.0 * dt, 3.0 * dt};   /* model.py:45-50 */
    const double vmin[3] = {100.0, 0.0, 0.0}, vmax[3] = {300.0, 38000.0, 0.0};
    for (int k = 0; k < 3; ++k) {
        double a = (double)act[k], target;
        if (o->discrete)
            target = a * fac_d[k] + off[k];                      /* atc_gym.py:327-330 */
        else
            target = a * fac_c[k] / 2.0 + fac_c[k] / 2.0 + off[k];   /* atc_gym.py:333-335 */
        if (k < 2 && (target < vmin[k] || target > vmax[k])) {   /* model.py:69-72,91-94; phi unvalidated */
            reward -= 1.0;                                       /* atc_gym.py:312-315 */
            continue;
        }
        double delta = target - *st[k];
        delta = delta < hi[k] ? delta : hi[k];                   /* min(delta, hi)  (model.py:75,97,115) */
        delta = delta > lo[k] ? delta : lo[k];                   /* max(delta, lo) */
        *st[k] = *st[k] + delta;
        if (!(fabs(target - o->last_action[3 * i + k]) < disc[k]))   /* atc_gym.py:305-306 */
            o->actions_taken[i / o->n_ac] += 1;
        o->last_action[3 * i + k] = target;
    }
    /* model.py:122-129 */
    double d = (o->v[i] / 3600.0) * dt;
    double rad = radians_(o->phi[i]);
    double dx = d * sin(rad), dy = d * cos(rad);
    if (o->wind) {
        double wx, wy;
        wind_at(o, o->x[i], o->y[i], &wx, &wy);
        dx = dx + (wx / 3600.0) * dt;
        dy = dy + (wy / 3600.0) * dt;
    }
    o->x[i] = o->x[i] + dx;
    o->y[i] = o->y[i] + dy;
    return reward;
}

static void reset_env(Oracle *o, int e, const double *spawn /* [A*5] or NULL */)
{
    int A = o->n_ac;
    if (spawn) {
        for (int a = 0; a < A; ++a) {
            int i = e * A + a;
            o->x[i] = spawn[5 * a]; o->y[i] = spawn[5 * a + 1]; o->h[i] = spawn[5 * a + 2];
            o->phi[i] = spawn[5 * a + 3]; o->v[i] = spawn[5 * a + 4];
        }
    } else {
        spawn_env(o, e);
    }
    o->episodes[e] += 1;
    o->timesteps[e] = 0;          /* atc_gym.py:352-356 */
    o->ep_return[e] = 0.0;
    o->actions_taken[e] = 0;
}

static void write_reset_obs(const Oracle *o, int e, float *obs)
{
    for (int a = 0; a < o->n_ac; ++a) {
        float raw[10];
        ObsAux aux;
        get_state(o, e * o->n_ac + a, 0.0, raw, &aux);       /* atc_gym.py:351 — mva = 0, raw */
        if (o->normalize && o->normalize_reset_obs)
            normalize_obs(o, raw, obs + 10 * (e * o->n_ac + a));
        else
            memcpy(obs + 10 * (e * o->n_ac + a), raw, sizeof raw);
    }
}

static void step_env(Oracle *o, int e, const float *actions, float *obs, float *raw_obs, double *reward,
                     uint8_t *done, int32_t *term, int autoreset)
{
    const int A = o->n_ac;
    int t = ++o->timesteps[e];                                   /* atc_gym.py:135 */
    double base[ORACLE_MAX_AC], mva[ORACLE_MAX_AC];
    int code[ORACLE_MAX_AC];
    int any = 0;
    for (int a = 0; a < A; ++a) {
        int i = e * A + a;
        base[a] = aircraft_advance(o, i, actions + 3 * i);
        code[a] = TERM_RUNNING;
        int m = find_mva(o, o->x[i], o->y[i]);                  /* atc_gym.py:146-161 */
        if (m < 0) {
            base[a] = -50.0; code[a] = TERM_LEFT_AIRSPACE; mva[a] = 0.0;
        } else {
            mva[a] = o->height[m];
            if (o->h[i] < mva[a]) { base[a] = -200.0; code[a] = TERM_BELOW_MVA; }
        }
        if (inside_corridor(o, o->x[i], o->y[i], o->h[i], o->phi[i])) {   /* atc_gym.py:163-169 */
            int bonus = (TIMESTEP_LIMIT - t) * 5;
            base[a] = (double)(10000 + (bonus > 0 ? bonus : 0));
            code[a] = TERM_CAPTURED;
        }
        any |= code[a] != TERM_RUNNING;
    }
    /* env-level overrides: separation (own spec, README.md:51) then timeout (atc_gym.py:171-173) */
    int env_code = TERM_RUNNING, override = 0;
    for (int a = 0; a < A; ++a)
        if (code[a] > env_code) env_code = code[a];
    if (A > 1) {
        int viol = 0;
        for (int a = 0; a < A; ++a)
            for (int b = a + 1; b < A; ++b) {
                double dx = o->x[e * A + a] - o->x[e * A + b], dy = o->y[e * A + a] - o->y[e * A + b];
                double dh = fabs(o->h[e * A + a] - o->h[e * A + b]);
                if (dx * dx + dy * dy < 9.0 && dh < 1000.0) viol = 1;
            }
        if (viol) { env_code = TERM_SEPARATION; override = 1; any = 1; }
    }
    if (t > TIMESTEP_LIMIT) { env_code = TERM_TIMEOUT; override = 1; any = 1; }
    if (override)
        for (int a = 0; a < A; ++a) base[a] = a == 0 ? -200.0 : 0.0;

    double r_ac[ORACLE_MAX_AC];
    int packed = env_code;
    for (int a = 0; a < A; ++a) {
        int i = e * A + a;
        float raw[10];
        ObsAux aux;
        get_state(o, i, mva[a], raw, &aux);                      /* atc_gym.py:175 */
        double r = base[a];
        if (o->shaping) {                                        /* atc_gym.py:179-185, 199-260 */
            double rel_faf = relative_angle(o->phi_to, aux.phi_rel_faf);
            double pos = sigmoid_distance(aux.d_faf, o->dmax) * pow(fabs(rel_faf) / 180.0, 1.5) * 0.8;
            double plane_to_runway = relative_angle(o->phi_to, o->phi[i]);
            double ang_in = sign_(rel_faf) * plane_to_runway;
            double q = (ang_in - 22.5) / 202.0;
            double ang = pow(-(q * q) + 1.0, 32.0) * pos * 1.2;
            double gs = sigmoid_distance(fabs(o->h[i] - aux.on_gp), 36000.0) * pos * 0.8;
            r += pos;
            r += ang;
            r += gs;
        }
        r_ac[a] = r;
        if (raw_obs) memcpy(raw_obs + 10 * i, raw, sizeof raw);
        if (o->normalize)
            normalize_obs(o, raw, obs + 10 * i);
        else
            memcpy(obs + 10 * i, raw, sizeof raw);
        packed |= code[a] << (8 + 3 * a);
    }
    double r_env = tree_sum(r_ac, A);
    reward[e] = r_env;
    done[e] = (uint8_t)any;
    term[e] = packed;
    o->ep_return[e] += r_env;                                    /* atc_gym.py:194-197 */
    if (any) {
        o->last_ep_return[e] = o->ep_return[e];
        o->last_ep_len[e] = t;
        o->win_ring[e] = ((o->win_ring[e] << 1) | (env_code == TERM_CAPTURED)) & 0xFFFF;
        if (autoreset) {
            reset_env(o, e, NULL);
            write_reset_obs(o, e, obs);
        }
    }
}

/* ------------------------------------------------------------------------------------------ C API (ctypes) */

Oracle *atc_oracle_create(int n_mva, const double *ring_xy, const int *ring_off, const double *height,
                          const double *runway /* x, y, h, phi_from */, int n_entry, const double *entry_xyphi,
                          const int *level_off, const int *levels, double dt, int shaping, int normalize,
                          int discrete, int normalize_reset_obs, int n_env, int n_ac, uint64_t seed,
                          int64_t env_index_base, int wind_gx, int wind_gy, const float *wind)
{
    if (n_ac < 1 || n_ac > ORACLE_MAX_AC || n_env < 1) return NULL;
    Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
    int nv = ring_off[n_mva];
    o->n_mva = n_mva;
    o->ring_xy = (double *)malloc(sizeof(double) * 2 * nv);
    memcpy(o->ring_xy, ring_xy, sizeof(double) * 2 * nv);
    o->ring_off = (int *)malloc(sizeof(int) * (n_mva + 1));
    memcpy(o->ring_off, ring_off, sizeof(int) * (n_mva + 1));
    o->height = (double *)malloc(sizeof(double) * n_mva);
    memcpy(o->height, height, sizeof(double) * n_mva);
    o->bounds = (double *)malloc(sizeof(double) * 4 * n_mva);
    o->rwy_x = runway[0]; o->rwy_y = runway[1]; o->rwy_h = runway[2]; o->phi_from = runway[3];
    o->n_entry = n_entry;
    o->entry = (double *)malloc(sizeof(double) * 3 * n_entry);
    memcpy(o->entry, entry_xyphi, sizeof(double) * 3 * n_entry);
    o->level_off = (int *)malloc(sizeof(int) * (n_entry + 1));
    memcpy(o->level_off, level_off, sizeof(int) * (n_entry + 1));
    o->levels = (int *)malloc(sizeof(int) * level_off[n_entry]);
    memcpy(o->levels, levels, sizeof(int) * level_off[n_entry]);
    o->dt = dt; o->shaping = shaping; o->normalize = normalize; o->discrete = discrete;
    o->normalize_reset_obs = normalize_reset_obs;
    derive_constants(o);
    if (wind && wind_gx >= 2 && wind_gy >= 2) {
        o->wind_gx = wind_gx; o->wind_gy = wind_gy;
        o->wind = (double *)malloc(sizeof(double) * 2 * wind_gx * wind_gy);
        for (int i = 0; i < 2 * wind_gx * wind_gy; ++i) o->wind[i] = (double)wind[i];
    }
    o->n_env = n_env; o->n_ac = n_ac; o->seed = seed; o->env_index_base = env_index_base;
    size_t na = (size_t)n_env * n_ac;
    o->x = (double *)calloc(na, 8); o->y = (double *)calloc(na, 8); o->h = (double *)calloc(na, 8);
    o->phi = (double *)calloc(na, 8); o->v = (double *)calloc(na, 8);
    o->last_action = (double *)calloc(na * 3, 8);                /* atc_gym.py:86 */
    o->timesteps = (int32_t *)calloc(n_env, 4); o->actions_taken = (int32_t *)calloc(n_env, 4);
    o->episodes = (int32_t *)calloc(n_env, 4); o->win_ring = (int32_t *)calloc(n_env, 4);
    o->last_ep_len = (int32_t *)calloc(n_env, 4);
    o->ep_return = (double *)calloc(n_env, 8); o->last_ep_return = (double *)calloc(n_env, 8);
    return o;
}

void atc_oracle_destroy(Oracle *o)
{
    if (!o) return;
    free(o->ring_xy); free(o->ring_off); free(o->height); free(o->bounds); free(o->entry); free(o->level_off);
    free(o->levels); free(o->wind); free(o->x); free(o->y); free(o->h); free(o->phi); free(o->v);
    free(o->last_action); free(o->timesteps); free(o->actions_taken); free(o->episodes); free(o->win_ring);
    free(o->last_ep_len); free(o->ep_return); free(o->last_ep_return);
    free(o);
}

/* faf, iaf, corner1, corner2, normal (10) | bbox (4) | dmax, faf_mva, phi_to (3) | nmin (10) | nmax (10) | rings 3x8 */
void atc_oracle_constants(const Oracle *o, double *out)
{
    int k = 0;
    const double *p2[5] = {o->faf, o->iaf, o->corner1, o->corner2, o->normal};
    for (int i = 0; i < 5; ++i) { out[k++] = p2[i][0]; out[k++] = p2[i][1]; }
    for (int i = 0; i < 4; ++i) out[k++] = o->bbox[i];
    out[k++] = o->dmax; out[k++] = o->faf_mva; out[k++] = o->phi_to;
    for (int i = 0; i < 10; ++i) out[k++] = (double)o->nmin[i];
    for (int i = 0; i < 10; ++i) out[k++] = (double)o->nmax[i];
    for (int i = 0; i < 8; ++i) out[k++] = o->tri_h[i];
    for (int i = 0; i < 8; ++i) out[k++] = o->tri_1[i];
    for (int i = 0; i < 8; ++i) out[k++] = o->tri_2[i];
}

/* out[i] = MVA height in ft, or -1 when outside the airspace */
void atc_oracle_mva(const Oracle *o, int n, const double *xy, int32_t *out)
{
    for (int i = 0; i < n; ++i) {
        int m = find_mva(o, xy[2 * i], xy[2 * i + 1]);
        out[i] = m < 0 ? -1 : (int32_t)o->height[m];
    }
}

void atc_oracle_mva_index(const Oracle *o, int n, const double *xy, int32_t *out)
{
    for (int i = 0; i < n; ++i) out[i] = find_mva(o, xy[2 * i], xy[2 * i + 1]);
}

void atc_oracle_inside_corridor(const Oracle *o, int n, const double *xyhphi, uint8_t *out)
{
    for (int i = 0; i < n; ++i)
        out[i] = (uint8_t)inside_corridor(o, xyhphi[4 * i], xyhphi[4 * i + 1], xyhphi[4 * i + 2], xyhphi[4 * i + 3]);
}

void atc_oracle_inside_corridor_angle(const Oracle *o, int n, const double *xyphi, uint8_t *out)
{
    for (int i = 0; i < n; ++i)
        out[i] = (uint8_t)inside_corridor_angle_full(o, xyphi[3 * i], xyphi[3 * i + 1], xyphi[3 * i + 2]);
}

/* mask NULL = all envs.  spawn NULL = device-spec RNG spawn, else [n_env * n_ac * 5] explicit states */
void atc_oracle_reset(Oracle *o, const uint8_t *mask, const double *spawn, float *obs)
{
    for (int e = 0; e < o->n_env; ++e) {
        if (mask && !mask[e]) continue;
        reset_env(o, e, spawn ? spawn + 5 * (size_t)e * o->n_ac : NULL);
        if (obs) write_reset_obs(o, e, obs);
    }
}

void atc_oracle_step(Oracle *o, const float *actions, float *obs, float *raw_obs, double *reward, uint8_t *done,
                     int32_t *term, int autoreset)
{
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o->n_env; ++e)
        step_env(o, e, actions, obs, raw_obs, reward, done, term, autoreset);
}

/* T fused steps, autoreset on; actions [T, n_env, n_ac, 3]; outputs [T, ...].  Used for the CPU baseline timing. */
void atc_oracle_rollout(Oracle *o, int T, const float *actions, float *obs, double *reward, uint8_t *done, int32_t *term)
{
    const size_t na = (size_t)o->n_env * o->n_ac;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o->n_env; ++e)
        for (int t = 0; t < T; ++t)
            step_env(o, e, actions + 3 * na * t, obs + 10 * na * t, NULL, reward + (size_t)o->n_env * t,
                     done + (size_t)o->n_env * t, term + (size_t)o->n_env * t, 1);
}

/* same, also returning info["original_state"] (atc_gym.py:192): the raw observation of the moved aircraft of every
 * step, terminal steps of auto-reset envs included.  raw_obs [T, n_env, n_ac, 10]. */
void atc_oracle_rollout_raw(Oracle *o, int T, const float *actions, float *obs, float *raw_obs, double *reward,
                            uint8_t *done, int32_t *term)
{
    const size_t na = (size_t)o->n_env * o->n_ac;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o->n_env; ++e)
        for (int t = 0; t < T; ++t)
            step_env(o, e, actions + 3 * na * t, obs + 10 * na * t, raw_obs ? raw_obs + 10 * na * t : NULL,
                     reward + (size_t)o->n_env * t, done + (size_t)o->n_env * t, term + (size_t)o->n_env * t, 1);
}

void atc_oracle_get_state(const Oracle *o, double *state /* [n_env*n_ac*5] x,y,h,phi,v */, int32_t *timesteps)
{
    size_t na = (size_t)o->n_env * o->n_ac;
    for (size_t i = 0; i < na; ++i) {
        state[5 * i] = o->x[i]; state[5 * i + 1] = o->y[i]; state[5 * i + 2] = o->h[i];
        state[5 * i + 3] = o->phi[i]; state[5 * i + 4] = o->v[i];
    }
    if (timesteps) memcpy(timesteps, o->timesteps, sizeof(int32_t) * o->n_env);
}

void atc_oracle_set_state(Oracle *o, const double *state, const int32_t *timesteps)
{
    size_t na = (size_t)o->n_env * o->n_ac;
    for (size_t i = 0; i < na; ++i) {
        o->x[i] = state[5 * i]; o->y[i] = state[5 * i + 1]; o->h[i] = state[5 * i + 2];
        o->phi[i] = state[5 * i + 3]; o->v[i] = state[5 * i + 4];
    }
    if (timesteps) memcpy(o->timesteps, timesteps, sizeof(int32_t) * o->n_env);
}

/* metrics: ep_return, last_ep_return [n_env] f64; actions_taken, episodes, win_ring, last_ep_len [n_env] i32 */
void atc_oracle_get_metrics(const Oracle *o, double *ep_return, double *last_ep_return, int32_t *actions_taken,
                            int32_t *episodes, int32_t *win_ring, int32_t *last_ep_len)
{
    size_t n = (size_t)o->n_env;
    if (ep_return) memcpy(ep_return, o->ep_return, 8 * n);
    if (last_ep_return) memcpy(last_ep_return, o->last_ep_return, 8 * n);
    if (actions_taken) memcpy(actions_taken, o->actions_taken, 4 * n);
    if (episodes) memcpy(episodes, o->episodes, 4 * n);
    if (win_ring) memcpy(win_ring, o->win_ring, 4 * n);
    if (last_ep_len) memcpy(last_ep_len, o->last_ep_len, 4 * n);
}

int atc_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void atc_oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
