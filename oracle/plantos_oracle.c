/*
 * TEST INFRASTRUCTURE ONLY -- scalar C restatement of the PlantOS env-step path.
 *
 * Second CPU oracle, for parity runs too large for the Python port
 * (oracle/plantos_oracle.py): the same algorithm, one env at a time, plain arrays, no
 * tricks.  Each function cites the reference lines it follows
 * (/root/reference/plantos_env.py).  Maps are always injected (the reference draws them
 * from Python's global `random`, which C cannot reproduce); auto-reset follows SB3's
 * DummyVecEnv (A2C_training.py:218).  Pinned by tests/test_oracle_golden.py against the
 * trajectories recorded from the unmodified reference and against the Python port.
 *
 * Never linked into or called from the product (rl_env_b200/).
 *   gcc -O2 -shared -fPIC -o oracle/libplantos_oracle.so oracle/plantos_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { EMPTY = 0, OBSTACLE = 1, HYDRATED = 2, THIRSTY = 3 };

typedef struct {
    uint8_t* cells;       /* [G*G] 0 empty 1 obstacle 2 hydrated 3 thirsty */
    int8_t* explored;     /* explored_map, plantos_env.py:234 */
    int32_t* visits;      /* visit_counts, :146 */
    int x, y;
    int step_count, collided, bonus_given, collisions, n_obstacles, n_plants;
    int cursor;           /* next injected map */
    double ep_return;
    int ep_len;
} env_t;

typedef struct {
    int n, G, P, O, R, C, D, max_steps;
    double r_goal, r_mistake, r_invalid, r_water_empty, r_step, r_exploration, r_revisit, r_complete;
    env_t* envs;
    const uint8_t* map_cells;   /* [n][E][G*G], borrowed */
    const int16_t* map_rover;   /* [n][E][2] */
    int episodes;
} oracle_t;

oracle_t* po_create(int n, int G, int P, int O, int R, int C, int max_steps) {
    oracle_t* o = (oracle_t*)calloc(1, sizeof(oracle_t));
    o->n = n; o->G = G; o->P = P; o->O = O; o->R = R; o->C = C; o->max_steps = max_steps;
    o->D = 5 * C + 2 + 25;                                   /* :55-57 */
    o->r_goal = 20; o->r_mistake = -10; o->r_invalid = -5; o->r_water_empty = -5;   /* :76-83 */
    o->r_step = -0.1; o->r_exploration = 10; o->r_revisit = -1; o->r_complete = 50;
    o->envs = (env_t*)calloc((size_t)n, sizeof(env_t));
    for (int i = 0; i < n; ++i) {
        o->envs[i].cells = (uint8_t*)calloc((size_t)G * G, 1);
        o->envs[i].explored = (int8_t*)calloc((size_t)G * G, 1);
        o->envs[i].visits = (int32_t*)calloc((size_t)G * G, 4);
    }
    return o;
}

void po_destroy(oracle_t* o) {
    if (!o) return;
    for (int i = 0; i < o->n; ++i) { free(o->envs[i].cells); free(o->envs[i].explored); free(o->envs[i].visits); }
    free(o->envs);
    free(o);
}

void po_set_rewards(oracle_t* o, const double* r8) {
    o->r_goal = r8[0]; o->r_mistake = r8[1]; o->r_invalid = r8[2]; o->r_water_empty = r8[3];
    o->r_step = r8[4]; o->r_exploration = r8[5]; o->r_revisit = r8[6]; o->r_complete = r8[7];
}

void po_set_maps(oracle_t* o, const uint8_t* cells, const int16_t* rover, int episodes) {
    o->map_cells = cells; o->map_rover = rover; o->episodes = episodes;
    for (int i = 0; i < o->n; ++i) o->envs[i].cursor = 0;
}

/* plantos_env.py:251-315 */
static void observe(const oracle_t* o, const env_t* e, float* obs) {
    const int G = o->G, C = o->C, R = o->R;
    memset(obs, 0, sizeof(float) * (size_t)o->D);
    for (int i = 0; i < C; ++i) {
        const double angle = (2 * M_PI * i) / C;                       /* :261 */
        int dist = R, kind = EMPTY;                                    /* :262-263 */
        for (int r = 1; r <= R; ++r) {
            const int cx = e->x + (int)(r * cos(angle));               /* :266,268 */
            const int cy = e->y + (int)(r * sin(angle));               /* :267,269 */
            if (!(0 <= cx && cx < G && 0 <= cy && cy < G)) { dist = r; kind = OBSTACLE; break; }  /* :271-274 */
            const int c = e->cells[cx * G + cy];
            if (c != EMPTY) { dist = r; kind = c; break; }             /* :277-284 */
        }
        obs[5 * i] = (float)((double)dist / (double)R);                /* :288 */
        obs[5 * i + 1 + kind] = 1.0f;                                  /* :290-292 */
    }
    obs[5 * C] = (float)((double)e->x / (double)G);                    /* :295 */
    obs[5 * C + 1] = (float)((double)e->y / (double)G);                /* :296 */
    for (int lx = 0; lx < 5; ++lx)                                     /* :302-311 */
        for (int ly = 0; ly < 5; ++ly) {
            const int gx = e->x + lx - 2, gy = e->y + ly - 2;
            float v = 1.0f;
            if (0 <= gx && gx < G && 0 <= gy && gy < G) {
                int32_t c = e->visits[gx * G + gy];
                if (c > 10) c = 10;
                v = (float)((double)c / 10.0);
            }
            obs[5 * C + 2 + lx * 5 + ly] = v;
        }
}

/* plantos_env.py:125-158 with an injected map */
static void reset_env(oracle_t* o, int i, float* obs) {
    env_t* e = &o->envs[i];
    const int G = o->G, gg = G * G;
    const int k = e->cursor < o->episodes ? e->cursor : o->episodes - 1;
    e->cursor += 1;
    memcpy(e->cells, o->map_cells + ((size_t)i * o->episodes + k) * gg, (size_t)gg);
    e->x = o->map_rover[((size_t)i * o->episodes + k) * 2];
    e->y = o->map_rover[((size_t)i * o->episodes + k) * 2 + 1];
    e->step_count = 0; e->collided = 0; e->bonus_given = 0; e->collisions = 0;   /* :130-133 */
    e->n_obstacles = 0; e->n_plants = 0;
    for (int c = 0; c < gg; ++c) {
        if (e->cells[c] == OBSTACLE) e->n_obstacles++;
        else if (e->cells[c] != EMPTY) e->n_plants++;
    }
    memset(e->explored, 0, (size_t)gg);
    memset(e->visits, 0, (size_t)gg * 4);
    e->explored[e->x * G + e->y] = 2;                                  /* :236 */
    e->visits[e->x * G + e->y] = 1;                                    /* :147 */
    e->ep_return = 0.0; e->ep_len = 0;
    if (obs) observe(o, e, obs);
}

void po_reset(oracle_t* o, float* obs) {
    for (int i = 0; i < o->n; ++i) reset_env(o, i, obs + (size_t)i * o->D);
}

static int explored_cells(const oracle_t* o, const env_t* e) {         /* :320 */
    int n = 0;
    for (int c = 0; c < o->G * o->G; ++c) n += e->explored[c] > 0;
    return n;
}

/* One VecEnv.step: PlantOSEnv.step (:160-183) per env + DummyVecEnv auto-reset.
 * ep_return/ep_len [n] receive Monitor's r (unrounded) and l where done. */
void po_step(oracle_t* o, const int64_t* actions, float* obs, double* reward, uint8_t* terminated,
             uint8_t* truncated, float* terminal_obs, double* ep_return, int32_t* ep_len,
             int32_t* term_sc /* i32 [11][n], pre-reset scalars where done; may be NULL */) {
    const int G = o->G;
    static const int DX[4] = {-1, 0, 1, 0}, DY[4] = {0, 1, 0, -1};     /* :186 */
    for (int i = 0; i < o->n; ++i) {
        env_t* e = &o->envs[i];
        const int64_t a = actions[i];
        e->step_count += 1;                                            /* :162 */
        double rew = o->r_step;                                        /* :164 */
        if (a < 4) {                                                   /* :166-167, :185-211 */
            const int d = (int)(((a % 4) + 4) % 4);                    /* python list indexing of negatives */
            const int nx = e->x + DX[d], ny = e->y + DY[d];
            if (0 <= nx && nx < G && 0 <= ny && ny < G && e->cells[nx * G + ny] != OBSTACLE) {
                const int fresh = e->visits[nx * G + ny] == 0;         /* :197 */
                e->explored[e->x * G + e->y] = 1;                      /* :198 */
                e->x = nx; e->y = ny;
                e->explored[nx * G + ny] = 2;                          /* :200 */
                e->visits[nx * G + ny] += 1;                           /* :203 */
                rew += fresh ? o->r_exploration : o->r_revisit;
            } else {
                e->collided = 1; e->collisions += 1;                   /* :209-210 */
                rew += o->r_invalid;
            }
        } else {                                                       /* :168-169, :213-222 */
            uint8_t* c = &e->cells[e->x * G + e->y];
            if (*c == THIRSTY) { *c = HYDRATED; rew += o->r_goal; }
            else if (*c == HYDRATED) rew += o->r_mistake;              /* documented result, README.md:46 */
            else rew += o->r_water_empty;
        }
        float* orow = obs + (size_t)i * o->D;
        observe(o, e, orow);                                           /* :173 */
        const int explored = explored_cells(o, e);
        const int total = G * G - e->n_obstacles;                      /* :321 */
        const double pct = ((double)explored / (double)total) * 100;   /* :331 */
        const int term = pct >= 100;                                   /* :176,244-246 */
        const int trunc = e->step_count >= o->max_steps;               /* :177 */
        if (pct >= 100 && !e->bonus_given) { rew += o->r_complete; e->bonus_given = 1; }   /* :179-181 */
        reward[i] = rew;
        terminated[i] = (uint8_t)term;
        truncated[i] = (uint8_t)trunc;
        e->ep_return += rew; e->ep_len += 1;
        if (term || trunc) {
            if (ep_return) ep_return[i] = e->ep_return;
            if (ep_len) ep_len[i] = e->ep_len;
            if (terminal_obs) memcpy(terminal_obs + (size_t)i * o->D, orow, sizeof(float) * (size_t)o->D);
            if (term_sc) {
                const int n = o->n;
                int thirsty = 0;
                for (int c = 0; c < G * G; ++c) thirsty += e->cells[c] == THIRSTY;
                term_sc[0 * n + i] = e->x; term_sc[1 * n + i] = e->y; term_sc[2 * n + i] = e->step_count;
                term_sc[3 * n + i] = explored; term_sc[4 * n + i] = total; term_sc[5 * n + i] = thirsty;
                term_sc[6 * n + i] = e->collisions; term_sc[7 * n + i] = e->collided;
                term_sc[8 * n + i] = e->bonus_given; term_sc[9 * n + i] = e->cursor; term_sc[10 * n + i] = 0;
            }
            if (e->cursor < o->episodes) reset_env(o, i, orow);
            /* else: out of maps -- the caller injects one and calls po_reset_one */
        }
    }
}

/* install a new map for env i (used when maps arrive one episode at a time) */
void po_reset_one(oracle_t* o, int i, const uint8_t* cells, int rx, int ry, float* obs_row) {
    env_t* e = &o->envs[i];
    const int G = o->G, gg = G * G;
    memcpy(e->cells, cells, (size_t)gg);
    e->x = rx; e->y = ry;
    e->step_count = 0; e->collided = 0; e->bonus_given = 0; e->collisions = 0;
    e->n_obstacles = 0; e->n_plants = 0;
    for (int c = 0; c < gg; ++c) {
        if (e->cells[c] == OBSTACLE) e->n_obstacles++;
        else if (e->cells[c] != EMPTY) e->n_plants++;
    }
    memset(e->explored, 0, (size_t)gg);
    memset(e->visits, 0, (size_t)gg * 4);
    e->explored[rx * G + ry] = 2;
    e->visits[rx * G + ry] = 1;
    e->ep_return = 0.0; e->ep_len = 0;
    if (obs_row) observe(o, e, obs_row);
}

/* cells u8 [n][G*G], visits i32 [n][G*G], scalars i32 [11][n] in PLANTOS_SC_* order */
void po_get_state(const oracle_t* o, uint8_t* cells, int32_t* visits, int32_t* sc) {
    const int gg = o->G * o->G, n = o->n;
    for (int i = 0; i < n; ++i) {
        const env_t* e = &o->envs[i];
        if (cells) memcpy(cells + (size_t)i * gg, e->cells, (size_t)gg);
        if (visits) memcpy(visits + (size_t)i * gg, e->visits, (size_t)gg * 4);
        if (sc) {
            int thirsty = 0;
            for (int c = 0; c < gg; ++c) thirsty += e->cells[c] == THIRSTY;
            sc[0 * n + i] = e->x; sc[1 * n + i] = e->y; sc[2 * n + i] = e->step_count;
            sc[3 * n + i] = explored_cells(o, e); sc[4 * n + i] = gg - e->n_obstacles;
            sc[5 * n + i] = thirsty; sc[6 * n + i] = e->collisions; sc[7 * n + i] = e->collided;
            sc[8 * n + i] = e->bonus_given; sc[9 * n + i] = e->cursor; sc[10 * n + i] = 0;
        }
    }
}

/* int8 [C][R][2], the offsets this file's observe() uses (for table cross-checks) */
void po_lidar_offsets(int C, int R, int8_t* out) {
    for (int i = 0; i < C; ++i) {
        const double angle = (2 * M_PI * i) / C;
        for (int r = 1; r <= R; ++r) {
            out[(i * R + r - 1) * 2] = (int8_t)(int)(r * cos(angle));
            out[(i * R + r - 1) * 2 + 1] = (int8_t)(int)(r * sin(angle));
        }
    }
}
