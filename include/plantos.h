/*
 * plantos.h -- C ABI of libplantos_b200.so, the B200 (sm_100a) batched PlantOS simulator.
 *
 * This is the drop-in boundary for the env-step hot path of GammaKing2000/RL-Env:
 * everything an SB3 `VecEnv.reset()/step_async()/step_wait()` does below the VecEnv
 * call (DummyVecEnv loop -> Monitor -> PlantOSEnv.step/reset, reference
 * A2C_training.py:116-125,216-218 and plantos_env.py:125-372) is one kernel launch
 * behind `plantos_step`.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative PLANTOS_E* code otherwise; the
 *     message for the calling thread's last failure is `plantos_last_error()`;
 *   - "dev" pointers are device pointers on the handle's GPU, owned by the caller
 *     (e.g. torch tensors' data_ptr()); "host" pointers are host memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work
 *     is enqueued on it asynchronously unless stated otherwise;
 *   - one caller thread per handle (SB3 is single-threaded); not re-entrant;
 *   - grid coordinates: x = row = first index, y = column (plantos_env.py:186-190);
 *   - cell codes: 0 empty, 1 obstacle, 2 hydrated plant, 3 thirsty plant
 *     (the reference's LIDAR entity ids, plantos_env.py:20-23).
 */
#ifndef PLANTOS_B200_H
#define PLANTOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLANTOS_ABI_VERSION 2

enum {
    PLANTOS_OK = 0,
    PLANTOS_EINVAL = -1,   /* bad argument / config rejected */
    PLANTOS_ECUDA = -2,    /* CUDA runtime error */
    PLANTOS_ENOMAPS = -3,  /* injected-map mode: an env ran past its pushed maps */
    PLANTOS_ESTATE = -4    /* call not valid in the handle's current state */
};

/* map_source: PHILOX = the reference's cluster generator (plantos_env.py:338-372) with Philox draws;
 * INJECTED = recorded maps (plantos_push_maps); MAZE = the Gradio fork's 'maze' generator
 * (gradio-app/plantos_env_new.py:408-604: randomised DFS over a (G-1)/6 meta grid, 5x5 rooms, 5-wide
 * corridors, random extensions / corner cuts / bulges) with Philox draws, on the device. */
enum { PLANTOS_MAPS_PHILOX = 0, PLANTOS_MAPS_INJECTED = 1, PLANTOS_MAPS_MAZE = 2 };

/* kernel selection (plantos_config_t.kernel) */
enum {
    PLANTOS_KERNEL_AUTO = 0,     /* fast kernel when the preset qualifies, else generic */
    PLANTOS_KERNEL_GENERIC = 1,  /* one warp per env, any supported config */
    PLANTOS_KERNEL_FAST = 2      /* persistent pipelined kernel; G<=28, G+R<=32, R<=7, C<=16 and an instantiated (R,C): presets T, DFLT */
};

/* indices into the reward table (plantos_upload_tables / plantos_compute_tables):
 * each entry is R_STEP + X summed in the reference's order (plantos_env.py:164-169),
 * the *_COMPLETE twins add R_COMPLETE_EXPLORATION afterwards (:179-181). */
enum {
    PLANTOS_RW_NEW = 0,          /* valid move onto a never-visited cell  (:204-205) */
    PLANTOS_RW_REVISIT = 1,      /* valid move onto a visited cell        (:206-207) */
    PLANTOS_RW_INVALID = 2,      /* wall / obstacle                       (:208-211) */
    PLANTOS_RW_WATER_GOAL = 3,   /* watered a thirsty plant               (:217-219) */
    PLANTOS_RW_WATER_EMPTY = 4,  /* watered an empty cell                 (:221-222) */
    PLANTOS_RW_WATER_MISTAKE = 5,/* watered a hydrated plant: documented R_MISTAKE
                                    (README.md:46; the reference raises TypeError here) */
    PLANTOS_RW_COUNT = 6         /* table holds 2*COUNT entries: [k] and [COUNT+k] (+complete) */
};

/* Mirrors PlantOSEnv.__init__ (plantos_env.py:25-27,76-83,120) plus sharding/seed. */
typedef struct plantos_config {
    int32_t struct_size;        /* sizeof(plantos_config_t), ABI check */
    int32_t num_envs;           /* envs simulated by THIS handle / GPU */
    int64_t env_id_base;        /* global id of local env 0 (rank * num_envs when sharded) */
    int32_t grid_size;          /* G, 5..128 */
    int32_t num_plants;         /* P */
    int32_t num_obstacles;      /* O; O/3 clusters are generated (plantos_env.py:341) */
    int32_t lidar_range;        /* R, 1..64 */
    int32_t lidar_channels;     /* C, 1..64 */
    int32_t max_steps;          /* truncation horizon, 1..65535 (reference: 1000) */
    float   thirsty_plant_prob; /* reference: 0.7 */
    int32_t map_source;         /* PLANTOS_MAPS_* */
    uint64_t seed;              /* Philox key */
    double  r_goal, r_mistake, r_invalid, r_water_empty, r_step,
            r_exploration, r_revisit, r_complete_exploration;
    int32_t kernel;             /* PLANTOS_KERNEL_* */
    /* Tuning / test knobs, all 0 by default (they never change results): */
    int32_t tune_fast_grid;     /* > 0: number of thread blocks of the specialised kernels' persistent grid */
    int32_t tune_fast_impl;     /* 0: k_step_tile (lane-per-env tiles); 1: k_step_fast (round 1's table-driven kernel, also
                                   the fallback for caller-supplied LIDAR offsets); 2: k_step_lane when the library was
                                   built with it */
    int32_t tune_no_pdl;        /* != 0: plain launches instead of programmatic dependent launches */
    int32_t tune_l2_keep_mb;    /* > 0: k_step_fast tags state accesses L2::evict_last inside a persisting set-aside of that size */
    int32_t reserved[3];
} plantos_config_t;

typedef struct plantos plantos_t;

/* Fill `cfg` with the reference ctor defaults (plantos_env.py:25-26,76-83,120):
 * G21 P8 O50 R2 C10, prob 0.7, max_steps 1000, DQN reward set, philox maps, seed 0. */
int plantos_default_config(plantos_config_t* cfg);

/* Observation length 5*C + 2 + 25 (plantos_env.py:55-57). */
int plantos_obs_dim(const plantos_config_t* cfg);

/* Host-only: the constant tables the kernels use, evaluated in double precision with
 * the reference's literal expressions, for cross-checking against another host language:
 *   lidar_off  int8  [C][R][2]  (int)(r*cos(2*pi*i/C)), (int)(r*sin(..))  plantos_env.py:261-267
 *   dist_tab   float [R+1]      (float)(r / (double)R)                     :288
 *   pos_tab    float [G]        (float)(x / (double)G)                     :295-296
 *   visit_tab  float [11]       (float)(k / 10.0)                          :308
 *   reward_tab double[2*PLANTOS_RW_COUNT]                                  :164-181
 * Any pointer may be NULL to skip that table. */
int plantos_compute_tables(const plantos_config_t* cfg, int8_t* lidar_off, float* dist_tab,
                           float* pos_tab, float* visit_tab, double* reward_tab);

/* Create a simulator for cfg->num_envs envs on CUDA device `device`.  Validates the
 * config (replaces the reference's ValueError, plantos_env.py:360-364: the border ring is
 * always obstacle-free, so 4G-4 >= P+1 is required), allocates persistent state, uploads
 * the default tables.  Envs hold no valid state until plantos_reset. */
int plantos_create(const plantos_config_t* cfg, int device, plantos_t** out);
int plantos_destroy(plantos_t* h);

/* Replace the device tables (layouts as in plantos_compute_tables; host pointers). Synchronous. */
int plantos_upload_tables(plantos_t* h, const int8_t* lidar_off, const float* dist_tab,
                          const float* pos_tab, const float* visit_tab, const double* reward_tab);

/* Injected-map mode: give every env a queue of `episodes` recorded maps, consumed one per
 * reset in order (replaces the procedural generator, plantos_env.py:338-372).  Host pointers:
 * cells u8 [num_envs][episodes][G*G], rover i16 [num_envs][episodes][2] (x,y).  Resets the
 * per-env map cursor to 0.  Synchronous.  A map may hold at most 255 plant cells (the env record counts thirsty
 * plants in 8 bits; EINVAL otherwise, and plantos_set_state flags the same through plantos_check); the info
 * surface reports total_plants from the config's num_plants, as the reference does (plantos_env.py:317-336), so
 * recorded maps are expected to hold exactly num_plants plants, which every reference map does. */
int plantos_push_maps(plantos_t* h, const uint8_t* cells, const int16_t* rover, int episodes);

/* VecEnv.reset(): start a new episode in every env (PlantOSEnv.reset, plantos_env.py:125-158)
 * and write obs_dev f32 [num_envs][D]. */
int plantos_reset(plantos_t* h, float* obs_dev, void* stream);

/* VecEnv.step(actions): one PlantOSEnv.step (plantos_env.py:160-183) per env, then SB3's
 * auto-reset for finished envs.  All pointers are device pointers:
 *   actions      i64 [N]    0 N, 1 E, 2 S, 3 W, >=4 water (plantos_env.py:166-169)
 *   obs          f32 [N][D] next observation (post-reset where done)
 *   reward       f32 [N]
 *   done         u8  [N]    terminated | truncated
 *   terminated   u8  [N]    may be NULL
 *   truncated    u8  [N]    may be NULL   (SB3 "TimeLimit.truncated" = truncated & !terminated)
 *   terminal_obs f32 [N][D] may be NULL; rows written only where done (SB3 "terminal_observation")
 */
int plantos_step(plantos_t* h, const int64_t* actions, float* obs, float* reward,
                 uint8_t* done, uint8_t* terminated, uint8_t* truncated,
                 float* terminal_obs, void* stream);

/* Same step with HOST buffers (what a numpy VecEnv consumer sees): copies actions H2D,
 * runs the step, copies obs / reward / done back D2H on `stream` and waits for it.
 * Pinned host memory makes the copies asynchronous with respect to the host until the wait. */
int plantos_step_host(plantos_t* h, const int64_t* actions_host, float* obs_host,
                      float* reward_host, uint8_t* done_host, void* stream);

/* Per-env scalar state, struct-of-arrays, i32 [PLANTOS_SC_COUNT][N] (device pointer). */
enum {
    PLANTOS_SC_X = 0, PLANTOS_SC_Y = 1, PLANTOS_SC_STEP_COUNT = 2, PLANTOS_SC_EXPLORED = 3,
    PLANTOS_SC_TOTAL_CELLS = 4, PLANTOS_SC_THIRSTY = 5, PLANTOS_SC_COLLISIONS = 6,
    PLANTOS_SC_COLLIDED = 7, PLANTOS_SC_BONUS_GIVEN = 8, PLANTOS_SC_EPISODE = 9,
    PLANTOS_SC_WATERED = 10, PLANTOS_SC_COUNT = 11
};
/* which: 0 = live state, 1 = snapshot taken at each env's most recent terminal step
 * (the info dict SB3 returns next to terminal_observation). */
int plantos_get_scalars(plantos_t* h, int which, int32_t* out_dev, void* stream);
/* Episode return so far (which=0) / of the last finished episode (which=1), f64 [N]:
 * sum of the python-float rewards in step order, as SB3 Monitor accumulates it. */
int plantos_get_returns(plantos_t* h, int which, double* out_dev, void* stream);

/* Full integer state for parity checks (device pointers, any may be NULL):
 *   cells  u8  [N][G*G] cell codes;  visits i32 [N][G*G] visit_counts (plantos_env.py:146) */
int plantos_get_state(plantos_t* h, uint8_t* cells_dev, int32_t* visits_dev, void* stream);
/* Overwrite state (MCTS-style copy, mcts_custom_trainer.py:218-243): cells/visits as above,
 * scalars i32 [PLANTOS_SC_COUNT][N] (TOTAL_CELLS/THIRSTY are recomputed from cells). */
int plantos_set_state(plantos_t* h, const uint8_t* cells_dev, const int32_t* visits_dev,
                      const int32_t* scalars_dev, void* stream);

/* Episode statistics accumulated on device since create / the last clear, f64 [8]:
 * {episodes, sum return, sum length, sum exploration %, sum collisions, sum plants watered,
 *  n terminated, n truncated}.  `out_dev` is a device pointer (the payload of the optional
 * NCCL all-reduce).  Sums are kept in fixed point (1e-6 units) so they do not depend on
 * the order in which envs finish. */
int plantos_stats(plantos_t* h, double* out_dev, int clear, void* stream);

/* CurriculumWrapper on the device.  mode PLANTOS_CURRICULUM_TERMINATE follows
 * A2C_training.py:37-109 (reaching the exploration threshold terminates the episode; the reference
 * constructs it with 40, 100, 10 and max_episodes_per_maze = 3), PLANTOS_CURRICULUM_MARK follows
 * trainingCode.py:24-98 (the threshold only marks the maze completed; 30, 100, 5, 50).  As in the
 * reference, visit counts persist into the next episode unless the maze was completed or has been
 * played max_episodes_per_maze times, the reset observation shows fresh counts, explored_map
 * restarts every episode, and a new map is drawn at every reset.  Call before plantos_reset; it
 * restarts every env's curriculum state.  Both step kernels implement it;
 * plantos_set_state is refused while a curriculum is active.  mode PLANTOS_CURRICULUM_OFF disables. */
enum { PLANTOS_CURRICULUM_OFF = 0, PLANTOS_CURRICULUM_TERMINATE = 1, PLANTOS_CURRICULUM_MARK = 2 };
int plantos_set_curriculum(plantos_t* h, int mode, double initial_threshold, double max_threshold,
                           double threshold_increment, int max_episodes_per_maze);
/* Current exploration thresholds, f64 [N] device pointer. */
int plantos_get_curriculum_thresholds(plantos_t* h, double* out_dev, void* stream);
/* "Same maze" as the reference's wrapper MEANS it (A2C_training.py:75-86 `reset(seed=self.current_maze_seed)`
 * -- which in the reference never reaches the map generator, so it draws a new map every time): with
 * enable != 0 every reset that keeps the current maze regenerates exactly the map (obstacles, plants,
 * thirsty flags, rover start) of the episode the maze started in.  Default 0 = the reference's actual
 * behaviour.  Needs an active curriculum. */
int plantos_set_curriculum_reuse_map(plantos_t* h, int enable);

/* PlantOSEnv.max_steps is a plain attribute callers may change between steps (plantos_env.py:120);
 * takes effect for every step enqueued afterwards. */
int plantos_set_max_steps(plantos_t* h, int max_steps);

/* Rollout policy of the reference's MCTS planner (mcts_custom_trainer.py:168-216): with probability
 * 0.7 move to the least visited valid neighbour (first minimum in N, E, S, W order), otherwise -- and
 * when every move is blocked -- a uniformly random action.  `uniforms_dev` holds two floats in [0, 1)
 * per env (u[2e] < 0.7 selects the heuristic, floor(5 * u[2e+1]) is the random action), `actions_dev`
 * receives int64 [N] actions ready for plantos_step.  Together with a K-step rollout graph
 * (PlantOSVecEnv.make_rollout) this gives device-side MCTS-style rollouts. */
int plantos_rollout_policy(plantos_t* h, const float* uniforms_dev, int64_t* actions_dev, void* stream);

/* Per-episode log = what SB3's Monitor wrapper records (A2C_training.py:124, trainingCode.py:109;
 * train_improved1/gym/env_0.monitor.csv: "r,l,t" rows).  Once enabled, every env that finishes an
 * episode appends one entry on the device (inside the step kernel, one atomic per warp); the host
 * drains them whenever it likes.  Entries beyond `capacity` between two drains are dropped and
 * counted. */
typedef struct plantos_episode {
    uint32_t env;          /* env index inside this handle (add env_id_base for the global id) */
    uint32_t length;       /* Monitor's l */
    uint32_t step_seq;     /* number of the plantos_step call (0-based) that finished the episode */
    uint32_t flags;        /* bit 0 terminated, bit 1 truncated */
    double episode_return; /* Monitor's r: rewards summed in step order, in double */
    uint16_t collisions;     /* info["total_collisions"] at the end of the episode */
    uint16_t watered;        /* plants hydrated during the episode */
    uint16_t explored_cells; /* info["explored_cells"] and ... */
    uint16_t total_cells;    /* ... info["total_cells"]: exploration_percentage = explored / total * 100, the
                                value the reference's EvaluationCallback tries to log per episode
                                (A2C_training.py:161-179; SB3 Monitor's info_keywords) */
} plantos_episode_t;
/* capacity > 0 enables (or resizes and clears) the log, capacity == 0 disables it. */
int plantos_episode_log_enable(plantos_t* h, int capacity);
/* Copies up to `max_entries` pending entries to host memory `out`, clears the log, reports how many
 * were copied (`n_out`) and how many were lost to overflow since the last drain (`dropped_out`, may
 * be NULL).  Synchronises `stream`. */
int plantos_episode_log_drain(plantos_t* h, plantos_episode_t* out, int max_entries, int* n_out,
                              int64_t* dropped_out, void* stream);

/* Sticky device-side error flag (e.g. PLANTOS_ENOMAPS); synchronises `stream`. */
int plantos_check(plantos_t* h, void* stream);

/* Number of simulator kernels launched by this handle so far. */
int64_t plantos_launch_count(const plantos_t* h);
/* Kernel family the handle selected at create ("generic" / "fast"). */
const char* plantos_kernel_name(const plantos_t* h);
/* num_steps consecutive steps with pre-generated actions in ONE call -- the open-loop rollout the
 * reference's MCTS planner runs on a copied env (mcts_custom_trainer.py:139-166: `_rollout` steps
 * `sim_env` up to max_depth with no learner in between).  actions [num_steps][N]; step k writes
 * obs + k * obs_step_stride (floats; >= N * obs_dim, a multiple of 4), reward / done / terminated /
 * truncated + k * N (the last two may be NULL); terminal_obs as in plantos_step.  Auto-reset applies
 * inside the rollout exactly as in single steps; the result is bit-identical to num_steps plantos_step
 * calls.  On the fast presets this is one launch of the state-resident kernel (each warp keeps its 32
 * envs' window rings and records on the SM for all num_steps steps); otherwise num_steps launches. */
int plantos_rollout(plantos_t* h, int num_steps, const int64_t* actions_dev, float* obs_dev, int64_t obs_step_stride,
                    float* reward_dev, uint8_t* done_dev, uint8_t* terminated_dev, uint8_t* truncated_dev,
                    float* terminal_obs_dev, void* stream);

/* Pipelined stepping for OPEN-LOOP sequences (rollouts with pre-generated actions, benchmarks): with
 * enable != 0 a plantos_step / plantos_rollout that directly follows another one of this handle on the same
 * stream no longer waits for the previous launch as a whole; per-tile counters on the device order the two
 * launches env by env, so the next launch's loads and simulation overlap the previous launch's observation
 * stores and the drain of its last wave.  A launch that follows a plantos_step must write a DIFFERENT obs
 * range (a single step publishes an env before its observation is stored); after a plantos_rollout the same
 * buffers may be written again.  Results are identical.  Contract: `actions` of such a launch must not be
 * produced by work enqueued after the previous launch (anything else enqueued on the stream in between must
 * not touch the launch's inputs or read buffers the next launch overwrites).  Default off; the reference has no
 * counterpart (its DummyVecEnv steps synchronously, A2C_training.py:218). */
int plantos_set_pipelining(plantos_t* h, int enable);
/* The kernel the latest plantos_step actually launched: "k_step_tile", "k_step_fast" (both are the
 * "fast" family: lane-per-env tiles resp. the table-driven half-warp kernel used when a caller uploads
 * LIDAR offsets other than the reference's), "k_step_generic" (also what a fast handle falls back to
 * for an obs pointer that is not 16-byte aligned), "" before the first step. */
const char* plantos_last_step_kernel(const plantos_t* h);
/* Bytes of persistent device state per env. */
int64_t plantos_state_bytes_per_env(const plantos_t* h);

const char* plantos_last_error(void);
int plantos_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PLANTOS_B200_H */
