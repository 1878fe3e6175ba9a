/* flatland_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-environment, single-threaded restatement of the reference hot path
 * (RoboEden/flatland-marl): RailEnv.step + DistanceMap + flatland_cutils TreeObsForRailEnv.
 * It exists to CHECK the CUDA product path; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product library
 * (flatland-marl_b200/csrc) never links, includes or calls anything in this directory.
 *
 * Pinned against golden vectors produced by the unmodified reference
 * (tests/golden/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
 */
#ifndef FLATLAND_ORACLE_H
#define FLATLAND_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_MAX_NODES 31
#define FO_NODE_F 12
#define FO_ATTR_F 83
#define FO_PRED_DEPTH 500
#define FO_ACTION_ABSENT 255

typedef struct FoEnv FoEnv;

/* Builds an environment from an already generated world (what RailEnv.reset leaves behind after the
 * rail/line/timetable generators ran).  pos arrays are [N][2] (row, col). */
FoEnv *fo_create(int H, int W, int N, int T, const uint16_t *grid, const int16_t *init_pos,
                 const uint8_t *init_dir, const int16_t *target, const double *speed,
                 const int32_t *earliest, const int32_t *latest);
void fo_free(FoEnv *e);

/* RailEnv.reset tail (rail_env.py:320-347): distance map, reset_agents, obs_builder.reset(). */
void fo_reset(FoEnv *e);

/* RailEnv.step (rail_env.py:501-634). actions[N] (FO_ACTION_ABSENT = key not in dict),
 * sched[N] = pre-drawn malfunction durations for this step (0 = none).
 * Returns 0, or -1 if the episode was already done (rail_env.py:508-509). */
int fo_step(FoEnv *e, const uint8_t *actions, const uint8_t *sched, int32_t *rewards, uint8_t *dones);

/* TreeObsForRailEnv.get_many + get_properties (treeobs.cpp:30-108,612-640). Any pointer may be NULL. */
int fo_obs(FoEnv *e, float *attr, float *forest, int32_t *adjacency, int32_t *node_order,
           int32_t *edge_order, uint8_t *valid_actions, float *dist_target, uint8_t *deadlocked);

/* MotionCheck on its own (agent_chains.py:19-37,151-236): cur/nxt are [n][2] cell tuples, off-map
 * agents use (-1, i).  Writes can_move[n]. */
void fo_motion_check(int n, const int16_t *cur, const int16_t *nxt, uint8_t *can_move);

/* State read-back ([N] each; pos [N][2], -1 = None). */
void fo_get_state(const FoEnv *e, int16_t *pos, uint8_t *dir, uint8_t *state, uint8_t *ctr, uint8_t *mal,
                  uint16_t *nmal, uint8_t *saved, int32_t *arrival, int16_t *old_pos, int8_t *old_dir,
                  uint8_t *sig_mal);
int fo_elapsed(const FoEnv *e);
int fo_done_all(const FoEnv *e);
int fo_num_targets(const FoEnv *e);
/* Distance map of unique-target slot s as uint16 [H][W][4], 65535 = unreachable; slot of agent i. */
void fo_get_dist_u16(const FoEnv *e, int slot, uint16_t *out);
int fo_target_slot(const FoEnv *e, int agent);

/* Runs n_steps lock-step steps (step + obs) over n_envs environments on n_threads host threads with
 * uniform pseudo-random actions; finished episodes are reset in place.  Used only as the CPU
 * baseline ("port") in bench.py.  Returns agent-steps executed. */
long long fo_bench_run(FoEnv **envs, int n_envs, int n_steps, int n_threads, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
