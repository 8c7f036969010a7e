/*
 * ss_env.h -- C-ABI of the B200 batched stepping-stone environment (libss_b200.so).
 *
 * The reference has no native FFI: its operator boundary for this path is the Python vectorised-env
 * protocol of common/envs_utils.py (ShmemVecEnv, :486-620) over N copies of
 * mocca_envs:Walker3DStepperEnv-v0 / MikeStepperEnv-v0.  Each entry point below names the reference
 * interface it replaces; steppingstone_b200/vec_env.py binds them with ctypes and INTEGRATION.md shows
 * the reference-side stub.
 *
 * Conventions: every function returns 0 on success and a negative code on error (ss_last_error()
 * gives the text, thread-local).  Pointers named *_dev are device pointers owned by the caller
 * (torch CUDA tensors); *_host are host pointers.  `stream` is a cudaStream_t passed as void*
 * (torch.cuda.current_stream().cuda_stream); all device work is enqueued on it and nothing
 * synchronises unless stated.  A handle belongs to one device and is not thread-safe (like the
 * reference: strictly call-and-wait, envs_utils.py:550-558).
 */
#ifndef SS_ENV_H
#define SS_ENV_H
#include <stdint.h>
#include <stddef.h>
#include "ss_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SS_OK 0
#define SS_ERR_ARG (-1)
#define SS_ERR_CUDA (-2)
#define SS_ERR_STATE (-3)

#define SS_NWARM 16      /* warm-start cache: ids of the last substep's <= 8 contacts (-1 = none), then their normal impulses */
#define SS_NSCAL 16      /* doubles per env in ss_state_view.scalars                              */
#define SS_NINT 8        /* int32 per env in ss_state_view.ints                                   */
#define SS_NEV 18        /* int32 per env and substep in the discrete-event trace (ss_substep_events)   */
#define SS_MT_WORDS 625  /* MT19937: 624 state words + position                                   */
#define SS_TABLE_RING 8  /* sampling-table versions kept alive (ss_set_sample_cdf)                  */

/* scalars[]: */
enum { SS_S_POTENTIAL = 0, SS_S_WALK_X, SS_S_WALK_Y, SS_S_WALK_Z, SS_S_INITIAL_Z,
       SS_S_FEET_CONTACT0, SS_S_FEET_CONTACT1, SS_S_FEET_X0 /* ..+5: feet_xyz[2][3] */,
       SS_S_EP_RET = 13, SS_S_CACHE_CONTACT0 = 14, SS_S_CACHE_CONTACT1 = 15 };
/* ints[]: */
enum { SS_I_NEXT_STEP = 0, SS_I_REACHED_COUNT, SS_I_MIRRORED, SS_I_ELAPSED, SS_I_EP_LEN,
       SS_I_UPDATE_TERRAIN, SS_I_PHANTOM_IDX, SS_I_STARTED };

/* Full simulator state of N envs in host memory (teacher forcing, checkpoints, parity tests).
 * NULL members are skipped.  Reference analogue: none (env state is never saved). */
typedef struct ss_state_view {
    double* q;         /* [N,28] base xyz, base quat (x,y,z,w), 21 hinge angles                   */
    double* qd;        /* [N,27] base angular vel (world), base linear vel (world), 21 hinge rates */
    double* warm;      /* [N,16] contact ids [0..8) + cached normal impulses [8..16) (manifold warm starting) */
    double* terrain;   /* [N,n_steps,6] terrain_info: x,y,z,phi,x_tilt,y_tilt                     */
    double* plank;     /* [N,rendered,6] pose each plank body was last given                      */
    double* scalars;   /* [N,SS_NSCAL]                                                            */
    int32_t* ints;     /* [N,SS_NINT]                                                             */
    uint32_t* mt;      /* [N,2,SS_MT_WORDS] env np_random stream, then the env's aux stream (see ss_create)      */
    int32_t* table_slot; /* [N] which of the SS_TABLE_RING sampling-table versions each env draws from (ss_set_sample_cdf) */
} ss_state_view;

typedef struct ss_env ss_env;

/* make_vec_envs(env_id, seed, num_processes, log_dir) -- common/envs_utils.py:48-56, one gym env per
 * worker seeded seed+rank (:25-40).  mt_env_host [n_envs,625] are the MT19937 states of gym's per-env
 * RandomState(seed+rank) (terrain generator + curriculum sampler draws).  mt_aux_host [n_envs,625] is the
 * second stream each env owns: for Walker3D the forked copy of the GLOBAL numpy stream its terrain generator
 * also draws from (walker3d_envs.py:728); for Mike the robot's own RandomState, which the
 * reference never re-seeds (env_base.py:26,65 vs :131-133) and which decides the left/right mirrored start
 * (robots.py:846).  The Python host derives both (steppingstone_b200/seeding.py).  Multi-GPU shards pass the
 * streams of their own global env indices. */
int ss_create(const ss_model* model, const ss_params* params, int32_t n_envs, int32_t device,
              const uint32_t* mt_env_host, const uint32_t* mt_aux_host, ss_env** out);
/* VecEnv.close -- envs_utils.py:273-281 */
void ss_destroy(ss_env* env);

/* ShmemVecEnv.reset -- envs_utils.py:542-548 -> Walker3DStepperEnv.reset walker3d_envs.py:812-867 /
 * MikeStepperEnv.reset mike.py:333-388.  obs_dev [N,60] f32. */
int ss_reset_all(ss_env* env, float* obs_dev, void* stream);

/* ShmemVecEnv.step_async+step_wait -- envs_utils.py:550-558; worker loop with auto-reset :645-649;
 * Walker3DStepperEnv.step walker3d_envs.py:875-903; Monitor.update envs_utils.py:130-153;
 * TimeLimitMask.step :60-65.  act_dev [N,21] f32; obs_dev [N,60] f32 (post-reset obs for finished
 * envs); rew_dev [N] f32; done_dev [N] u8; timeout_dev [N] u8 (info["bad_transition"]);
 * ep_ret_dev [N] f32 / ep_len_dev [N] i32 = info["episode"]["r"/"l"], valid where done. */
int ss_step(ss_env* env, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev,
            uint8_t* timeout_dev, float* ep_ret_dev, int32_t* ep_len_dev, void* stream);

/* Same call with HOST buffers (the reference's numpy-in / numpy-out contract, envs_utils.py:550-558, 608-620):
 * rew_host is float64 like the reference's stacked python floats, done_host / timeout_host are 0/1 bytes (numpy
 * bool storage).  One H2D copy of the actions, ONE launch over the batch; the kernel writes every output straight into
 * pinned host memory -- into the caller's arrays themselves where they are pinned (ss_host_alloc / cudaHostAlloc /
 * cudaHostRegister; checked per array), else into a pinned buffer of the library's that is copied out.
 * Returns after everything has landed in the caller's arrays.  A non-finite action fails with SS_ERR_ARG before anything
 * runs (robots.py:290: fatal in the reference too, where the other workers have stepped by then). */
int ss_step_host(ss_env* env, const float* act_host, float* obs_host, double* rew_host,
                 uint8_t* done_host, uint8_t* timeout_host, float* ep_ret_host, int32_t* ep_len_host);
int ss_reset_host(ss_env* env, float* obs_host);
/* pinned host memory for observation arrays the kernel can write into directly (no reference analogue: ShmemVecEnv's
 * shared-memory observation buffers, envs_utils.py:507-513, play this role there) */
int ss_host_alloc(size_t bytes, void** out);
void ss_host_free(void* p);

/* update_curriculum / update_specialist / update_sample_prob -- walker3d_envs.py:1346-1356, :604-622,
 * :1317-1320 (mike.py:856-866, :827-830) through envs_utils.py:568-586.  The caller passes the
 * normalised cumulative table exactly as RandomState.choice builds it (cdf = cumsum(p/p.sum());
 * cdf /= cdf[-1]); n_cells must be 11^5.  Every call installs a new table VERSION (a ring of SS_TABLE_RING versions
 * lives on the device, each env holds the index of the one it draws from):
 *   apply_now == 0  update_curriculum / update_specialist: every env takes the new table (the reference calls the method
 *                   on every env unconditionally, envs_utils.py:580-586);
 *   apply_now != 0  update_sample_prob (mike.py:827-830): ONLY the envs whose update_terrain flag is set -- they advanced
 *                   a stone in their last step -- take it and re-sample their next-next stone at once; every other env
 *                   keeps drawing from the version it has, exactly like a reference env ignores the table it is sent.
 * When the ring wraps, envs still on the version being overwritten move to the oldest surviving one (the only deviation:
 * an env that has not advanced a stone across SS_TABLE_RING consecutive table updates). */
/* current head of the version ring and how many versions were ever installed (checkpoints) */
int ss_table_ring_state(const ss_env* env, int32_t* head, int64_t* count);
int ss_set_sample_cdf(ss_env* env, const double* cdf_host, int32_t n_cells, int32_t apply_now, void* stream);

/* set_env_params / set_robot_params -- env_base.py:73-85 through envs_utils.py:560-566: the reference setattr()s
 * fields of every worker's env / robot (and recomputes the torque limits, robots.py:385-388).  Here the host
 * rebuilds the two constant tables and this call swaps them in (synchronises `stream`; takes effect from the next
 * step / reset).  n_steps and rendered cannot change (they size the per-env terrain arrays). */
int ss_update_tables(ss_env* env, const ss_model* model, const ss_params* params, void* stream);

/* create_temp_states / create_temp_states_1 / create_temp_states_2 -- walker3d_envs.py:1189-1243, mike.py:699-753 through
 * envs_utils.py:573-578 (value-guided sampling probes read by train.py:247-271).  variant 1: yaw x pitch at r = 0.7
 * (121 cells); 2: yaw x pitch x r (1331); 5: yaw x pitch x r x x_tilt x y_tilt (161051).  grid_host [5][11] = the
 * env's fake_{yaw,pitch,r,x_tilt,y_tilt}_samples.  For every env whose update_terrain flag is set, row g of
 * out_dev [N, cells, 60] f32 is get_temp_state() with the next-next stone moved to grid cell g; other envs get zeros
 * (self.temp_states).  Reference side effects kept: walk_target stays at the last cell's stone; the stone is then
 * re-placed from restore_host = (pitch, yaw, dr, x_tilt, y_tilt) (the env's next_next_*; NULL = leave the row as the
 * sampler wrote it, which is what Mike's restore amounts to).  Stream-ordered, no synchronisation. */
int64_t ss_temp_state_cells(int32_t variant);
int ss_create_temp_states(ss_env* env, int32_t variant, const double* grid_host, const double* restore_host,
                          float* out_dev, void* stream);

/* State access (synchronous). */
int ss_get_state(ss_env* env, ss_state_view* view);
int ss_set_state(ss_env* env, const ss_state_view* view);

/* Per-substep parity probe (SURVEY 8b ss_substep_trace): after each ss_step the library keeps the
 * generalised coordinates reached after each of the 4 physics substeps.  q_host [4,N,28],
 * qd_host [4,N,27] doubles.  Must be enabled first (costs 1.76 KB/env of extra stores per step). */
int ss_enable_substep_trace(ss_env* env, int32_t on);
int ss_substep_trace(ss_env* env, double* q_host, double* qd_host);
/* The DISCRETE decisions each of those substeps took: ev_host [4,N,SS_NEV] int32 = bit j set: hinge j got a lower-limit row
 * (btMultiBodyJointLimitConstraint: q - lo <= 0), same for the upper limit (bit 30 of that word: the max_coord_vel safety
 * clamp engaged in this substep), then the <= 8 contact points that entered the solver (-1 = none): contact id in the low
 * 16 bits, above it the branch of the contact geometry that produced the point (bit 16 btPlaneSpace1's |n.z| > 0.7071
 * friction basis, bit 17 capsule segment off the plank's footprint, bit 18 point inside the box, bits 19-20 its face axis).
 * Words 10..17: per contact, bit `it` set iff its normal impulse was > 0 when solver iteration `it` refreshed the friction
 * bounds.  Bullet skips a friction row while its normal impulse is zero (`if (totalImpulse > 0)` in
 * btMultiBodyConstraintSolver::solveSingleIteration), so the row keeps whatever impulse it had: the iteration at which a
 * normal lets go is a discrete event too, with a first-order effect on light links.  A limit row or a contact point exists or not, so fp32 rounding can flip such a decision
 * where the fp64 reference sits within rounding of the threshold; the parity tests compare trajectories conditional on
 * equal decisions and account for every flip.  Needs ss_enable_substep_trace. */
int ss_substep_events(ss_env* env, int32_t* ev_host);

/* Stand-alone articulated-body forward-dynamics substep (no contacts, no limits): the
 * `aba_substep` kernel boundary of SURVEY 8(d) -- reads q [N,28], qd [N,27], tau [N,21] (f32, device),
 * writes q, qd in place; 524 algorithmic bytes per env-substep. */
int ss_aba_substep(ss_env* env, float* q_dev, float* qd_dev, const float* tau_dev, int32_t n, void* stream);

/* Device-side GAE scan: RolloutStorage.compute_returns(use_gae=True) -- algorithms/storage.py:59-71.
 * rewards [T,N], value_preds [T+1,N], masks [T+1,N], bad_masks [T+1,N], returns [T+1,N] (f32, device). */
int ss_gae(const float* rewards_dev, const float* value_preds_dev, const float* masks_dev,
           const float* bad_masks_dev, float* returns_dev, int32_t T, int32_t N, float gamma,
           float gae_lambda, void* stream);

/* Sticky error flags of the device entry points (synchronises).  Bit 0 (SS_FLAG_BAD_ACTION): a non-finite action reached
 * ss_step.  The reference asserts inside every worker's step() (robots.py:290); ss_step_host does the same check on the host
 * and refuses the batch, but ss_step is stream-ordered and cannot raise: the env steps such an action as all-zero (state
 * and reward stay finite) and sets the flag -- check it once per rollout (steppingstone_b200.collector does). */
#define SS_FLAG_BAD_ACTION 1
int ss_error_flags(ss_env* env, int32_t* flags_out, int32_t clear);

/* Golden-tape probe: one step() with the PHYSICS REPLACED by the state the caller has just injected with ss_set_state
 * (q, qd and the two foot-contact flags scalars[SS_S_CACHE_CONTACT0..1]) -- everything after pybullet.stepSimulation runs:
 * calc_state, calc_env_state, rewards, TimeLimit / Monitor bookkeeping, auto-reset.  The reference-generated fixtures
 * (tests/golden/ *.npz, tools/gen_golden.py) replay through this entry point on the GPU
 * (tests/test_gpu_parity.py::test_gpu_replays_reference_goldens).  Same buffers as ss_step. */
int ss_step_injected(ss_env* env, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev,
                     uint8_t* timeout_dev, float* ep_ret_dev, int32_t* ep_len_dev, void* stream);

/* Parity probes (tests only; no reference analogue).
 * ss_env_post: calc_state + calc_env_state on the CURRENT state without stepping the physics
 *   (robots.py:304-365 + walker3d_envs.py:1026-1044), no TimeLimit/Monitor/auto-reset.
 * ss_enable_debug / ss_debug_read: per-env dump of the last substep's joint-space inertia H [27,27], bias C [27],
 *   contact candidates [32,8], row constants and Delassus matrix; out_host [N, ss_debug_floats()] f32. */
int ss_env_post(ss_env* env, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, void* stream);
int32_t ss_debug_floats(void);
int ss_enable_debug(ss_env* env, int32_t on);
int ss_debug_read(ss_env* env, float* out_host);

/* host-side breakdown of the last ss_step_host (seconds): out4 = action copy, enqueue, waiting on the
 * device, copying out whatever the caller did not give pinned memory for (bench.py reports it next to e2e) */
int ss_host_timing(const ss_env* env, double* out4);
int32_t ss_num_envs(const ss_env* env);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t ss_launch_count(const ss_env* env);
const char* ss_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
