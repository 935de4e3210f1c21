/*
 * dial_b200.h — C ABI of the B200-native DIAL-MPC sampling core.
 *
 * The reference (LeCAR-Lab/dial-mpc) has no FFI for this path: its boundary is the
 * Python class `MBDPI` plus the env registry, and every function below replaces a
 * piece of jitted JAX that `MBDPI` calls.  Each entry point cites the reference
 * code it stands in for (paths relative to the reference root).
 *
 * Conventions
 *   - extern "C", opaque handles, plain pointers and sizes, no torch / C++ types.
 *   - every `const float* / float*` argument marked [dev] is a CUDA device pointer to
 *     contiguous row-major fp32 owned by the caller; `stream` is a `cudaStream_t`
 *     passed as `void*`; all work is enqueued asynchronously on it.
 *   - return 0 on success, negative on error; `dial_last_error()` returns a static,
 *     thread-local, NUL-terminated description of the last failure.
 *   - one plan per GPU; a plan is not thread-safe.
 *   - no hidden allocations after `dial_plan_create`.
 */
#ifndef DIAL_B200_H_
#define DIAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIAL_ABI_VERSION 8

/* capacities of the fixed-size device model */
#define DIAL_MAXB 24   /* bodies incl. world            */
#define DIAL_MAXV 28   /* dofs                          */
#define DIAL_MAXQ 29   /* generalized positions         */
#define DIAL_MAXU 20   /* actuators                     */
#define DIAL_MAXG 8    /* collision geoms               */
#define DIAL_MAXP 16   /* contact pairs                 */
#define DIAL_MAXC 20   /* contacts                      */
#define DIAL_MAXS 8    /* sites                         */
#define DIAL_MAXNODE 8 /* Hnode+1                       */
#define DIAL_MAXH 64   /* Hsample+1                     */
#define DIAL_MAXSTAGE 12
#define DIAL_MAXUSER 64 /* user constants of a custom reward */
#define DIAL_MAXRANK 8  /* GPUs of one NVLink domain sharing the samples */
#define DIAL_IPC_HANDLE_BYTES 64

/* environments (reward functors fused into the rollout kernel) */
enum {
  DIAL_ENV_GO2_WALK = 0,    /* UnitreeGo2Env.step         envs/unitree_go2_env.py:126-261 */
  DIAL_ENV_GO2_SEQJUMP = 1, /* UnitreeGo2SeqJumpEnv.step  envs/unitree_go2_env.py:403-521 */
  DIAL_ENV_H1_WALK = 2,     /* UnitreeH1WalkEnv.step      envs/unitree_h1_env.py:181-321  */
  DIAL_ENV_ALLEGRO = 3,     /* AllegroReorientEnv.step    envs/manipulation.py:63-100     */
  DIAL_ENV_H1_LOCO = 4,     /* UnitreeH1LocoEnv.step      envs/unitree_h1_env.py:686-830  */
  DIAL_ENV_CUSTOM = 5,      /* user reward (README.md:223-312 "Writing Custom Environment"):
                               a device functor compiled into a dedicated build of this library,
                               see include/dial_custom_reward.h                              */
};

/* Compiled robot model: what `brax.io.mjcf.load` + `mjx.put_model` give the reference
 * (envs/unitree_go2_env.py:95-99).  Filled by the host-side model compiler. */
typedef struct dial_model_desc {
  int32_t nq, nv, nu, nbody, njnt, ngeom, nsite, npair, ncon;
  int32_t iterations, ls_iterations, eulerdamp, cone;
  float timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  /* bodies (index 0 = world) */
  int32_t body_parentid[DIAL_MAXB], body_rootid[DIAL_MAXB], body_depth[DIAL_MAXB];
  int32_t body_jntadr[DIAL_MAXB], body_dofadr[DIAL_MAXB], body_dofnum[DIAL_MAXB];
  float body_pos[DIAL_MAXB][3], body_quat[DIAL_MAXB][4];
  float body_ipos[DIAL_MAXB][3], body_iquat[DIAL_MAXB][4];
  float body_mass[DIAL_MAXB], body_inertia[DIAL_MAXB][3], body_invweight0[DIAL_MAXB];
  float body_invweight0_rot[DIAL_MAXB];
  /* joints (at most one per body) */
  int32_t jnt_type[DIAL_MAXB], jnt_qposadr[DIAL_MAXB], jnt_dofadr[DIAL_MAXB], jnt_limited[DIAL_MAXB];
  float jnt_pos[DIAL_MAXB][3], jnt_axis[DIAL_MAXB][3], jnt_range[DIAL_MAXB][2], jnt_margin[DIAL_MAXB];
  float jnt_solref[DIAL_MAXB][2], jnt_solimp[DIAL_MAXB][5];
  /* dofs */
  int32_t dof_bodyid[DIAL_MAXV], dof_jntid[DIAL_MAXV], dof_parentid[DIAL_MAXV];
  float dof_armature[DIAL_MAXV], dof_damping[DIAL_MAXV], dof_invweight0[DIAL_MAXV];
  float qpos0[DIAL_MAXQ];
  /* collision geoms + static contact pairs (MJX fixed-size contact arrays) */
  int32_t geom_type[DIAL_MAXG], geom_bodyid[DIAL_MAXG];
  float geom_pos[DIAL_MAXG][3], geom_quat[DIAL_MAXG][4], geom_size[DIAL_MAXG][3];
  int32_t pair_kind[DIAL_MAXP], pair_geom1[DIAL_MAXP], pair_geom2[DIAL_MAXP], pair_ncon[DIAL_MAXP];
  int32_t pair_condim[DIAL_MAXP];
  float pair_friction[DIAL_MAXP][5], pair_margin[DIAL_MAXP], pair_gap[DIAL_MAXP];
  float pair_solref[DIAL_MAXP][2], pair_solimp[DIAL_MAXP][5];
  /* sites */
  int32_t site_bodyid[DIAL_MAXS];
  float site_pos[DIAL_MAXS][3];
  /* actuators (joint transmissions): force = gain*ctrl + b0 + b1*q + b2*qd */
  int32_t actuator_dofadr[DIAL_MAXU], actuator_qposadr[DIAL_MAXU];
  int32_t actuator_ctrllimited[DIAL_MAXU], actuator_forcelimited[DIAL_MAXU];
  float actuator_gear[DIAL_MAXU], actuator_gain[DIAL_MAXU], actuator_bias[DIAL_MAXU][3];
  float actuator_ctrlrange[DIAL_MAXU][2], actuator_forcerange[DIAL_MAXU][2];
} dial_model_desc;

/* Environment + planner configuration: DialConfig (core/dial_config.py:4-23),
 * BaseEnvConfig (config/base_env_config.py:4-20) and the env-specific constants. */
typedef struct dial_plan_desc {
  int32_t env_id;
  int32_t Nsample;   /* samples rolled by THIS rank (shard size)                  */
  int32_t Ntotal;    /* Nsample of the whole job (== Nsample when not sharded)    */
  int32_t shard_offset; /* global index of this rank's first sample               */
  int32_t Hsample, Hnode;
  int32_t n_frames;  /* int(dt / timestep), base_env.py:17                        */
  int32_t leg_control_torque; /* 1: act2tau, 0: act2joint -> position actuator    */
  float temp_sample;
  float dt, action_scale;
  float kp[DIAL_MAXU], kd[DIAL_MAXU];
  float joint_range[DIAL_MAXU][2];          /* env.joint_range (sampling range)   */
  float physical_joint_range[DIAL_MAXU][2]; /* sys.jnt_range[1:]                  */
  float joint_torque_range[DIAL_MAXU][2];   /* sys.actuator_ctrlrange (+-inf ok)  */
  float joint_offset[DIAL_MAXU];            /* added to the joint target (Allegro: init_q[7:], manipulation.py:106) */
  float M_n2u[DIAL_MAXH][DIAL_MAXNODE];     /* node -> action spline matrix       */
  /* reward constants */
  int32_t torso_body;       /* MuJoCo body id of the torso (x index + 1)          */
  int32_t nfeet;
  int32_t feet_site[4];
  float gait_duty, gait_cadence, gait_amplitude, gait_phase[4];
  float vel_cmd[3], ang_cmd[3], ramp_up_time, pos_tar[3];
  /* seq-jump */
  int32_t n_stage;
  float jump_dt;
  float pose_seq[DIAL_MAXSTAGE][3], yaw_seq[DIAL_MAXSTAGE];
  float contact_targets[DIAL_MAXSTAGE][4][3], contact_radius[DIAL_MAXSTAGE][4];
  /* DIAL_ENV_CUSTOM: constants handed to dial_custom_reward() (ctx->user) */
  int32_t n_user;
  float user[DIAL_MAXUSER];
  /* randomize_tasks (unitree_go2_env.py:141-155, unitree_h1_env.py:198-212): the env step whose
   * info["step"] equals cmd_step uses (cmd_vel, cmd_ang) instead of (vel_cmd, ang_cmd); -1 = none.
   * Set through dial_plan_set_command only (dial_plan_create ignores the value and starts at -1). */
  int32_t cmd_step;
  float cmd_vel[3], cmd_ang[3];
} dial_plan_desc;

/* State handed to the planner: Brax `State.pipeline_state` (qpos, qvel,
 * qacc_warmstart) + `State.info` counters the rewards read
 * (envs/unitree_go2_env.py:106-118, :366-381). */
typedef struct dial_state {
  const float* qpos;           /* [dev] [nq] */
  const float* qvel;           /* [dev] [nv] */
  const float* qacc_warmstart; /* [dev] [nv] */
  int32_t step;                /* info["step"]          */
  int32_t stage;               /* info["contact_stage"] */
} dial_state;

typedef struct dial_plan dial_plan;

int dial_abi_version(void);
const char* dial_last_error(void);
/* sizeof() of the descriptor structs as compiled into the library: which = 0 model, 1 plan,
 * 2 state, 3 mpc buffers (lets foreign-language bindings verify their struct layout). */
size_t dial_sizeof(int which);

/* Create / destroy a plan (uploads model + config, allocates all workspaces). */
dial_plan* dial_plan_create(const dial_model_desc* model, const dial_plan_desc* cfg);
void dial_plan_destroy(dial_plan* plan);

/* rollout_us_vmap — core/dial_core.py:36-42,80-81: roll B action sequences from one
 * state.  us [dev] [B, Hsample+1, nu]; outputs (nullable except rewss):
 * rewss [B,Hs+1], q [B,Hs+1,nq], qd [B,Hs+1,nv], xpos [B,Hs+1,nbody-1,3]. */
int dial_rollout(dial_plan* plan, const dial_state* s, const float* us, int B, int H,
                 float* rewss, float* q, float* qd, float* xpos, void* stream);

/* env.step for ONE instance — the `step_env(state, Y0[0])` call at
 * core/dial_core.py:245.  Writes the successor state (qpos,qvel,qacc_warmstart
 * [dev]), the reward [dev][1] and ctrl [dev][nu]; step/stage are advanced by the host. */
int dial_env_step(dial_plan* plan, const dial_state* s, const float* action,
                  float* qpos_out, float* qvel_out, float* warm_out, float* reward,
                  float* ctrl_out, void* stream);

/* randomize_tasks support: replaces the command of ONE env step (the one whose info["step"]
 * == cmd_step; pass -1 for none) in every later launch of this plan — the reference draws a
 * one-step random command whenever step % 500 == 0 (unitree_go2_env.py:141-163), from a key chain
 * that depends only on the reset key, so the host can compute it ahead of the horizon reaching it.
 * Stream-ordered (safe between replays of the control-step graph). */
int dial_plan_set_command(dial_plan* plan, int cmd_step, const float vel[3], const float ang[3], void* stream);

/* randomize_tasks of UnitreeGo2SeqJumpEnv (unitree_go2_env.py:383-394, 594-631): `reset` draws a
 * whole jump sequence (11 stages) instead of using the configured one.  Replaces the stage tables
 * of the plan (n_stage <= DIAL_MAXSTAGE; pose [n][3], yaw [n], contact_targets [n][4][3],
 * contact_radius [n][4], host pointers) for every later launch.  Stream-ordered like
 * dial_plan_set_command. */
int dial_plan_set_stages(dial_plan* plan, int n_stage, const float* pose_seq, const float* yaw_seq,
                         const float* contact_targets, const float* contact_radius, void* stream);

/* The same step, also returning what the envs' `_get_obs` (envs/unitree_go2_env.py:263-286,
 * unitree_h1_env.py:323-346) reads of pipeline_state.x / xd: kin_out [dev][13] = x.pos(3),
 * x.rot(4) of the torso body, then global_to_body_velocity(xd.vel) (3) and
 * global_to_body_velocity(xd.ang * pi/180) (3) (nullable). */
int dial_env_step_kin(dial_plan* plan, const dial_state* s, const float* action,
                      float* qpos_out, float* qvel_out, float* warm_out, float* reward,
                      float* ctrl_out, float* kin_out, void* stream);

/* pipeline_init — envs/unitree_go2_env.py:104: mjx.forward at (qpos, qvel=0):
 * normalises the quaternion and produces the initial qacc_warmstart. */
int dial_pipeline_init(dial_plan* plan, const float* qpos, const float* qvel,
                       float* qpos_out, float* warm_out, void* stream);

/* Stage 1 of MBDPI.reverse_once (core/dial_core.py:103-125): sample Y0s
 * (eps injected [dev][Ntotal,Hnode+1,nu], or NULL -> Threefry stream keyed by
 * `key`), pin node 0, append the mean row, clip, spline to actions, roll out
 * this rank's shard + the mean sample and write per-sample mean rewards
 * rews_local [dev][Nsample+1] (mean sample last).  Trajectories
 * (q/qd/xpos [Nsample+1,Hs+1,*]) are kept in the plan's workspace, which is double-buffered:
 * dial_reverse_trajbar enqueued (on any stream, after the matching update) before the NEXT
 * dial_reverse_rollout reads the buffer of this rollout, so it may overlap the next rollout;
 * the buffer is reused by the rollout after next. */
int dial_reverse_rollout(dial_plan* plan, const dial_state* s, const float* eps,
                         const uint32_t key[2], const float* Ybar /*[dev][Hn+1,nu]*/,
                         const float* noise_scale /*[dev][Hn+1]*/, float* rews_local,
                         void* stream);

/* Stage 2 (core/dial_core.py:126-135): population std, softmax weights over the
 * Ntotal+1 rewards rews_all [dev] (sample-major, mean sample last) and the weighted
 * control update Ybar_out [dev][Hn+1,nu].  Every rank recomputes all Y0s from
 * eps/key, so sharded runs need only the one allgather of rewards.  Optional
 * (nullable) weights [dev][Ntotal+1]. */
int dial_reverse_update(dial_plan* plan, const float* eps, const uint32_t key[2],
                        const float* Ybar, const float* noise_scale, const float* rews_all,
                        float* Ybar_out, float* weights, void* stream);

/* dial_reverse_update with the rewards taken from this rank's exchange mailbox when
 * rews_all == NULL (see dial_exchange_*); rews_gathered (nullable) [dev][Ntotal+1] receives a
 * compact copy of the gathered rewards (the `rews` of the reference's info dict). */
int dial_reverse_update_x(dial_plan* plan, const float* eps, const uint32_t key[2],
                          const float* Ybar, const float* noise_scale, const float* rews_all,
                          float* Ybar_out, float* weights, float* rews_gathered, void* stream);

/* qbar/qdbar/xbar (core/dial_core.py:133-135): weighted sums of this rank's stored
 * trajectories with `weights` [dev][Ntotal+1]; sharded runs sum the outputs across
 * ranks (the mean sample is counted on rank 0 only).  Outputs [dev]:
 * qbar [Hs+1,nq], qdbar [Hs+1,nv], xbar [Hs+1,nbody-1,3]. */
int dial_reverse_trajbar(dial_plan* plan, const float* weights, int rank,
                         float* qbar, float* qdbar, float* xbar, void* stream);

/* The stored trajectories of the LAST dial_reverse_rollout (the `pipeline_statess` that
 * core/dial_core.py:120-124 keeps alive: q, qd, x.pos of every sample): device-to-device copies
 * into caller buffers q [Nsample+1,Hs+1,nq], qd [..,nv], xpos [..,nbody-1,3] (each nullable). */
int dial_reverse_trajectories(dial_plan* plan, float* q, float* qd, float* xpos, void* stream);

/* ---- Multi-GPU reward exchange over NVLink peer memory (one process per GPU) -----------------
 * The reference is single-device; sharding the samples needs ONE exchange per reverse_once: all
 * ranks need all Ntotal rewards for std / softmax (core/dial_core.py:125-128).  Instead of a host-
 * issued collective the exchange is fused into the kernels on either side of it: the epilogue of
 * the rollout kernel stores each finished row's reward straight into the mailbox of EVERY rank
 * (peer stores through NVSwitch) and its last CTA raises a flag per rank; the weights kernel
 * spins on the W flags of its own mailbox (bounded), then reads the rewards locally.  The
 * info-only bars are summed the same way (push partial, flag, wait, add in rank order).  No host
 * round trip, no extra launch: a sharded control step is graph-capturable (dial_mpc_step).
 *   1. every rank: dial_exchange_create(plan, rank, world, handle)   -> 64-byte CUDA IPC handle
 *   2. all-gather the handles with any host transport (torch.distributed, MPI, a file)
 *   3. every rank: dial_exchange_connect(plan, handles[world][64])
 * From then on dial_reverse_rollout publishes, dial_reverse_update(_x) with rews_all == NULL
 * consumes, dial_reverse_trajbar returns the all-rank sums.  dial_exchange_status reads
 * {sequence, CTAs done, error (1 = a wait timed out after ~4 s), bars sequence, total ns the
 * update kernels waited for their slowest peer's flag, the same for the bars kernels}. */
int dial_exchange_create(dial_plan* plan, int rank, int world, unsigned char handle_out[DIAL_IPC_HANDLE_BYTES]);
int dial_exchange_connect(dial_plan* plan, const unsigned char* handles);
int dial_exchange_status(dial_plan* plan, uint32_t out[6]);

/* ---- Device-resident synchronous MPC loop -------------------------------------------------
 * The reference's main loop (core/dial_core.py:242-268) is, per control step,
 *     state = step_env(state, Y0[0]); Y0 = shift(Y0);
 *     for i < n_diffuse: rng, Y0, info = reverse_once(state, rng, Y0, noise[i])
 * with one jitted XLA program per piece.  Here the whole step is ONE CUDA graph: state, step
 * counters, rng and control knots stay in HBM (caller-owned `dial_mpc_buffers`), the key
 * splitting / shift / counter bookkeeping between kernels runs as tiny glue kernels, and
 * `dial_mpc_step` only replays the graph (captured on the second use of an
 * (n_diffuse, env_step) shape; the first use runs eagerly).  Results equal the eager
 * `dial_env_step` + `dial_reverse_*` sequence.  Sharded plans (Ntotal > Nsample) need a connected
 * exchange: every rank replays the same graph, the env step runs redundantly on every rank. */
typedef struct dial_mpc_buffers { /* all [dev], caller-owned, fixed while bound */
  float* qpos;            /* [nq]  state, advanced in place by the env step                  */
  float* qvel;            /* [nv]                                                            */
  float* qacc_warmstart;  /* [nv]                                                            */
  int32_t* counters;      /* [2]   info["step"], info["contact_stage"]; advanced in place    */
  uint32_t* rng;          /* [2]   planner rng; each reverse_once splits it                  */
  float* Y;               /* [Hn+1,nu] control knots, in/out                                 */
  float* ctrl;            /* [nu]  out: control applied by the env step                      */
  float* reward;          /* [1]   out: reward of the env step                               */
  float* rews;            /* [Nsample+1] out: this rank's sample rewards of the last reverse_once */
  float* rews_all;        /* [Ntotal+1]  out, sharded plans only (else NULL): all ranks' rewards  */
  float* qbar;            /* [Hs+1,nq]        nullable (all three or none): bars of the last  */
  float* qdbar;           /* [Hs+1,nv]        reverse_once                                    */
  float* xbar;            /* [Hs+1,nbody-1,3]                                                 */
  const float* noise;     /* [>= n_diffuse][Hn+1] annealing schedule (dial_core.py:259-261)  */
} dial_mpc_buffers;

/* Bind the state block; M_shift [host][Hn+1][Hn+1] = u2node . roll(-1, last row 0) . node2u
 * (MBDPI.shift, core/dial_core.py:160-165).  Drops previously captured graphs. */
int dial_mpc_bind(dial_plan* plan, const dial_mpc_buffers* buffers, const float* M_shift);

/* One control step on `stream`: [env_step == 1: state <- env.step(state, Y[0])], [env_step == 1
 * or 2: Y <- shift(Y)], then n_diffuse x reverse_once with noise rows 0..n_diffuse-1.
 * env_step = 0 plans from the bound state as it is (deploy/dial_plan.py: the state comes from
 * the robot); env_step = 2 shifts and plans without advancing the state (a planner whose state
 * is written by somebody else once per control period). */
int dial_mpc_step(dial_plan* plan, int n_diffuse, int env_step, void* stream);

/* jax.random.split(rng) / the planner's key threading (core/dial_core.py:106,145):
 * host-side Threefry-2x32; out[0] is the new rng, out[1] the sampling key. */
void dial_key_split(const uint32_t key[2], uint32_t out0[2], uint32_t out1[2]);

/* tuning aid: with DIAL_DEBUG_COUNTERS=1 in the environment at plan creation the dense solver
 * path counts [0] physics steps and [1] Newton iterations; reads and resets the counters. */
int dial_debug_counters(dial_plan* plan, float out[8]);

/* Which solver instantiation `model` maps to (1 star<3,6>, 2 star<5,7>, 3 dense nv=22,
 * 4 star<5,6>, 0 generic tree, <0 unsupported): custom-reward builds compile only this one
 * (-DDIAL_ONLY_VARIANT=v). */
int dial_solver_variant(const dial_model_desc* model);

/* "" for the stock library; the identifier (-DDIAL_CUSTOM_REWARD_ID) of the reward source a
 * custom build was compiled with.  Only such a build accepts env_id == DIAL_ENV_CUSTOM. */
const char* dial_custom_reward_id(void);

/* Measured fp32 FFMA throughput of the current device in TFLOP/s (independent FMA chains at full
 * occupancy, best of 3, CUDA events; synchronous): the roofline denominator bench.py reports the
 * rollout kernel's flop rate against.  iters = loop trips of 128 FFMAs per thread. */
int dial_fp32_peak(int iters, float* tflops_out);

/* kernel launches issued by this plan since creation (bench bookkeeping) */
int64_t dial_launch_count(const dial_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* DIAL_B200_H_ */
