/*
 * qmpc.h — C-ABI of the batched B200 quaternion-MPC solver.
 *
 * Drop-in boundary for the GRF solve of zixinz990/quaternion-mpc's legged_ctrl package:
 *   legged::LeggedMpc::grf_update(LeggedState&)      legged_ctrl/include/mpc/LeggedMpc.h:26
 *     QuatMpc::grf_update                            legged_ctrl/src/mpc/QuatMpc.cpp:109-276
 *     ConvexMpc::grf_update                          legged_ctrl/src/mpc/ConvexMpc.cpp:81-198
 * The structs below carry exactly the LeggedState fields those two functions read and write
 * (file:line next to every field).  Everything is plain C: fixed-width types, caller-owned
 * buffers, no exceptions, int return codes (0 = ok, <0 = error), per-problem `status`.
 *
 * One handle = one CUDA device + one solver configuration + all workspace (allocated once in
 * qmpc_create, nothing is allocated per call).  Calls on one handle are not concurrent; use one
 * handle per calling thread / stream.  There is NO CPU fallback: if the CUDA device or kernel
 * image is unavailable the calls fail with QMPC_ERR_CUDA.
 */
#ifndef QMPC_H_
#define QMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMPC_ABI_VERSION 3

/* ---- models ------------------------------------------------------------- */
#define QMPC_MODEL_QUAT_4FOOT   0 /* QuatMpc: 13-state quaternion SRB, 4 feet, 24 cone rows
                                     (AltroUtils.cpp:363-439, QuatMpc.cpp:109-276) */
#define QMPC_MODEL_QUAT_2FOOT   1 /* 2-contact quaternion SRB, m=6, 12 cone rows
                                     (AltroUtils.cpp:441-513, TestAltroTrotQuatMpc.cpp) */
#define QMPC_MODEL_EULER_CONVEX 2 /* ConvexMpc: 12-state Euler SRB, LQR cost
                                     (AltroUtils.cpp:224-359, ConvexMpc.cpp:81-198) */

#define QMPC_MAX_HORIZON 32

/* ---- return codes ----------------------------------------------------------------------------- */
#define QMPC_OK            0
#define QMPC_ERR_ARG      -1 /* null pointer, bad model / horizon / batch */
#define QMPC_ERR_CUDA     -2 /* CUDA runtime error (no device, launch failure, ...) */
#define QMPC_ERR_CAPACITY -3 /* batch larger than the handle's max_batch */

/* ---- per-problem solve status (ALTRO's SolveStatus is discarded by the reference,
 *      QuatMpc.cpp:256; we report it) ------------------------------------------------------------ */
#define QMPC_STATUS_SUCCESS            0 /* stationarity and feasibility tolerances met */
#define QMPC_STATUS_MAX_ITERATIONS     1 /* stopped at iterations_max (the usual case with active cones) */
#define QMPC_STATUS_LINESEARCH_FAILED  2 /* no step length satisfied the Armijo test; last iterate returned */
#define QMPC_STATUS_BACKWARD_FAILED    3 /* Quu not positive definite; last iterate returned */
#define QMPC_STATUS_NONFINITE          4 /* NaN/Inf in the inputs or the initial rollout */

/* Solver + robot configuration, constant over a batch.  Defaults: qmpc_default_config(). */
typedef struct QmpcConfig {
  int32_t model;             /* QMPC_MODEL_*                                                     */
  int32_t horizon;           /* N = state.param.mpc_horizon                    QuatMpc.cpp:35      */
  double  dt;                /* seconds = mpc_update_period/1000; rounded to float32 inside, as
                                ALTRO passes `float h`                          QuatMpc.cpp:224,
                                                                                AltroUtils.cpp:10   */
  double  q_weights[13];     /* state.param.q_weights (12 used by EULER_CONVEX) QuatMpc.cpp:227     */
  double  r_weights[12];     /* state.param.r_weights                                              */
  double  w;                 /* quaternion (geodesic) weight state.param.w                         */
  double  mu;                /* friction coefficient                            QuatMpc.cpp:37      */
  double  fz_max;            /*                                                 QuatMpc.cpp:38      */
  double  robot_mass;        /* state.param.robot_mass                          QuatMpc.cpp:122,185 */
  double  inertia[9];        /* row-major body inertia ACTUALLY used by the dynamics, i.e.
                                1.2 * trunk_inertia for QuatMpc                 QuatMpc.cpp:182     */
  double  com_offset[3];     /* body_com (0.0223, 0.002, -0.0005)               AltroUtils.cpp:373  */
  double  com_mass;          /* 5.204                                           AltroUtils.cpp:374  */
  double  gravity;           /* 9.81                                                               */
  double  quat_d_dt;         /* desired-attitude integration step, fixed 5 ms   QuatMpc.cpp:132     */
  /* AltroOptions in force (QuatMpc.cpp:21-26, ConvexMpc.cpp:36-38; rest = library defaults) */
  int32_t iterations_max;    /* 10 (QuatMpc) / 5 (ConvexMpc)                                       */
  int32_t drop_omega0;       /* 1 = reproduce the reference's x_init quirk: measured angular
                                velocity is NOT put into x0                     QuatMpc.cpp:232-245 */
  double  penalty_initial;   /* 1.0                                                                */
  double  penalty_scaling;   /* 20.0 (QuatMpc) / 10.0 (ConvexMpc)                                  */
  double  penalty_max;       /* 1e8                                                                */
  double  tol_cost_intermediate;   /* 1e-4 */
  double  tol_primal_feasibility;  /* 1e-4 */
  double  tol_stationarity;        /* 1e-4 */
} QmpcConfig;

/* One QuatMpc solve: the LeggedState fields QuatMpc::grf_update reads (QUAT_4FOOT / QUAT_2FOOT). */
typedef struct QmpcProblem {
  double  torso_quat[4];           /* fbk.torso_quat (w,x,y,z)                  QuatMpc.cpp:236-239 */
  double  torso_lin_vel_world[3];  /* fbk.torso_lin_vel_world                   QuatMpc.cpp:231     */
  double  torso_ang_vel_body[3];   /* fbk.torso_ang_vel_body (unused if drop_omega0) :243-245       */
  double  foot_pos_body[12];       /* fbk.foot_pos_body, 3x4 column-major FL,FR,RL,RR  :185
                                      (QUAT_2FOOT uses the first two columns)                       */
  double  torso_pos_d_body[3];     /* filtered desired position offset          QuatMpc.cpp:156-158 */
  double  torso_lin_vel_d_body[3]; /* filtered desired velocity                 QuatMpc.cpp:156-169 */
  double  torso_quat_d[4];         /* ctrl.torso_quat_d BEFORE this tick's integration  :128-131    */
  double  torso_ang_vel_d_body[3]; /* ctrl.torso_ang_vel_d_body                 QuatMpc.cpp:132     */
  int32_t plan_contacts[4];        /* ctrl.plan_contacts (0/1)                  QuatMpc.cpp:119     */
} QmpcProblem;

/* One ConvexMpc solve: the fields ConvexMpc::grf_update reads (EULER_CONVEX). */
typedef struct QmpcConvexProblem {
  double  torso_euler[3];          /* fbk.torso_euler                           ConvexMpc.cpp:156   */
  double  torso_pos_world[3];      /* fbk.torso_pos_world                       ConvexMpc.cpp:159   */
  double  torso_ang_vel_world[3];  /* fbk.torso_ang_vel_world                   ConvexMpc.cpp:162   */
  double  torso_lin_vel_world[3];  /* fbk.torso_lin_vel_world                   ConvexMpc.cpp:165   */
  double  foot_pos_abs_com[12];    /* fbk.foot_pos_abs_com 3x4 column-major     ConvexMpc.cpp:117   */
  double  torso_rot_mat[9];        /* fbk.torso_rot_mat row-major (output map)  ConvexMpc.cpp:191   */
  double  torso_pos_d_world[3];    /* ctrl.torso_pos_d_world                    ConvexMpc.cpp:99-101 */
  double  torso_lin_vel_d_world[3];/* ctrl.torso_lin_vel_d_world                ConvexMpc.cpp:105   */
  double  yaw_rate_d;              /* ctrl.torso_ang_vel_d_body[2]              ConvexMpc.cpp:98    */
  int32_t plan_contacts[4];        /* ctrl.plan_contacts                        ConvexMpc.cpp:91    */
  int32_t pad_[2];
} QmpcConvexProblem;

/* What grf_update writes back. */
typedef struct QmpcResult {
  double  grf_body[12];      /* ctrl.optimized_input[0:12]   QuatMpc.cpp:269 / ConvexMpc.cpp:191  */
  double  grf_world[12];     /* ctrl.mpc_grf_world           QuatMpc.cpp:268                      */
  double  torso_quat_d[4];   /* ctrl.torso_quat_d after the 5 ms integration, QuatMpc.cpp:133-137
                                (unchanged identity for EULER_CONVEX)                             */
  double  max_violation;     /* max_k max(0, c) of the cone rows at the returned iterate          */
  int32_t iterations;        /* AL-iLQR iterations executed                                       */
  int32_t status;            /* QMPC_STATUS_*                                                     */
} QmpcResult;

typedef struct QmpcHandle QmpcHandle;

/* Fill cfg with the shipped Go1 values: legged_ctrl/config/gazebo_go1_quat_mpc.yaml:35-75,115-122
 * and QuatMpc.cpp:21-26 for QUAT_*;  gazebo_go1_convex_mpc.yaml + ConvexMpc.cpp:36-38 for
 * EULER_CONVEX.  `horizon` <= QMPC_MAX_HORIZON. */
int qmpc_default_config(int32_t model, int32_t horizon, QmpcConfig* cfg);

/* Create a solver on CUDA device `device` able to solve up to `max_batch` problems per call.
 * Replaces: QuatMpc::QuatMpc (QuatMpc.cpp:8-55) + the per-call `ALTROSolver solver(horizon)`
 * construction (QuatMpc.cpp:218-229). */
int qmpc_create(const QmpcConfig* cfg, int32_t max_batch, int32_t device, QmpcHandle** out);

/* qmpc_create with explicit, per-handle build choices.  The library reads NO environment variables:
 * everything that selects a kernel or a launch geometry is resolved here, once, and stored in the
 * handle (the 200 Hz mpc_thread of Main.cpp:88-120 must not see hidden dispatch on its solve path).
 * `opt` may be NULL (= qmpc_create).  The dense and srb kernels are the on-device cross-checks of the
 * tests (independent implementations of the same solve); the product default is QMPC_KERNEL_AUTO. */
#define QMPC_KERNEL_AUTO   -1 /* coop for every model                                              */
#define QMPC_KERNEL_DENSE   0 /* generic dense algebra, one thread per problem: cross-check, compiled only
                                 into the test-only libqmpc_b200_xcheck.so (QMPC_ERR_ARG in the product)  */
#define QMPC_KERNEL_SRB     1 /* structured, one thread per problem, QUAT models only: cross-check, ditto    */
#define QMPC_KERNEL_COOP    2 /* 16 lanes per problem, shared-memory resident, one persistent launch */
#define QMPC_KERNEL_PHASED  3 /* coop bodies split into set-up / backward / forward launches        */
typedef struct QmpcCreateOptions {
  int32_t kernel;           /* QMPC_KERNEL_*                                                        */
  int32_t smem_residents;   /* -1 = chosen by occupancy query; else bit 0: per-knot linearisation
                               blocks, bit 1: duals kept in shared memory (coop kernel)            */
  int32_t packed_launch;    /* 1 = fill blocks one by one instead of spreading a partial wave       */
  int32_t host_chunks;      /* *_host entry points: 0 / 1 = one copy-solve-copy sequence (default); 2..4 = a batch of
                               several problem waves is copied and solved in up to that many chunks of whole waves,
                               copies overlapping the solves (measured: no gain on a B200, see qmpc_api.cu)        */
} QmpcCreateOptions;
int qmpc_create_ex(const QmpcConfig* cfg, int32_t max_batch, int32_t device, const QmpcCreateOptions* opt,
                   QmpcHandle** out);

/* Solve `batch` independent problems.  `d_in` / `d_out` are DEVICE pointers; the launch is
 * enqueued on `cuda_stream` (a cudaStream_t, may be NULL = default stream) and the call returns
 * without synchronising.  Replaces the body of QuatMpc::grf_update (QuatMpc.cpp:109-276). */
int qmpc_solve_batch(QmpcHandle* h, const QmpcProblem* d_in, int32_t batch, QmpcResult* d_out,
                     void* cuda_stream);
int qmpc_solve_batch_convex(QmpcHandle* h, const QmpcConvexProblem* d_in, int32_t batch,
                            QmpcResult* d_out, void* cuda_stream);

/* Same with HOST buffers: H2D copy, solve, D2H copy, synchronise.  This is the call the
 * CudaQuatMpc shim makes with batch = 1 from the 200 Hz mpc_thread (Main.cpp:88-120). */
int qmpc_solve_batch_host(QmpcHandle* h, const QmpcProblem* in, int32_t batch, QmpcResult* out);
int qmpc_solve_batch_convex_host(QmpcHandle* h, const QmpcConvexProblem* in, int32_t batch,
                                 QmpcResult* out);

/* ================================================================================================
 * Rows "next" of the hot-path scope: the data either side of the solve.
 * ================================================================================================ */

/* ---- N1: per-step contact schedules -------------------------------------------------------------
 * The reference plans with ONE contact mask held over the horizon (QuatMpc.cpp:119-125,202) and
 * flags the per-step schedule as TODO (ConvexMpc.cpp:82); LeggedContactFSM::predict_contact_state
 * (LeggedContactFSM.cpp:272-286) exists but is unused.  mask[k] bit i = foot i (FL,FR,RL,RR) in
 * contact at knot k, k < horizon.  Knot k uses its own mask for u_ref (weight shared by the feet in
 * contact; 0 if none) and for the fz bound (fz <= fz_max * contact); the initial input guess stays
 * SetInput(u_traj_ref.at(0)) (QuatMpc.cpp:253).  A schedule that repeats plan_contacts at every
 * knot gives bit-identical results to the plain entry points. */
typedef struct QmpcContactSchedule {
  uint8_t mask[QMPC_MAX_HORIZON];
} QmpcContactSchedule;

#define QMPC_GAIT_TROT            0 /* set_default_gait_pattern          LeggedContactFSM.cpp:87-108  */
#define QMPC_GAIT_TROT_WITH_STAND 1 /* set_trot_with_stand_gait_pattern  LeggedContactFSM.cpp:110-150 */
#define QMPC_GAIT_CRAWL           2 /* set_crawl_gait_pattern            LeggedContactFSM.cpp:152-193 */
#define QMPC_GAIT_STAND           3 /* set_default_stand_pattern         LeggedContactFSM.cpp:195-206 */

/* The per-robot state of the four LeggedContactFSM objects that predict_contact_state reads. */
typedef struct QmpcGaitState {
  double  gait_phase[4];   /* leg_FSM[i].gait_phase, 0..1                LeggedContactFSM.h:76       */
  double  gait_freq;       /* cycles per second (param.gait_freq)        LeggedState.cpp:77          */
  int32_t gait;            /* QMPC_GAIT_*                                                            */
  int32_t pad_;
} QmpcGaitState;

/* mask[k] bit i = (leg_FSM[i].predict_contact_state(k * cfg.dt) == STANCE), k = 0..horizon-1;
 * bytes k >= horizon are 0.  Device pointers; enqueued on `cuda_stream`. */
int qmpc_predict_contact_schedule(QmpcHandle* h, const QmpcGaitState* d_gait, int32_t batch,
                                  QmpcContactSchedule* d_sched, void* cuda_stream);

/* qmpc_solve_batch / _convex with a per-problem schedule (d_sched may be NULL = plain solve). */
int qmpc_solve_batch_sched(QmpcHandle* h, const QmpcProblem* d_in, const QmpcContactSchedule* d_sched,
                           int32_t batch, QmpcResult* d_out, void* cuda_stream);
int qmpc_solve_batch_convex_sched(QmpcHandle* h, const QmpcConvexProblem* d_in,
                                  const QmpcContactSchedule* d_sched, int32_t batch, QmpcResult* d_out,
                                  void* cuda_stream);
int qmpc_solve_batch_sched_host(QmpcHandle* h, const QmpcProblem* in, const QmpcContactSchedule* sched,
                                int32_t batch, QmpcResult* out);
int qmpc_solve_batch_convex_sched_host(QmpcHandle* h, const QmpcConvexProblem* in,
                                       const QmpcContactSchedule* sched, int32_t batch, QmpcResult* out);

/* ---- N2: leg kinematics in, joint torques out ---------------------------------------------------
 * Producer of the solve's foot_pos_body and consumer of its GRFs:
 *   fbk.foot_pos_body(:, i) = a1_kin.fk(joint_pos_i, rho_opt_i, rho_fix_i)   BaseInterface.cpp:204-207
 *   fbk.jac_foot(:, 3i:3i+3) = a1_kin.jac(joint_pos_i, ...)                  BaseInterface.cpp:208-212
 *   ctrl.joint_tau_tgt_i = -jac_i^T * optimized_input_i   (0 for a swing leg when movement_mode > 0)
 *                                                                            BaseInterface.cpp:343-405 */
typedef struct QmpcLegParams {
  double rho_fix[4][5];  /* per leg: leg_offset_x, leg_offset_y, motor_offset, upper_leg_length,
                            lower_leg_length                                  BaseInterface.cpp:12-30 */
  double rho_opt[4][3];  /* foot contact offset, zero in the reference       BaseInterface.cpp:31    */
} QmpcLegParams;
int qmpc_default_leg_params(QmpcLegParams* lp); /* Go1 values, BaseInterface.cpp:12-33, LeggedParams.h */

/* d_joint_pos: batch x 12 (FL hip, thigh, calf, FR ..., RL ..., RR ...).
 * d_foot_pos_body: batch x 12, 3x4 column-major (layout of QmpcProblem.foot_pos_body), may be NULL.
 * d_jac_foot: batch x 36, the 3x12 Eigen (column-major) fbk.jac_foot, may be NULL. */
int qmpc_leg_kinematics(QmpcHandle* h, const QmpcLegParams* lp, const double* d_joint_pos, int32_t batch,
                        double* d_foot_pos_body, double* d_jac_foot, void* cuda_stream);

/* d_tau: batch x 12 joint torque targets from the solve results.  d_plan_contacts: batch x 4 (the
 * QmpcProblem.plan_contacts values) or NULL = all legs in stance; movement_mode as ctrl.movement_mode. */
int qmpc_joint_torques(QmpcHandle* h, const QmpcResult* d_results, const double* d_jac_foot,
                       const int32_t* d_plan_contacts, int32_t movement_mode, int32_t batch, double* d_tau,
                       void* cuda_stream);

/* ---- N3: reference generation ------------------------------------------------------------------
 * The step immediately before the solve: QuatMpc::goal_update (QuatMpc.cpp:68-107) turns joystick
 * commands and the torso feedback into the filtered references QuatMpc::grf_update reads
 * (torso_pos_d_body / torso_lin_vel_d_body through six MovingWindowFilter(100) channels,
 * MovingWindowFilter.hpp:16-71; torso_ang_vel_d_body), and the Raibert heuristic
 * (BaseInterface.cpp:265-288) places the foot-hold targets.  Both are batched here with the
 * per-robot controller state (desired torso position, filter windows) resident on the device. */
typedef struct QmpcGoalInput {
  double joy_vel[2];             /* state.joy.velx, vely                         QuatMpc.cpp:80-81     */
  double joy_ang_rate[3];        /* state.joy.roll_rate, pitch_rate, yaw_rate    QuatMpc.cpp:93-95     */
  double joy_body_height;        /* state.joy.body_height                        QuatMpc.cpp:100       */
  double torso_pos_world[3];     /* fbk.torso_pos_world                          QuatMpc.cpp:74,102    */
  double torso_quat[4];          /* fbk.torso_quat (w,x,y,z); torso_rot_mat and the yaw-only
                                    torso_rot_mat_z are derived as in           BaseInterface.cpp:196-200 */
  double torso_lin_vel_world[3]; /* fbk.torso_lin_vel_world                      BaseInterface.cpp:266 */
} QmpcGoalInput;

typedef struct QmpcRaibertParams {
  double gait_freq;                  /* param.gait_freq                          LeggedState.cpp:77    */
  double default_foot_pos_rel[12];   /* param.default_foot_pos_rel, 3x4 column-major  yaml:16-30       */
  double delta_x_limit, delta_y_limit; /* FOOT_DELTA_X_LIMIT / _Y_LIMIT          LeggedParams.h        */
} QmpcRaibertParams;
int qmpc_default_raibert_params(QmpcRaibertParams* rp);

/* Bytes of device memory holding the goal_update state of max_batch robots (zero-filled = freshly
 * constructed controllers: no desired position yet, empty filter windows). */
int64_t qmpc_goal_state_bytes(const QmpcHandle* h);
/* One goal_update tick for `batch` robots (robot i keeps using slot i of d_goal_state).  Writes
 * torso_pos_d_body, torso_lin_vel_d_body (filtered), torso_ang_vel_d_body, torso_quat and
 * torso_lin_vel_world of d_problems[i]; the other QmpcProblem fields are left untouched. */
int qmpc_goal_update(QmpcHandle* h, void* d_goal_state, const QmpcGoalInput* d_in, int32_t batch,
                     QmpcProblem* d_problems, void* cuda_stream);
/* ctrl.foot_pos_target_world / foot_pos_target_rel (batch x 12 each, 3x4 column-major; either may
 * be NULL) from the Raibert heuristic. */
int qmpc_raibert_targets(QmpcHandle* h, const QmpcRaibertParams* rp, const QmpcGoalInput* d_in, int32_t batch,
                         double* d_foot_pos_target_world, double* d_foot_pos_target_rel, void* cuda_stream);

/* ---- N3, gait-FSM half: QuatMpc::foot_update (QuatMpc.cpp:278-305) --------------------------------------
 * One tick of the four LeggedContactFSM objects of every robot (LeggedContactFSM.cpp:10-78, 208-260): phase
 * advance, stance -> swing at the pattern's switch time, swing -> stance at the end of the swing or on early
 * contact (more than 90 % through the swing and the foot-force flag set), swing-foot targets from the quintic
 * curve (Utils.cpp:236-293), stance feet hold the touch-down position.  The per-leg state (contact state, gait
 * phase, pattern index, swing start / end, targets) is resident on the device; the outputs are what the reference
 * writes into the LeggedState: ctrl.plan_contacts, ctrl.gait_counter and the FSM targets passed through by
 * grf_update (QuatMpc.cpp:270-272).  movement_mode == 0 resets the FSMs and plans all feet in contact (:283-289). */
typedef struct QmpcFootUpdateInput {
  double  foot_pos_world[12];         /* fbk.foot_pos_world, 3x4 column-major          QuatMpc.cpp:293 */
  double  foot_pos_target_world[12];  /* ctrl.foot_pos_target_world (Raibert targets)  QuatMpc.cpp:294 */
  int32_t foot_contact_flag[4];       /* fbk.foot_contact_flag                         QuatMpc.cpp:295 */
  int32_t movement_mode;              /* ctrl.movement_mode                            QuatMpc.cpp:283 */
  int32_t pad_[3];
} QmpcFootUpdateInput;
typedef struct QmpcFootUpdateOutput {
  double  foot_pos_target[12];   /* FSM_foot_pos_target_world -> ctrl.optimized_state[6 + 3 i]   QuatMpc.cpp:270 */
  double  foot_vel_target[12];   /* FSM_foot_vel_target_world -> ctrl.optimized_input[12 + 3 i]  QuatMpc.cpp:271 */
  double  foot_acc_target[12];   /* FSM_foot_acc_target_world -> ctrl.optimized_input[24 + 3 i]  QuatMpc.cpp:272 */
  double  gait_counter[4];       /* ctrl.gait_counter = the leg's gait phase                     QuatMpc.cpp:292 */
  int32_t plan_contacts[4];      /* ctrl.plan_contacts (1 = STANCE)                              QuatMpc.cpp:300 */
} QmpcFootUpdateOutput;
/* Bytes of device memory holding the four leg FSMs of max_batch robots. */
int64_t qmpc_leg_fsm_state_bytes(const QmpcHandle* h);
/* reset_params + reset for `batch` robots (LeggedContactFSM.cpp:4-31): robot i gets the gait pattern d_gait[i]
 * (QMPC_GAIT_*; NULL = the default trot of reset_params). */
int qmpc_leg_fsm_init(QmpcHandle* h, void* d_fsm_state, const int32_t* d_gait, int32_t batch, void* cuda_stream);
/* One foot_update tick: dt = the FSM step (5 ms in QuatMpc.cpp:292, h in ConvexMpc.cpp:208), gait_freq =
 * param.gait_freq.  d_problems (may be NULL): plan_contacts of the solve inputs are written in place;
 * d_gait_out (may be NULL): the gait states qmpc_predict_contact_schedule reads - so that a tick
 * goal_update -> foot_update -> predict schedule -> solve -> torques never leaves the device. */
int qmpc_foot_update(QmpcHandle* h, void* d_fsm_state, const QmpcFootUpdateInput* d_in, double dt, double gait_freq,
                     int32_t batch, QmpcFootUpdateOutput* d_out, QmpcProblem* d_problems, QmpcGaitState* d_gait_out,
                     void* cuda_stream);

/* ---- N4: warm start (trajectory shift) ------------------------------------------------------------
 * legged_ctrl builds a fresh ALTROSolver every tick and starts from u_ref (QuatMpc.cpp:218,253); the
 * ALTRO API offers ShiftTrajectory() for receding-horizon use (pattern shown in
 * test_altro/TestBicycle.cpp:181-199) which the controller never calls.  This extension keeps the
 * previous input trajectory per problem: when `valid`, knot k starts from u_prev[min(k + 1, N - 1)]
 * (one-knot shift, last input repeated) instead of u_ref; duals and penalty start fresh as in the
 * reference.  After the solve the buffer holds the new trajectory (valid = 0 after a non-finite
 * solve).  With valid = 0 the result is bit-identical to the plain entry points. */
typedef struct QmpcWarmStart {
  double  u[QMPC_MAX_HORIZON][12];  /* input trajectory of the last solve (body-frame GRFs per knot) */
  int32_t valid;
  int32_t pad_;
} QmpcWarmStart;

/* qmpc_solve_batch_sched with a per-problem warm-start buffer (device pointers; d_sched may be NULL,
 * d_warm is read and updated in place). */
int qmpc_solve_batch_warm(QmpcHandle* h, const QmpcProblem* d_in, const QmpcContactSchedule* d_sched,
                          QmpcWarmStart* d_warm, int32_t batch, QmpcResult* d_out, void* cuda_stream);

void qmpc_destroy(QmpcHandle* h);

/* ---- multi-GPU host entry point -----------------------------------------------------
 * The solves are independent (QuatMpc.cpp:218 builds a fresh solver per call), so a host batch is cut
 * into contiguous, balanced shards - shard g = [batch*g/G, batch*(g+1)/G) - one per device; every
 * device has its own stream and pinned staging, the copies and launches of all devices are in flight
 * together and the results land in the caller's single `out` array in batch order.  No collective is
 * involved: a C++ caller (the reference's host language) spreads a batch over the GPUs of one box with
 * this one call.  `max_batch` is the capacity of the WHOLE batch.  Works for every model: `in` is an
 * array of QmpcProblem (QUAT_*) or QmpcConvexProblem (EULER_CONVEX) according to cfg->model. */
typedef struct QmpcMultiHandle QmpcMultiHandle;
int  qmpc_create_multi(const QmpcConfig* cfg, int32_t max_batch, const int32_t* devices, int32_t n_devices,
                       QmpcMultiHandle** out);
int  qmpc_solve_batch_host_multi(QmpcMultiHandle* mh, const void* in, int32_t batch, QmpcResult* out);
void qmpc_destroy_multi(QmpcMultiHandle* mh);
int32_t     qmpc_multi_device_count(const QmpcMultiHandle* mh);
int64_t     qmpc_multi_launch_count(const QmpcMultiHandle* mh);
const char* qmpc_multi_last_error(const QmpcMultiHandle* mh);

/* Introspection: number of kernel launches issued by this handle so far; last CUDA error text. */
int64_t     qmpc_launch_count(const QmpcHandle* h);
const char* qmpc_last_error(const QmpcHandle* h);
const char* qmpc_status_string(int32_t status);
/* One-line description of the kernel and launch geometry the handle uses (for logs / bench.py). */
int         qmpc_describe(const QmpcHandle* h, char* buf, int32_t n);
int32_t     qmpc_abi_version(void);

/* Measurement utility (not on the solve path): sustained FP64 / FP32 vector-FMA throughput of
 * `device` in TFLOP/s — the roofline denominator of this FMA-bound path (bench.py). */
int qmpc_measure_fma_peak(int32_t device, double* fp64_tflops, double* fp32_tflops);

#ifdef __cplusplus
}
#endif
#endif /* QMPC_H_ */
