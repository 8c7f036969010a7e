/*
 * ss_model.h -- plain-C model + parameter tables shared by the CPU oracle
 * (oracle/ss_oracle.c) and the CUDA library (steppingstone_b200/csrc).
 *
 * The tables are produced on the host by steppingstone_b200/model.py from the
 * reference's own model files (mocca_envs/data/custom/walker3d.xml, mike.xml,
 * mocca_envs/data/objects/steps/plank.urdf) and the constants the reference
 * sets in code (mocca_envs/robots.py:446-468,702-724 torque limits;
 * mocca_envs/bullet_utils.py:313-332 physics parameters;
 * mocca_envs/bullet_objects.py:44-93 stone material).
 *
 * Body numbering: body 0 is the floating base ("walker3d"/"mike" MJCF root,
 * with its fixed children head/torso merged in); body k (1..21) is the link
 * moved by hinge k-1 of the reference's `ordered_joints`
 * (mocca_envs/robots.py:400-431; order pinned by robots.py:514-521).
 * Fixed children (hands) are merged into their parents.  Every body frame has
 * its origin at its own hinge anchor, so the hinge motion subspace is [axis;0].
 */
#ifndef SS_MODEL_H
#define SS_MODEL_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_NB 22        /* bodies incl. floating base                      */
#define SS_NJ 21        /* hinges == actions (robots.py:474)               */
#define SS_NDOF 27      /* 6 base + 21 hinge velocities                    */
#define SS_NQ 28        /* base xyz + quat(xyzw) + 21 hinge angles         */
#define SS_OBS 60       /* walker3d_envs.py:552-556                        */
#define SS_ACT 21
#define SS_NCOL 4       /* foot capsules (geoms 0..3 of the geom table)    */
#define SS_NGEOM 22     /* collision geoms: capsules and spheres (W3D 22, Mike 20) */
#define SS_MAX_PAIRS 160 /* self-collision geom pairs (W3D 141, Mike 100)  */
#define SS_MAX_STONES 100 /* rendered_step_count upper bound (W3D 100)     */
#define SS_MAX_CONTACTS 8 /* contact points kept per env per substep       */
#define SS_MAX_LIMIT_ROWS 8
#define SS_MT_N 624     /* MT19937 state words                             */
#define SS_GRID 11      /* curriculum grid edge (walker3d_envs.py:522)     */
#define SS_GRID_CELLS 161051 /* 11^5, mike.py:82-84                        */

/* Humanoid topology shared by walker3d.xml and mike.xml (checked at load). */
#define SS_PARENT_TABLE { -1, 0, 1, 2, 3, 4, 5, 6, 7, 3, 9, 10, 11, 12, 0, 14, 15, 16, 0, 18, 19, 20 }

typedef struct ss_model {
    int32_t parent[SS_NB];
    int32_t depth[SS_NB];
    double anchor[SS_NB][3];   /* hinge anchor of body i in its parent's frame            */
    double R0[SS_NB][9];       /* row-major; child-frame coords -> parent-frame at q=0    */
    double axis[SS_NB][3];     /* hinge axis in the child frame (unit)                    */
    double mass[SS_NB];
    double com[SS_NB][3];      /* centre of mass in the body frame                        */
    double inertia[SS_NB][6];  /* about the COM, body axes: xx yy zz xy xz yz             */
    double q_lo[SS_NJ], q_hi[SS_NJ];
    double tau_max[SS_NJ];     /* robots.py:385-388                                       */
    double joint_damping[SS_NJ]; /* MJCF joint damping (walker3d.xml:4)                   */
    double pd_p[SS_NJ], pd_d[SS_NJ]; /* robots.py:482-483 (opt-in PD actuator)            */
    int32_t col_body[SS_NCOL];
    int32_t col_foot[SS_NCOL]; /* 0 right, 1 left                                         */
    double col_radius[SS_NCOL];
    double col_p0[SS_NCOL][3], col_p1[SS_NCOL][3]; /* capsule segment, body frame         */
    int32_t foot_body[2];
    double foot_ref[2][3];     /* point getLinkState()[0] reports for the foot link       */
    double contact_breaking;   /* manifold breaking threshold for foot-vs-stone           */
    double init_q[SS_NJ];      /* "running_start", robots.py:535-545                      */
    double init_pos[3];        /* walker3d_envs.py:823 / mike.py:344                      */
    double total_mass;
    /* Every collision geom of the MJCF (walker3d.xml:17-97): a capsule [p0,p1] or a sphere (p0 == p1) of `radius` in the
     * frame of body geom_body.  Geoms 0..3 are the foot capsules (= col_*).  geom_stone: the link's contype/conaffinity
     * (the LAST geom of the link decides, as in Bullet's MJCF importer) passes the static stones' filter
     * (group StaticFilter, mask AllFilter ^ StaticFilter) -- torso, waist and butt do not.  geom_breaking = 0.02 x the
     * angular-motion disc of the link's compound shape (manifold breaking threshold).  pair_a/pair_b: geom pairs that
     * can self-collide under URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS (robots.py:496-500):
     * different links, neither an ancestor of the other, (group_a & mask_b) && (group_b & mask_a). */
    int32_t n_geom, n_pair;
    int32_t geom_body[SS_NGEOM], geom_foot[SS_NGEOM] /* -1, 0 right, 1 left */, geom_stone[SS_NGEOM], geom_sphere[SS_NGEOM];
    double geom_radius[SS_NGEOM], geom_p0[SS_NGEOM][3], geom_p1[SS_NGEOM][3], geom_breaking[SS_NGEOM];
    int32_t pair_a[SS_MAX_PAIRS], pair_b[SS_MAX_PAIRS];
} ss_model;

typedef struct ss_params {
    /* physics: mocca_envs/bullet_utils.py:313-332, env_base.py:50-55 */
    double gravity;            /* 9.81                                                    */
    double dt;                 /* 1/240                                                   */
    int32_t n_substeps;        /* 4                                                       */
    int32_t n_iters;           /* 20                                                      */
    double erp_limit;          /* joint-limit ERP (Bullet default erp 0.2)                */
    double contact_stiffness;  /* series(30000, robot default) bullet_objects.py:61-69    */
    double contact_damping;    /* 1000 + robot default 0.1                                */
    double friction;           /* 1.2 (walker3d.xml:5) * 1.0 (bullet_objects.py:64)       */
    double linear_slop;
    double warmstart;          /* 0.85                                                    */
    double lin_damping, ang_damping; /* btMultiBody defaults 0.04                         */
    double max_coord_vel;      /* 100                                                     */
    double limit_max_impulse;  /* 100                                                     */
    double stone_half[3];      /* 0.25, 0.375, 0.125: plank.urdf scaled 0.5, boxes merged */
    double stone_box_z;        /* box centre below the top face, link frame: -0.125       */
    double stone_com_z;        /* COM offset below the top face: 0.1375 (q12)             */
    double broadphase_margin;
    /* env: walker3d_envs.py:482-557, mike.py:25-101 */
    int32_t robot_kind;        /* 0 Walker3D, 1 Mike                                      */
    int32_t terrain_mode;      /* 0 W3D as shipped, 1 Mike as shipped, 2 W3D flat         */
    int32_t n_steps;           /* 100 / 40                                                */
    int32_t rendered;          /* 100 / 30                                                */
    int32_t stop_frames;       /* 30                                                      */
    int32_t max_episode_steps; /* 1000 (__init__.py:126-136)                              */
    int32_t n_phantoms;        /* 20 / 0: RNG draws only (walker3d_envs.py:566-580)       */
    int32_t use_sampler;       /* Mike 1 (mike.py:593), W3D 0 (walker3d_envs.py:1083)     */
    int32_t actuator_mode;     /* 0 torque held over the step (robots.py:289-293), 1 the abandoned PD block of
                                  robots.py:486-493 once per control step, 2 that PD law at every substep */
    int32_t collision_mode;    /* 0 feet vs stones only, 1 every link vs stones, 2 + self-collision (reference) */
    double done_height;        /* 0.7                                                     */
    double electricity_cost, stall_torque_cost, joints_at_limit_cost;
    double step_radius;        /* 0.25                                                    */
    double r_max;              /* r_range[1] = 1.5                                        */
    double tilt_limit_deg;     /* 0                                                       */
    double mirror_prob;        /* 0 (robots.py:583) / 0.5 (robots.py:846)                 */
    double control_dt;         /* 1/60 (scene.dt)                                         */
    double yaw_samples[SS_GRID], pitch_samples[SS_GRID], r_samples[SS_GRID];
    double xtilt_samples[SS_GRID], ytilt_samples[SS_GRID];
    /* link-link contacts of the robot with itself: hard contacts (no stiffness/damping on the robot's links) */
    double self_friction;      /* 1.2 * 1.2 (walker3d.xml:5, both links)                  */
    double self_erp;           /* setDefaultContactERP(0.9), bullet_utils.py:324          */
    double stone_breaking;     /* 0.02 x angular-motion disc of the plank's compound      */
    int32_t self_cross_terms;  /* 0: Bullet's row denominator J_a H^-1 J_a^T + J_b H^-1 J_b^T (no cross terms when both
                                  sides are the same multibody); 1: the exact (J_a-J_b) H^-1 (J_a-J_b)^T */
    int32_t reserved1;
} ss_params;

#ifdef __cplusplus
}
#endif
#endif
