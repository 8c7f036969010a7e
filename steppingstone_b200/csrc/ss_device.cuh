// ss_device.cuh -- warp-per-env device code of the B200 stepping-stone environment (sm_100a).
//
// One warp owns one env for a whole control step.  Four lane mappings are used:
//   lane = body  (22 bodies)   kinematics, composite inertia, bias forces      -> tree scans with __shfl
//   lane = dof   (27 dofs)     LDL^T factorisation, velocity / position update
//   lane = candidate (32 per pass)  collision: (geom, plank, segment end) against the stones, geom pairs against each other
//   lane = row   (<=32 rows)   constraint rows, Delassus matrix, projected Gauss-Seidel
// Spatial quantities live in ONE frame per substep: world-aligned axes, origin at the base link's
// origin ("common frame"), so tree accumulations are plain sums (no per-link spatial transforms).
//
// Reference call sites replaced: pybullet.stepSimulation via World.step (mocca_envs/bullet_utils.py:331-332)
// and the env logic of mocca_envs/walker3d_envs.py:875-1093 / mocca_envs/mike.py:396-603.
#pragma once
#ifndef SS_SYNC
#define SS_SYNC 1
#endif
#ifndef SS_SPARSE_LDL
#define SS_SPARSE_LDL 1      // tree-sparse U D U^T factorisation of the joint-space inertia (0: dense LDL^T)
#endif
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "ss_env.h"

#define FULL 0xffffffffu
#define SS_PI_D 3.14159265358979323846
#define SS_DEG_D (SS_PI_D / 180.0)

namespace ssb {

constexpr int NB = SS_NB, NJ = SS_NJ, ND = SS_NDOF;
#ifndef SS_DEL_UNROLL
#define SS_DEL_UNROLL 7      // the Delassus accumulation loop fully unrolled over its 7 dof chunks: every shared-memory address an
#endif                       // immediate, the live-group tests hoisted -- 0.7066 -> 0.6569 ms per 8192-env step (1 = the rolled loop)
constexpr int DEL_UNROLL = SS_DEL_UNROLL;
constexpr int HS = 28;   // row stride of H / L in shared memory: 16-byte aligned rows; a lane reads ITS row with 7 LDS.128
                         // (quarter-warp addresses 112 B apart hit disjoint bank groups -> conflict-free)
constexpr int ZS = 28;   // row stride of Z (same access pattern)
constexpr int MAXC = SS_MAX_CONTACTS, MAXL = SS_MAX_LIMIT_ROWS;

// ---------------------------------------------------------------------------------------------
// constant tables (fp32 copy of ss_model + ss_params), staged into shared memory with one TMA bulk copy
// ---------------------------------------------------------------------------------------------
struct alignas(16) DevModel {
    float anchor[NB][3]; float R0[NB][9]; float axis[NB][3]; float mass[NB]; float com[NB][3]; float inertia[NB][6];
    float q_lo[NJ], q_hi[NJ], tau_max[NJ], damping[NJ], pd_p[NJ], pd_d[NJ], init_q[NJ];
    // collision geoms (capsule / sphere per body), stone candidate table (geom | slot << 5 | end << 7), self-collision pairs
    float geom_p0[SS_NGEOM][3], geom_p1[SS_NGEOM][3], geom_radius[SS_NGEOM], geom_half[SS_NGEOM], geom_breaking[SS_NGEOM];
    unsigned char geom_body[SS_NGEOM + 2];
    unsigned char pair_a[SS_MAX_PAIRS], pair_b[SS_MAX_PAIRS];
    unsigned short scand[SS_NGEOM * 8];
    int n_geom, n_pair, n_scand, collision_mode, self_cross_terms, pad2;
    float stone_breaking, self_friction, self_erp, stone_radius;
    int foot_body[2]; float foot_ref[2][3];
    int anc_mask[NB];
    float contact_breaking, gravity, dt, erp_limit, cfm_contact, erp_contact, friction, linear_slop, warmstart,
        max_coord_vel, limit_max_impulse, stone_half[3], stone_box_off, stone_com_z;
    int n_substeps, n_iters, robot_kind, terrain_mode, n_steps, rendered, stop_frames, max_episode_steps,
        n_phantoms, use_sampler, actuator_mode, pad0;
    float done_height, electricity_cost, stall_torque_cost, joints_at_limit_cost;
    double init_pos[3], step_radius, r_max, tilt_limit_deg, mirror_prob, control_dt;
    double samples[5][SS_GRID];     // yaw, pitch, r, x_tilt, y_tilt
    double pad1;
};
static_assert(sizeof(DevModel) % 16 == 0, "DevModel must be a multiple of 16 bytes for cp.async.bulk");

struct DevState {
    double* pos;        // [N][3]   base position (world, double: courses run to x ~ 70 m)
    float* sv;          // [N][64]  0..3 quat xyzw, 4..24 hinge angles, 32..58 generalised velocity (ang3, lin3, hinge21)
    float* warm;        // [N][SS_NWARM]
    double* terrain;    // [N][n_steps][6]
    double* plank;      // [N][rendered][6]
    double* scal;       // [N][SS_NSCAL]
    int* ints;          // [N][SS_NINT]
    uint32_t* mt;       // [N][2][SS_MT_WORDS]
    const double* cdf;  // [SS_TABLE_RING][n_cells] sampling-table versions
    int* tslot;         // [N] the version each env draws from
    int n_cells;
    double* trace_q;    // [4][N][28] or null
    double* trace_qd;   // [4][N][27] or null
    int* trace_ev;      // [4][N][SS_NEV] or null: limit-row masks and contact ids of each substep
    float* dbg;         // [N][DBG_FLOATS] or null
    int* err;           // [1] sticky error flags (bit 0: a non-finite action reached ss_step), read by ss_error_flags
    int* perm;          // [N] env handled by the k-th warp of the step kernel: envs ordered by their contact count
    int n;
};
constexpr int DBG_H = 0, DBG_C = 729, DBG_CAND = 756, DBG_ROWS = 756 + 256, DBG_FLOATS = 756 + 256 + 4 + 32 * 6 + 32 * 32;

// per-warp scratch
struct alignas(16) WarpSmem {
    float S[NB][8];           // hinge motion vectors (6) padded to 32 bytes: two LDS.128 per body
    float pose[NB][12];       // body poses of this substep: R (9, to world axes), origin rel. the base origin (3)
    float gw[SS_NGEOM][8];    // geom segments in the common frame: p0 (3), p1 (3), radius, pad
    float stone[4][16];       // per candidate slot: box centre rel. step-start base pos (3), R (9), valid, pad
    float vec[64];            // v* (27) / obs staging (60)
    float con[MAXC][16];      // n3, point on A 3, depth, body A, candidate id, warm impulse, point on B 3, body B (-1 stone)
    float lim[MAXL][4];       // hinge, dir, penetration
    float dinv[32];
    double parked[4];         // base position (fp64) parked during the substeps
    float pd[2][32];          // actuator_mode 2: clipped action and the step's joint-damping torque (lane = dof)
    union {
        struct { float H[HS * HS]; float Z[32 * ZS]; } s;
        double d[(HS * HS + 32 * ZS) / 2];
    } u;
};

static_assert(sizeof(((WarpSmem*)0)->u.d) >= 7 * SS_MAX_STONES * sizeof(double), "terrain generator scratch (7 arrays of n_steps doubles) must fit the per-warp union");

// scratch of the dynamics-only kernel (aba_substep): same member names, no constraint-solver storage
struct alignas(16) WarpSmemDyn {
    float S[NB][8];
    float pose[NB][12];
    float dinv[32];
    struct { struct { float H[HS * HS]; } s; } u;
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// Reciprocal / division / square root of the hot path.  The IEEE-rounded forms compile to a MUFU + Newton steps + a
// slow-path check each (~8-10 instructions, 40 of them per env-substep, 27 on the critical path of the factorisation);
// the approximate forms (<= 1-2 ulp) are one MUFU.  The difference is below the fp32 rounding noise the parity
// tolerances already carry (GPU fp32 vs oracle fp64).  -DSS_FAST_DIV=0 restores the IEEE forms.
#ifndef SS_FAST_DIV
#define SS_FAST_DIV 1
#endif
#if SS_FAST_DIV && !defined(SS_SIMT_EMU)
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
__device__ __forceinline__ float frcp(float x) { return 1.f / x; }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float fsqrt(float x) { return sqrtf(x); }
#endif
__device__ __forceinline__ float shfl(float v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }

__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
    float t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
    c[0] = t0; c[1] = t1; c[2] = t2;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void matvec3(const float* R, const float* x, float* y) {
    float t0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
    float t1 = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
    float t2 = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
    y[0] = t0; y[1] = t1; y[2] = t2;
}
__device__ __forceinline__ void matTvec3(const float* R, const float* x, float* y) {
    float t0 = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
    float t1 = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
    float t2 = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
    y[0] = t0; y[1] = t1; y[2] = t2;
}
__device__ __forceinline__ void matmul3(const float* A, const float* B, float* C) {
    float T[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = T[i];
}
__device__ __forceinline__ void quat_to_mat(const float* q, float* R) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}

// topology of the humanoid tree (SS_PARENT_TABLE): chains 1..8, 9..13 (off body 3), 14..17 and 18..21 (off the base)
__device__ __forceinline__ int seg_start(int l) { return l == 0 ? 0 : (l <= 8 ? 1 : (l <= 13 ? 9 : (l <= 17 ? 14 : (l <= 21 ? 18 : l)))); }
__device__ __forceinline__ int seg_end(int l) { return l == 0 ? 0 : (l <= 8 ? 8 : (l <= 13 ? 13 : (l <= 17 ? 17 : (l <= 21 ? 21 : l)))); }

// sum over the subtree rooted at each body (lanes >= 22 must pass 0)
__device__ __forceinline__ float subtree_sum(float x, int lane, int send) {
#pragma unroll
    for (int off = 1; off <= 4; off <<= 1) { float y = __shfl_down_sync(FULL, x, off); if (lane + off <= send) x += y; }
    float t9 = shfl(x, 9); if (lane >= 1 && lane <= 3) x += t9;
    float t1 = shfl(x, 1), t14 = shfl(x, 14), t18 = shfl(x, 18);
    if (lane == 0) x += t1 + t14 + t18;
    return x;
}
// sum over the ancestors of each body including itself and the base (lane 0)
__device__ __forceinline__ float tree_prefix(float x, int lane, int sstart) {
#pragma unroll
    for (int off = 1; off <= 4; off <<= 1) { float y = __shfl_up_sync(FULL, x, off); if (lane >= 1 && lane - off >= sstart) x += y; }
    float t3 = shfl(x, 3); if (lane >= 9 && lane <= 13) x += t3;
    float t0 = shfl(x, 0); if (lane >= 1 && lane < NB) x += t0;
    return x;
}

struct Xf { float R[9]; float p[3]; };
__device__ __forceinline__ Xf xf_shfl_up(const Xf& a, int off) {
    Xf r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.R[i] = __shfl_up_sync(FULL, a.R[i], off);
#pragma unroll
    for (int i = 0; i < 3; i++) r.p[i] = __shfl_up_sync(FULL, a.p[i], off);
    return r;
}
__device__ __forceinline__ Xf xf_shfl(const Xf& a, int src) {
    Xf r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.R[i] = shfl(a.R[i], src);
#pragma unroll
    for (int i = 0; i < 3; i++) r.p[i] = shfl(a.p[i], src);
    return r;
}
// a o b : x -> a.R (b.R x + b.p) + a.p
__device__ __forceinline__ Xf xf_compose(const Xf& a, const Xf& b) {
    Xf r; float t[3];
    matmul3(a.R, b.R, r.R);
    matvec3(a.R, b.p, t);
    r.p[0] = a.p[0] + t[0]; r.p[1] = a.p[1] + t[1]; r.p[2] = a.p[2] + t[2];
    return r;
}

// ---------------------------------------------------------------------------------------------
// forward kinematics: lane b gets the pose of body b in the common frame (rotation to world axes,
// origin relative to the base origin)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ Xf fk_warp(const DevModel& M, const float* quat, float q_dof, int lane) {
    float qb = shfl(q_dof, (lane + 5) & 31);
    Xf X;
#pragma unroll
    for (int i = 0; i < 9; i++) X.R[i] = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
    X.p[0] = X.p[1] = X.p[2] = 0.f;
    if (lane == 0) quat_to_mat(quat, X.R);
    else if (lane < NB) {
        float s, c; sincosf(qb, &s, &c);
        float t = 1.f - c, x = M.axis[lane][0], y = M.axis[lane][1], z = M.axis[lane][2];
        float Rj[9] = { t * x * x + c, t * x * y - s * z, t * x * z + s * y,
                        t * x * y + s * z, t * y * y + c, t * y * z - s * x,
                        t * x * z - s * y, t * y * z + s * x, t * z * z + c };
        matmul3(M.R0[lane], Rj, X.R);
        X.p[0] = M.anchor[lane][0]; X.p[1] = M.anchor[lane][1]; X.p[2] = M.anchor[lane][2];
    }
    int ss = seg_start(lane);
#pragma unroll
    for (int off = 1; off <= 4; off <<= 1) {
        Xf Y = xf_shfl_up(X, off);
        if (lane >= 1 && lane < NB && lane - off >= ss) X = xf_compose(Y, X);
    }
    { Xf Y = xf_shfl(X, 3); if (lane >= 9 && lane <= 13) X = xf_compose(Y, X); }
    { Xf Y = xf_shfl(X, 0); if (lane >= 1 && lane < NB) X = xf_compose(Y, X); }
    return X;
}

// LDL^T elimination step K (lane i owns row i of the symmetric matrix); compile-time recursion keeps Hr in registers
#if !defined(SS_SPARSE_LDL) || !SS_SPARSE_LDL
template <int K>
__device__ __forceinline__ void ldl_steps(float (&Hr)[HS], float& myDinv, int lane) {
    if constexpr (K < ND) {
        float dk = shfl(Hr[K], K);
        float inv = frcp(dk);
        if (lane == K) myDinv = inv;
        float lik = Hr[K] * inv;
#pragma unroll
        for (int j = K + 1; j < ND; j++) {
            float hkj = shfl(Hr[j], K);
            if (lane > K) Hr[j] -= lik * hkj;
        }
        if (lane > K) Hr[K] = lik;
        ldl_steps<K + 1>(Hr, myDinv, lane);
    }
}
#endif

#if SS_SPARSE_LDL
// Branch-induced sparsity (default since r02b: 0.7207 -> 0.6942 ms per 8192-env step on a B200; -DSS_SPARSE_LDL=0 keeps the dense
// row-ordered LDL^T for A/B runs).
// H[i][k] != 0 only if dof i is an ancestor of dof k or vice versa.  Eliminating the dofs leaves-first (k descending)
// creates no fill-in: H = U D U^T with U unit upper triangular and U[i][k] != 0 only for i a proper ancestor of k
// (Featherstone's L^T D L).  206 instead of 351 + 27 non-zeros for this humanoid: the factorisation needs one shuffle
// per (k, ancestor of k) pair and the triangular solves skip the structural zeros at compile time.
__host__ __device__ constexpr bool body_anc(int a, int b) {          // body a is a proper ancestor of body b
    constexpr int parent[NB] = SS_PARENT_TABLE;
    for (int x = parent[b]; x >= 0; x = parent[x]) if (x == a) return true;
    return false;
}
// dof numbering: 0..5 the floating base (a chain in front of everything), 6 + j the hinge of body j + 1
__host__ __device__ constexpr bool dof_anc(int i, int k) { return i < k && k < ND && (i < 6 || body_anc(i - 5, k - 5)); }
__host__ __device__ constexpr unsigned anc_lanes(int k) { unsigned m = 0; for (int i = 0; i < k; i++) if (dof_anc(i, k)) m |= 1u << i; return m; }

template <int K, int J>
__device__ __forceinline__ void udu_cols(float (&Hr)[HS], float uik, bool mine) {
    if constexpr (J >= 0) {
        if constexpr (dof_anc(J, K)) {
            float hkj = shfl(Hr[J], K);
            if (mine) Hr[J] -= uik * hkj;
        }
        udu_cols<K, J - 1>(Hr, uik, mine);
    }
}
// U D U^T elimination step K, K descending (lane i owns the full symmetric row i): after the last step lane i holds
// U[i][k] in Hr[k] for k > i; columns k <= i are scratch
template <int K>
__device__ __forceinline__ void ldl_steps(float (&Hr)[HS], float& myDinv, int lane) {
    if constexpr (K >= 0) {
        float dk = shfl(Hr[K], K);
        float inv = frcp(dk);
        if (lane == K) myDinv = inv;
        constexpr unsigned lanes = anc_lanes(K);
        const bool mine = (lanes >> lane) & 1u;
        float uik = Hr[K] * inv;
        udu_cols<K, K - 1>(Hr, uik, mine);
        if (mine) Hr[K] = uik;
        ldl_steps<K - 1>(Hr, myDinv, lane);
    }
}
#endif

// ---------------------------------------------------------------------------------------------
// dynamics: builds H (joint-space inertia, common frame) and its LDL^T in shared memory, returns the
// unconstrained generalised acceleration in lane = dof layout.
//   in : quat[4] (all lanes), q_dof, v_dof (lane = dof: 0..2 base angular, 3..5 base linear, 6.. hinges),
//        tau_dof (hinge torque for lanes 6..26, else 0)
//   out: acc (lane = dof);  W.u.s.H holds L (unit lower, row stride HS), W.dinv holds 1/D
//        W.S holds the hinge motion vectors, W.pose the body poses
// ---------------------------------------------------------------------------------------------
template <bool DIAG, typename WS>
__device__ __forceinline__ float dynamics_warp(const DevModel& M, WS& W, const float* quat, float q_dof, float v_dof,
                                               float tau_dof, int lane, float* dbg) {
    const int ss = seg_start(lane), se = seg_end(lane);
    const bool isb = lane < NB, isj = lane >= 1 && lane < NB;
    Xf X = fk_warp(M, quat, q_dof, lane);
    // hinge motion vector S = (a ; o x a)
    float S[6] = { 0, 0, 0, 0, 0, 0 };
    if (isj) { matvec3(X.R, M.axis[lane], S); cross3(X.p, S, S + 3); }
    if (isb) {
        float4* Sp = reinterpret_cast<float4*>(W.S[lane]);
        Sp[0] = make_float4(S[0], S[1], S[2], S[3]); Sp[1] = make_float4(S[4], S[5], 0.f, 0.f);
    }
    if (isb) {
#pragma unroll
        for (int i = 0; i < 9; i++) W.pose[lane][i] = X.R[i];
#pragma unroll
        for (int i = 0; i < 3; i++) W.pose[lane][9 + i] = X.p[i];
    }
    // rigid-body inertia about the common origin: m, h = m c, Ibar (xx yy zz xy xz yz)
    float m = 0, h[3] = { 0, 0, 0 }, I[6] = { 0, 0, 0, 0, 0, 0 };
    if (isb) {
        m = M.mass[lane];
        float c[3]; matvec3(X.R, M.com[lane], c);
        c[0] += X.p[0]; c[1] += X.p[1]; c[2] += X.p[2];
        const float* in = M.inertia[lane];
        float Il[9] = { in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2] }, T[9], Iw[9], Rt[9];
        matmul3(X.R, Il, T);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rt[3 * i + j] = X.R[3 * j + i];
        matmul3(T, Rt, Iw);
        float cc = dot3(c, c);
        I[0] = Iw[0] + m * (cc - c[0] * c[0]); I[1] = Iw[4] + m * (cc - c[1] * c[1]); I[2] = Iw[8] + m * (cc - c[2] * c[2]);
        I[3] = Iw[1] - m * c[0] * c[1]; I[4] = Iw[2] - m * c[0] * c[2]; I[5] = Iw[5] - m * c[1] * c[2];
        h[0] = m * c[0]; h[1] = m * c[1]; h[2] = m * c[2];
    }
    // composite inertia of each subtree
    float mc = subtree_sum(m, lane, se), hc[3], Ic[6];
#pragma unroll
    for (int i = 0; i < 3; i++) hc[i] = subtree_sum(h[i], lane, se);
#pragma unroll
    for (int i = 0; i < 6; i++) Ic[i] = subtree_sum(I[i], lane, se);

    // ---- H (composite rigid body algorithm) ----
    float* H = W.u.s.H;
    for (int i = lane; i < HS * HS; i += 32) H[i] = 0.f;
    __syncwarp();
    if (isj) {
        // F = Ic * S : n = Ibar a + h x l ; f = m l - h x a
        float F[6], t[3];
        F[0] = Ic[0] * S[0] + Ic[3] * S[1] + Ic[4] * S[2];
        F[1] = Ic[3] * S[0] + Ic[1] * S[1] + Ic[5] * S[2];
        F[2] = Ic[4] * S[0] + Ic[5] * S[1] + Ic[2] * S[2];
        cross3(hc, S + 3, t); F[0] += t[0]; F[1] += t[1]; F[2] += t[2];
        cross3(hc, S, t);
        F[3] = mc * S[3] - t[0]; F[4] = mc * S[4] - t[1]; F[5] = mc * S[5] - t[2];
        int d = 5 + lane;
        float hdd = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) hdd += S[i] * F[i];
        H[d * HS + d] = hdd;
        int j = lane;
        // parents: chain neighbours except at the three branch points
        for (int it = 0; it < 8; it++) {
            j = (j == 9) ? 3 : ((j == 14 || j == 18 || j == 1) ? 0 : j - 1);
            if (j < 1) break;
            const float4* Sj = reinterpret_cast<const float4*>(W.S[j]);
            const float4 s0 = Sj[0], s1 = Sj[1];
            float v = s0.x * F[0] + s0.y * F[1] + s0.z * F[2] + s0.w * F[3] + s1.x * F[4] + s1.y * F[5];
            H[d * HS + 5 + j] = v; H[(5 + j) * HS + d] = v;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { H[i * HS + d] = F[i]; H[d * HS + i] = F[i]; }
    }
    if (lane == 0) {
        H[0 * HS + 0] = Ic[0]; H[1 * HS + 1] = Ic[1]; H[2 * HS + 2] = Ic[2];
        H[0 * HS + 1] = H[1 * HS + 0] = Ic[3]; H[0 * HS + 2] = H[2 * HS + 0] = Ic[4]; H[1 * HS + 2] = H[2 * HS + 1] = Ic[5];
        H[3 * HS + 3] = H[4 * HS + 4] = H[5 * HS + 5] = mc;
        // angular rows x linear cols = [h]x ; linear rows x angular cols = -[h]x
        H[0 * HS + 4] = -hc[2]; H[0 * HS + 5] = hc[1]; H[1 * HS + 3] = hc[2]; H[1 * HS + 5] = -hc[0]; H[2 * HS + 3] = -hc[1]; H[2 * HS + 4] = hc[0];
        H[4 * HS + 0] = -hc[2]; H[5 * HS + 0] = hc[1]; H[3 * HS + 1] = hc[2]; H[5 * HS + 1] = -hc[0]; H[3 * HS + 2] = -hc[1]; H[4 * HS + 2] = hc[0];
    }
    __syncwarp();

    // ---- bias forces (recursive Newton-Euler with zero generalised acceleration, gravity as base acceleration) ----
    float qd_b = shfl(v_dof, (lane + 5) & 31);      // hinge rate of body `lane`
    float wb[3], vb[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { wb[i] = shfl(v_dof, i); vb[i] = shfl(v_dof, 3 + i); }
    float vel[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float contrib = isj ? S[i] * qd_b : 0.f;
        if (lane == 0) contrib = i < 3 ? wb[i] : vb[i - 3];
        vel[i] = tree_prefix(contrib, lane, ss);
    }
    float acc6[6];
    {
        float t6[6] = { 0, 0, 0, 0, 0, 0 };
        if (isj) {
            // (v x_m S) qd : angular w x a ; linear w x l + v x a
            float c1[3], c2[3], c3[3];
            cross3(vel, S, c1); cross3(vel, S + 3, c2); cross3(vel + 3, S, c3);
#pragma unroll
            for (int i = 0; i < 3; i++) { t6[i] = c1[i] * qd_b; t6[3 + i] = (c2[i] + c3[i]) * qd_b; }
        }
        if (lane == 0) {
            float c[3]; cross3(wb, vb, c);
            t6[3] = -c[0]; t6[4] = -c[1]; t6[5] = -c[2] + M.gravity;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) acc6[i] = tree_prefix(t6[i], lane, ss);
    }
    float f6[6];
    {
        // f = I a + v x* (I v) with the body's own inertia
        float Iv[6], Ia[6], t[3];
        auto apply = [&](const float* x, float* y) {
            y[0] = I[0] * x[0] + I[3] * x[1] + I[4] * x[2];
            y[1] = I[3] * x[0] + I[1] * x[1] + I[5] * x[2];
            y[2] = I[4] * x[0] + I[5] * x[1] + I[2] * x[2];
            cross3(h, x + 3, t); y[0] += t[0]; y[1] += t[1]; y[2] += t[2];
            cross3(h, x, t);
            y[3] = m * x[3] - t[0]; y[4] = m * x[4] - t[1]; y[5] = m * x[5] - t[2];
        };
        apply(vel, Iv); apply(acc6, Ia);
        float c1[3], c2[3], c3[3];
        cross3(vel, Iv, c1); cross3(vel + 3, Iv + 3, c2); cross3(vel, Iv + 3, c3);
#pragma unroll
        for (int i = 0; i < 3; i++) { f6[i] = Ia[i] + c1[i] + c2[i]; f6[3 + i] = Ia[3 + i] + c3[i]; }
        if (!isb) {
#pragma unroll
            for (int i = 0; i < 6; i++) f6[i] = 0.f;
        }
    }
    float Fs[6];
#pragma unroll
    for (int i = 0; i < 6; i++) Fs[i] = subtree_sum(f6[i], lane, se);
    float Cb = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) Cb += S[i] * Fs[i];
    // move to lane = dof: base rows from lane 0's Fs, hinge rows from lane dof-5
    float rhs = shfl(Cb, (lane + 27) & 31);       // lane d reads body d-5
#pragma unroll
    for (int i = 0; i < 6; i++) { float t = shfl(Fs[i], 0); if (lane == i) rhs = t; }
    float Cdof = rhs;
    rhs = tau_dof - rhs;
    if (lane >= ND) rhs = 0.f;

    // ---- LDL^T, lane i owns row i ----
    float Hr[HS];
    {
        const float4* row = reinterpret_cast<const float4*>(H + (lane < ND ? lane : ND) * HS);   // row ND is all zero
#pragma unroll
        for (int j = 0; j < HS / 4; j++) { float4 v = row[j]; Hr[4 * j] = v.x; Hr[4 * j + 1] = v.y; Hr[4 * j + 2] = v.z; Hr[4 * j + 3] = v.w; }
#pragma unroll
        for (int j = 0; j < ND; j++) if (lane >= ND && j == lane) Hr[j] = 1.f;
    }
    if (DIAG && dbg) {
        if (lane < ND) {
#pragma unroll
            for (int j = 0; j < ND; j++) dbg[DBG_H + lane * ND + j] = Hr[j];
            dbg[DBG_C + lane] = Cdof;
        }
    }
    float myDinv = 1.f;
#if SS_SPARSE_LDL
    ldl_steps<ND - 1>(Hr, myDinv, lane);
#else
    ldl_steps<0>(Hr, myDinv, lane);
#endif
#if SS_SPARSE_LDL
    // rows go to shared memory with the scratch part (columns <= lane) zeroed: the U^T solves below then need no
    // `lane > k` predicate and unroll to SHFL + LDS + FFMA per column
#pragma unroll
    for (int j = 0; j < ND; j++) if (j <= lane) Hr[j] = 0.f;
#endif
    if (lane < ND) {
        float4* row = reinterpret_cast<float4*>(H + lane * HS);
#pragma unroll
        for (int j = 0; j < HS / 4; j++) row[j] = make_float4(Hr[4 * j], Hr[4 * j + 1], Hr[4 * j + 2], Hr[4 * j + 3]);
    }
    W.dinv[lane] = myDinv;
    __syncwarp();
    // solve H acc = rhs
    float y = rhs;
#if SS_SPARSE_LDL
    // U y' = rhs (k descending; Hr[k] = U[lane][k], structurally zero where lane is no ancestor of k), D, U^T x = y''
#pragma unroll
    for (int k = ND - 1; k >= 1; k--) { float yk = shfl(y, k); if (lane < k) y -= Hr[k] * yk; }
    y *= myDinv;
#pragma unroll
    for (int k = 0; k < ND - 1; k++) { const float xk = shfl(y, k); y = fmaf(-H[k * HS + lane], xk, y); }     // (lanes >= ND: discarded below)
#else
#pragma unroll
    for (int k = 0; k < ND - 1; k++) { float yk = shfl(y, k); if (lane > k && lane < ND) y -= Hr[k] * yk; }
    y *= myDinv;
#pragma unroll 1
    for (int k = ND - 1; k >= 1; k--) { float xk = shfl(y, k); if (lane < k) y -= H[k * HS + lane] * xk; }
#endif
    return lane < ND ? y : 0.f;
}

#if SS_SPARSE_LDL
// one row of the back-substitution z = U^-1 b (lane = constraint row; U rows are broadcast LDS.128 reads): quads of row I
// that hold no descendant of dof I are not even loaded
template <int I, int Q>
__device__ __forceinline__ void sparse_urow(const float4* Ui, const float (&z)[HS], float& acc) {
    if constexpr (4 * Q < ND) {
        constexpr bool a0 = dof_anc(I, 4 * Q), a1 = dof_anc(I, 4 * Q + 1), a2 = dof_anc(I, 4 * Q + 2), a3 = dof_anc(I, 4 * Q + 3);
        if constexpr (a0 || a1 || a2 || a3) {
            const float4 u = Ui[Q];
            if constexpr (a0) acc = fmaf(-u.x, z[4 * Q], acc);
            if constexpr (a1) acc = fmaf(-u.y, z[4 * Q + 1], acc);
            if constexpr (a2) acc = fmaf(-u.z, z[4 * Q + 2], acc);
            if constexpr (a3) acc = fmaf(-u.w, z[4 * Q + 3], acc);
        }
        sparse_urow<I, Q + 1>(Ui, z, acc);
    }
}
template <int I>
__device__ __forceinline__ void sparse_usolve(const float* U, float (&z)[HS]) {
    if constexpr (I >= 0) {
        float acc = z[I];
        sparse_urow<I, (I + 1) / 4>(reinterpret_cast<const float4*>(U + I * HS), z, acc);
        z[I] = acc;
        sparse_usolve<I - 1>(U, z);
    }
}
#endif

// back-substitution L^T x = u (lane = dof)
__device__ __forceinline__ float solve_lt(const WarpSmem& W, float u, int lane) {
    const float* H = W.u.s.H;
#if SS_SPARSE_LDL
#pragma unroll
    for (int k = 0; k < ND - 1; k++) { const float xk = shfl(u, k); u = fmaf(-H[k * HS + lane], xk, u); }   // U^T x = u (scratch columns are zero)
#else
#pragma unroll 1
    for (int k = ND - 1; k >= 1; k--) { float xk = shfl(u, k); if (lane < k) u -= H[k * HS + lane] * xk; }
#endif
    return u;
}

__device__ __forceinline__ void plane_space(const float* n, float* p, float* q) {
    if (fabsf(n[2]) > 0.70710678f) {
        float a = n[1] * n[1] + n[2] * n[2], k = rsqrtf(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        float a = n[0] * n[0] + n[1] * n[1], k = rsqrtf(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

// ---------------------------------------------------------------------------------------------
// position integration (btMultiBody::stepPositionsMultiDof): exponential map for the base orientation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void integrate_warp(const DevModel& M, float* quat, float& q_dof, float vnew, float* dpos, int lane) {
    const float hdt = M.dt;
    float w[3], vl[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { w[i] = shfl(vnew, i); vl[i] = shfl(vnew, 3 + i); }
#pragma unroll
    for (int i = 0; i < 3; i++) dpos[i] += vl[i] * hdt;
    float ang = fsqrt(dot3(w, w));
    if (ang * hdt > 0.78539816f) ang = fdiv(0.78539816f, hdt);
    float f = ang < 0.001f ? 0.5f * hdt - hdt * hdt * hdt * 0.020833333333f * ang * ang : fdiv(sinf(0.5f * ang * hdt), ang);
    float dq[4] = { w[0] * f, w[1] * f, w[2] * f, cosf(0.5f * ang * hdt) }, rq[4];
    rq[0] = dq[3] * quat[0] + dq[0] * quat[3] + dq[1] * quat[2] - dq[2] * quat[1];
    rq[1] = dq[3] * quat[1] - dq[0] * quat[2] + dq[1] * quat[3] + dq[2] * quat[0];
    rq[2] = dq[3] * quat[2] + dq[0] * quat[1] - dq[1] * quat[0] + dq[2] * quat[3];
    rq[3] = dq[3] * quat[3] - dq[0] * quat[0] - dq[1] * quat[1] - dq[2] * quat[2];
    float nn = rsqrtf(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) quat[i] = rq[i] * nn;
    if (lane >= 6 && lane < ND) q_dof += vnew * hdt;
}

// ---------------------------------------------------------------------------------------------
// collision detection.  Candidates are evaluated 32 at a time (lane = candidate):
//   stone passes  (geom, plank slot, segment end) from M.scand -- feet first, then limbs, then spheres; id geom*8+slot*2+end
//   pair passes   one closest-point pair per self-collision geom pair (collision_mode 2); id SS_NGEOM*8 + pair
// and merged into a pool of the MAXC deepest (ties by id) kept in W.con.  The pool is finally ordered by id and the cached
// normal impulse of each contact is looked up in the previous substep's list (`warm`: lanes 0..7 ids, 8..15 impulses).
// Passes whose candidates are all out of reach (bounding test) are skipped.  Returns the number of contacts.
// ---------------------------------------------------------------------------------------------
struct Cand { bool active; float depth; int id; float n[3], x[3], xb[3]; int ba, bb, code; };    // code: geometry branch (parity probe, see ss_substep_events)

__device__ __forceinline__ void con_write(float* c, const Cand& k) {
    c[0] = k.n[0]; c[1] = k.n[1]; c[2] = k.n[2]; c[3] = k.x[0]; c[4] = k.x[1]; c[5] = k.x[2];
    c[6] = k.depth; c[7] = (float)k.ba; c[8] = (float)k.id; c[9] = 0.f;
    c[10] = k.xb[0]; c[11] = k.xb[1]; c[12] = k.xb[2]; c[13] = (float)k.bb; c[14] = (float)k.code;
}

__device__ __noinline__ int pool_merge_slow(WarpSmem* Wp, int npool, bool active, float depth, int id, const float* rec, int lane) {
    // more than MAXC candidates: keep the MAXC deepest of (pool U new), ties by id
    WarpSmem& W = *Wp;
    const unsigned amask = __ballot_sync(FULL, active);
    float pd = 1e30f; int pid = 0x7fffffff;
    if (lane < npool) { pd = W.con[lane][6]; pid = (int)W.con[lane][8]; }
    int rank_new = 0, rank_pool = 0;
    for (int m = 0; m < 32; m++) {
        float od = shfl(depth, m); int oid = shfl(id, m);
        if ((amask >> m) & 1) {
            if (od < depth || (od == depth && oid < id)) rank_new++;
            if (od < pd || (od == pd && oid < pid)) rank_pool++;
        }
    }
    for (int k = 0; k < npool; k++) {
        float od = shfl(pd, k); int oid = shfl(pid, k);
        if (od < depth || (od == depth && oid < id)) rank_new++;
        if (od < pd || (od == pd && oid < pid)) rank_pool++;
    }
    const bool keep_new = active && rank_new < MAXC, keep_pool = lane < npool && rank_pool < MAXC;
    const unsigned pm = __ballot_sync(FULL, keep_pool), nm = __ballot_sync(FULL, keep_new);
    const unsigned lt = (1u << lane) - 1u;
    float tmp[16];
    if (keep_pool) {
#pragma unroll
        for (int i = 0; i < 16; i++) tmp[i] = W.con[lane][i];
    }
    __syncwarp();
    if (keep_pool) {
        float* c = W.con[__popc(pm & lt)];
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = tmp[i];
    }
    if (keep_new) {
        float* c = W.con[__popc(pm) + __popc(nm & lt)];
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = rec[i];
    }
    __syncwarp();
    return __popc(pm) + __popc(nm);
}

__device__ __forceinline__ void pool_add(WarpSmem& W, int& npool, const Cand& k, int lane) {
    const unsigned amask = __ballot_sync(FULL, k.active);
    const int na = __popc(amask);
    if (na == 0) return;
    if (npool + na <= MAXC) {
        if (k.active) con_write(W.con[npool + __popc(amask & ((1u << lane) - 1u))], k);
        npool += na;
        __syncwarp();
        return;
    }
    float rec[16];
    con_write(rec, k); rec[15] = 0.f;
    npool = pool_merge_slow(&W, npool, k.active, k.depth, k.id, rec, lane);
}

// capsule segment (a -> b, plank-local coordinates) of radius r against the plank box: one end of the part of the segment
// above the plank's footprint
__device__ __forceinline__ void stone_candidate(const DevModel& M, const float* st, const float* g, int endp, float breaking,
                                                const float* dpos, Cand& k) {
    float p0[3], p1[3], a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { p0[i] = g[i] + dpos[i] - st[i]; p1[i] = g[3 + i] + dpos[i] - st[i]; }
    matTvec3(st + 3, p0, a); matTvec3(st + 3, p1, b);
    float t0 = 0.f, t1 = 1.f; bool empty = false;
#pragma unroll
    for (int ax = 0; ax < 2; ax++) {
        float d = b[ax] - a[ax], hh = M.stone_half[ax];
        if (fabsf(d) < 1e-12f) { if (fabsf(a[ax]) > hh) empty = true; }
        else {
            const float id = frcp(d);
            float ta = (-hh - a[ax]) * id, tb = (hh - a[ax]) * id;
            if (ta > tb) { float s = ta; ta = tb; tb = s; }
            t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
        }
    }
    if (empty || t0 > t1) { t0 = 0.f; t1 = 1.f; k.code |= 2; }
    if (endp == 1 && t1 - t0 < 1e-6f) return;
    float tt = endp ? t1 : t0, s[3], q[3], nl[3], dist;
    bool inside = true;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s[i] = a[i] + tt * (b[i] - a[i]);
        float hh = M.stone_half[i];
        q[i] = fminf(fmaxf(s[i], -hh), hh);
        if (q[i] != s[i]) inside = false;
    }
    if (!inside) {
        float d[3] = { s[0] - q[0], s[1] - q[1], s[2] - q[2] };
        dist = fsqrt(dot3(d, d));
        float inv = frcp(dist);
        nl[0] = d[0] * inv; nl[1] = d[1] * inv; nl[2] = d[2] * inv;
    } else {
        int ax = 0; float best = 1e30f;
#pragma unroll
        for (int i = 0; i < 3; i++) { float mm = M.stone_half[i] - fabsf(s[i]); if (mm < best) { best = mm; ax = i; } }
        nl[0] = nl[1] = nl[2] = 0.f;
        float sg = (ax == 0 ? s[0] : (ax == 1 ? s[1] : s[2])) >= 0.f ? 1.f : -1.f;
        if (ax == 0) nl[0] = sg; else if (ax == 1) nl[1] = sg; else nl[2] = sg;
        dist = -best;
        k.code |= 4 | (ax << 3);
    }
    const float r = g[6];
    k.depth = dist - r;
    if (k.depth < breaking) {
        k.active = true;
        matvec3(st + 3, nl, k.n);
        float sw[3]; matvec3(st + 3, s, sw);
        // contact point on the robot, relative to the CURRENT base origin
#pragma unroll
        for (int i = 0; i < 3; i++) { k.x[i] = sw[i] + st[i] - dpos[i] - r * k.n[i]; k.xb[i] = k.x[i]; }
    }
}

// closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9), as the oracle's segment_segment
__device__ __forceinline__ void segment_segment(const float* p1, const float* q1, const float* p2, const float* q2, float* c1, float* c2) {
    float d1[3], d2[3], r[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { d1[i] = q1[i] - p1[i]; d2[i] = q2[i] - p2[i]; r[i] = p1[i] - p2[i]; }
    const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    const float EPS = 1e-12f;
    float s, t;
    auto clamp01 = [](float x) { return fminf(fmaxf(x, 0.f), 1.f); };
    if (a <= EPS && e <= EPS) { s = t = 0.f; }
    else if (a <= EPS) { s = 0.f; t = clamp01(fdiv(f, e)); }
    else {
        const float c = dot3(d1, r);
        if (e <= EPS) { t = 0.f; s = clamp01(fdiv(-c, a)); }
        else {
            const float b = dot3(d1, d2), denom = a * e - b * b;
            s = denom > EPS * a * e ? clamp01(fdiv(b * f - c * e, denom)) : 0.f;
            t = fdiv(b * s + f, e);
            if (t < 0.f) { t = 0.f; s = clamp01(fdiv(-c, a)); }
            else if (t > 1.f) { t = 1.f; s = clamp01(fdiv(b - c, a)); }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) { c1[i] = p1[i] + d1[i] * s; c2[i] = p2[i] + d2[i] * t; }
}

__device__ __forceinline__ int collide_warp(const DevModel& M, WarpSmem& W, const float* dpos, float warm, int lane, int& cflags) {
    // geom segments in the common frame
    if (lane < M.n_geom) {
        const float* P = W.pose[M.geom_body[lane]];
        float t0[3], t1[3];
        matvec3(P, M.geom_p0[lane], t0); matvec3(P, M.geom_p1[lane], t1);
        float* g = W.gw[lane];
#pragma unroll
        for (int i = 0; i < 3; i++) { g[i] = P[9 + i] + t0[i]; g[3 + i] = P[9 + i] + t1[i]; }
        g[6] = M.geom_radius[lane]; g[7] = 0.f;
    }
    __syncwarp();
    int npool = 0;
    // ---- stones ----
    for (int base = 0; base < M.n_scand; base += 32) {
        Cand k; k.active = false; k.depth = 0.f; k.id = 0; k.ba = 0; k.bb = -1; k.code = 0;
        const int ci = base + lane;
        bool near = false;
        int g = 0, slot = 0, endp = 0;
        if (ci < M.n_scand) {
            const int code = M.scand[ci];
            g = code & 31; slot = (code >> 5) & 3; endp = (code >> 7) & 1;
            const float* st = W.stone[slot];
            if (st[12] != 0.f) {
                // bounding test: distance of the segment midpoint to the plank box vs half length + radius
                const float* gg = W.gw[g];
                float d[3], dl[3];
#pragma unroll
                for (int i = 0; i < 3; i++) d[i] = 0.5f * (gg[i] + gg[3 + i]) + dpos[i] - st[i];
                matTvec3(st + 3, d, dl);
#pragma unroll
                for (int i = 0; i < 3; i++) dl[i] = fmaxf(fabsf(dl[i]) - M.stone_half[i], 0.f);
                float reach = M.geom_half[g] + gg[6] + M.stone_breaking + 1e-4f;
                near = dot3(dl, dl) < reach * reach;
            }
        }
        if (base > 0 && !__any_sync(FULL, near)) continue;
        if (near) {
            stone_candidate(M, W.stone[slot], W.gw[g], endp, fminf(M.geom_breaking[g], M.stone_breaking), dpos, k);
            k.id = g * 8 + slot * 2 + endp; k.ba = M.geom_body[g];
        }
        pool_add(W, npool, k, lane);
    }
    // ---- the robot against itself ----
    if (M.collision_mode >= 2) {
        for (int base = 0; base < M.n_pair; base += 32) {
            Cand k; k.active = false; k.depth = 0.f; k.id = 0; k.ba = 0; k.bb = -1; k.code = 0;
            const int pi = base + lane;
            bool near = false;
            int ga = 0, gb = 0;
            if (pi < M.n_pair) {
                ga = M.pair_a[pi]; gb = M.pair_b[pi];
                const float *A = W.gw[ga], *B = W.gw[gb];
                float d[3];
#pragma unroll
                for (int i = 0; i < 3; i++) d[i] = 0.5f * ((A[i] + A[3 + i]) - (B[i] + B[3 + i]));
                float reach = M.geom_half[ga] + A[6] + M.geom_half[gb] + B[6] + 0.01f;
                near = dot3(d, d) < reach * reach;
            }
            if (!__any_sync(FULL, near)) continue;
            if (near) {
                const float *A = W.gw[ga], *B = W.gw[gb];
                float ca[3], cb[3], d[3];
                segment_segment(A, A + 3, B, B + 3, ca, cb);
#pragma unroll
                for (int i = 0; i < 3; i++) d[i] = ca[i] - cb[i];
                const float dist = fsqrt(dot3(d, d)), ra = A[6], rb = B[6];
                k.depth = dist - ra - rb;
                if (k.depth < fminf(M.geom_breaking[ga], M.geom_breaking[gb])) {
                    k.active = true;
                    if (dist > 1e-9f) { float inv = frcp(dist); k.n[0] = d[0] * inv; k.n[1] = d[1] * inv; k.n[2] = d[2] * inv; }
                    else { k.n[0] = k.n[1] = 0.f; k.n[2] = 1.f; }
#pragma unroll
                    for (int i = 0; i < 3; i++) { k.x[i] = ca[i] - ra * k.n[i]; k.xb[i] = cb[i] + rb * k.n[i]; }
                    k.id = SS_NGEOM * 8 + pi; k.ba = M.geom_body[ga]; k.bb = M.geom_body[gb];
                }
            }
            pool_add(W, npool, k, lane);
        }
    }
    cflags = 0;
    if (npool == 0) return 0;
    // ---- order by id, look the cached impulses up, foot flags ----
    {
        float rec[16]; int myid = 0x7fffffff;
        if (lane < npool) {
#pragma unroll
            for (int i = 0; i < 16; i++) rec[i] = W.con[lane][i];
            myid = (int)rec[8];
        }
        int rank = 0; float lam0 = 0.f;
        for (int j = 0; j < npool; j++) { int oid = shfl(myid, j); if (oid < myid) rank++; }
#pragma unroll
        for (int w = 0; w < MAXC; w++) {
            float wid = shfl(warm, w), wl = shfl(warm, MAXC + w);
            if ((int)wid == myid) lam0 = wl;
        }
        __syncwarp();
        bool f0 = false, f1 = false;
        if (lane < npool) {
            rec[9] = lam0;
            float* c = W.con[rank];
#pragma unroll
            for (int i = 0; i < 15; i++) c[i] = rec[i];
            const int ba = (int)rec[7], bb = (int)rec[13];
            f0 = ba == M.foot_body[0] || bb == M.foot_body[0];
            f1 = ba == M.foot_body[1] || bb == M.foot_body[1];
        }
        if (__any_sync(FULL, f0)) cflags |= 1;
        if (__any_sync(FULL, f1)) cflags |= 2;
        __syncwarp();
    }
    return npool;
}

// ---------------------------------------------------------------------------------------------
// Projected Gauss-Seidel in constraint space (btMultiBodyConstraintSolver::solveSingleIteration order: limit rows in
// alternating direction, then all normals, then all friction rows), lane = row, Delassus row in registers.
//   Every lane evaluates its own row update each row step (branch-free); the owner of row RR publishes its impulse
//   change with one shuffle and all lanes fold it into their accumulator c = lam (1 - cfm) + rhs - (J dv) / D.
//   Contact normals (lo = 0, hi = "infinity") keep e = c - lam instead of c: the impulse change is max(e, -lam), one
//   FMNMX and no subtraction on the dependent chain; their diagonal coefficient carries +cfm instead of -(1 - cfm).
//   Friction bounds depend on the normal impulse, which only changes in the normal block: one shuffle per iteration.
//   NC = number of contacts (exact): rows 0..NC-1 normals, MAXC + 2c, MAXC + 2c + 1 the friction rows of contact c.
// ---------------------------------------------------------------------------------------------
#define SS_ROW_STEP(RR)                                                        \
        {                                                                      \
            float sum_ = fminf(fmaxf(c, lo_c), hi_c);                          \
            float dl_ = shfl(sum_ - lam, (RR));                                \
            lam = (lane == (RR)) ? sum_ : lam;                                 \
            c = fmaf(-Ar[(RR)], dl_, c);                                       \
        }
#define SS_NORMAL_STEP(RR)                                                     \
        {                                                                      \
            float d_ = fmaxf(c, -lam);                                         \
            float dl_ = shfl(d_, (RR));                                        \
            lam = (lane == (RR)) ? lam + d_ : lam;                             \
            c = fmaf(-Ar[(RR)], dl_, c);                                       \
        }
template <bool DIAG, int NC>
__device__ __forceinline__ void pgs_solve(const float (&Ar)[32], float& c, float& lam, float& lo_c, float& hi_c, const int nl,
                                          const int n_iters, const bool isfric, const int nsrc, const float mu, const int lane,
                                          unsigned* push_hist) {
    // warm start: fold the cached normal impulses in (dv = sum_s W_s lam_s)
#pragma unroll
    for (int k = 0; k < NC; k++) { const float l0 = shfl(lam, k); c = fmaf(-Ar[k], l0, c); }
    for (int it = 0; it < n_iters; it++) {
        if (it & 1) {
#pragma unroll
            for (int k = 0; k < MAXL; k += 2)
                if (nl > k) { SS_ROW_STEP(3 * MAXC + k); SS_ROW_STEP(3 * MAXC + k + 1); }
        } else {
#pragma unroll
            for (int k = MAXL - 2; k >= 0; k -= 2)
                if (nl > k) { SS_ROW_STEP(3 * MAXC + k + 1); SS_ROW_STEP(3 * MAXC + k); }
        }
        if constexpr (NC > 0) {
#pragma unroll
            for (int k = 0; k < NC; k++) SS_NORMAL_STEP(k);
            const float ln = shfl(lam, nsrc);
            if (isfric) { const bool live = ln > 0.f; hi_c = live ? mu * ln : lam; lo_c = live ? -hi_c : lam; }
            // friction rows of a contact that does not push are no-ops (bounds pinned to the current impulse; Bullet skips
            // them too: `if (totalImpulse > 0)` in solveSingleIteration): skipped, warp-uniformly
            const unsigned pushing = __ballot_sync(FULL, lam > 0.f) & ((1u << MAXC) - 1u);
            if (DIAG && push_hist && it < 31) *push_hist |= ((pushing >> (lane & (MAXC - 1))) & 1u) << it;   // lane = contact (parity probe)
#pragma unroll
            for (int k = 0; k < NC; k++)
                if ((pushing >> k) & 1u) { SS_ROW_STEP(MAXC + 2 * k); SS_ROW_STEP(MAXC + 2 * k + 1); }
        }
    }
}
#undef SS_ROW_STEP
#undef SS_NORMAL_STEP

// ---------------------------------------------------------------------------------------------
// one physics substep.  State in registers: quat (all lanes), q_dof / v_dof (lane = dof), dpos =
// displacement of the base since the start of the control step (all lanes), warm (lane = candidate slot).
// Returns the per-foot contact flags of this substep's collision pass (bit f).
// ---------------------------------------------------------------------------------------------
template <bool DIAG>
__device__ __forceinline__ int substep_warp(const DevModel& M, WarpSmem& W, float* quat, float& q_dof, float& v_dof,
                                            float tau_dof, float* dpos, float& warm, int lane, float* dbg, bool* clamp_hit = nullptr,
                                            unsigned* push_hist = nullptr) {
    const float hdt = M.dt;
    float acc = dynamics_warp<DIAG>(M, W, quat, q_dof, v_dof, tau_dof, lane, dbg);
    float vs = v_dof + hdt * acc;
    if (DIAG && clamp_hit) *clamp_hit = __any_sync(FULL, lane < ND && fabsf(vs) > M.max_coord_vel);
    vs = fminf(fmaxf(vs, -M.max_coord_vel), M.max_coord_vel);
    if (lane >= ND) vs = 0.f;

    // ---- collision (see collide_warp): <= MAXC contacts in W.con, ordered by candidate id, warm impulses looked up ----
    int cflags = 0;
    const int nc = collide_warp(M, W, dpos, warm, lane, cflags);
    bool anyself = false;
    if (lane < nc) anyself = W.con[lane][13] >= 0.f;
    anyself = __any_sync(FULL, anyself);
    if (DIAG && dbg) {
        float* d = dbg + DBG_CAND + lane * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = 0.f;
        __syncwarp();
        if (lane < nc) {
            const float* c = W.con[lane]; float* o = dbg + DBG_CAND + lane * 16;
            o[0] = 1.f; o[1] = c[6]; o[2] = c[0]; o[3] = c[1]; o[4] = c[2]; o[5] = c[3]; o[6] = c[4]; o[7] = c[5];
            o[8] = c[8]; o[9] = c[7]; o[10] = c[13]; o[11] = c[10]; o[12] = c[11]; o[13] = c[12];
        }
    }

    // ---- violated hinge limits (lane = dof) ----
    int lviol = 0; float lpen = 0.f;
    if (lane >= 6 && lane < ND) {
        float plo = q_dof - M.q_lo[lane - 6], phi = M.q_hi[lane - 6] - q_dof;
        if (plo <= 0.f) { lviol = 1; lpen = plo; } else if (phi <= 0.f) { lviol = 2; lpen = phi; }
    }
    unsigned lmask = __ballot_sync(FULL, lviol != 0);
    const int lidx = __popc(lmask & ((1u << lane) - 1u));
    const int nl = min(__popc(lmask), MAXL);
    const int nr = nl + 3 * nc;

    float dv = 0.f;          // constraint velocity change (lane = dof)
    float newwarm = lane < MAXC ? -1.f : 0.f;
#if SS_SYNC >= 2
    __syncthreads();         // phase alignment: row build
#endif
    if (nr > 0) {
        // publish row definitions
        if (lviol && lidx < MAXL) { W.lim[lidx][0] = (float)(lane - 6); W.lim[lidx][1] = lviol == 1 ? 1.f : -1.f; W.lim[lidx][2] = lpen; }
        if (lane < ND) W.vec[lane] = vs;
        __syncwarp();

        // ---- lane = row slot: Jacobian row, z = L^-1 J^T ----
        // Fixed slots (compile-time row indices in the solver): 0..7 contact normals, 8..23 friction (two per contact,
        // contact-major as Bullet orders them), 24..31 violated hinge limits.  Inactive slots carry all-zero rows.
        // A link-link contact acts on two bodies of the tree: J = J_a - J_b.  When the env has one (warp-uniform), rows
        // go through the substitution twice (a side, b side): Bullet's row denominator is J_a H^-1 J_a^T + J_b H^-1 J_b^T
        // (no cross terms), which needs both halves separately.
        const int r = lane;
        float z[HS];
        int rtype = -1;      // 0 limit, 1 normal, 2/3 friction
        int nidx = -1; float pen = 0.f, lam = 0.f;
        int ccon = -1, hj = -1, maskA = 0, maskB = 0; bool selfrow = false;
        float dir[3] = { 0.f, 0.f, 0.f }, xa[3] = { 0.f, 0.f, 0.f }, xb[3] = { 0.f, 0.f, 0.f }, ldir = 0.f;
        {
            int a = 0;
            if (r < MAXC) { if (r < nc) ccon = r; }
            else if (r < 3 * MAXC) { int cc = (r - MAXC) >> 1; if (cc < nc) { ccon = cc; a = 1 + ((r - MAXC) & 1); nidx = cc; } }
            else if (r - 3 * MAXC < nl) {
                rtype = 0;
                const float* ld = W.lim[r - 3 * MAXC];
                hj = (int)ld[0]; ldir = ld[1]; pen = ld[2];
            }
            if (ccon >= 0) {
                rtype = 1 + a;
                const float* cd = W.con[ccon];
                float n[3] = { cd[0], cd[1], cd[2] }, t1[3], t2[3];
                plane_space(n, t1, t2);
#pragma unroll
                for (int i = 0; i < 3; i++) { dir[i] = a == 0 ? n[i] : (a == 1 ? t1[i] : t2[i]); xa[i] = cd[3 + i]; xb[i] = cd[10 + i]; }
                pen = cd[6]; if (a == 0) lam = cd[9] * M.warmstart;
                maskA = M.anc_mask[(int)cd[7]];
                const int bodyB = (int)cd[13];
                selfrow = bodyB >= 0;
                if (selfrow) maskB = M.anc_mask[bodyB];
            }
        }
        float rel = 0.f, dsides = 0.f;
        const int npass = anyself ? 2 : 1;
        float* Z = W.u.s.Z;
#pragma unroll 1
        for (int pass = 0; pass < npass; pass++) {
#pragma unroll
            for (int k = 0; k < HS; k++) z[k] = 0.f;
            if (pass == 0 && rtype == 0) {
#pragma unroll
                for (int k = 6; k < ND; k++) z[k] = (k - 6 == hj) ? ldir : 0.f;
            }
            if (ccon >= 0 && (pass == 0 || selfrow)) {
                float x[3], wr[3];
#pragma unroll
                for (int i = 0; i < 3; i++) x[i] = pass == 0 ? xa[i] : xb[i];
                const int mask = pass == 0 ? maskA : maskB;
                cross3(x, dir, wr);
                z[0] = wr[0]; z[1] = wr[1]; z[2] = wr[2]; z[3] = dir[0]; z[4] = dir[1]; z[5] = dir[2];
#pragma unroll
                for (int i = 1; i < NB; i++) {
                    const float4* Si = reinterpret_cast<const float4*>(W.S[i]);
                    const float4 s0 = Si[0], s1 = Si[1];
                    float v = s0.x * wr[0] + s0.y * wr[1] + s0.z * wr[2] + s0.w * dir[0] + s1.x * dir[1] + s1.y * dir[2];
                    z[5 + i] = ((mask >> i) & 1) ? v : 0.f;
                }
            }
            // relative velocity along the row before the solve: J v*
            float relp = 0.f;
#pragma unroll
            for (int k = 0; k < ND; k++) relp += z[k] * W.vec[k];
            rel += pass == 0 ? relp : -relp;
            // forward substitution with the unit lower factor, row by row (broadcast LDS.128 reads of L)
            const float* L = W.u.s.H;
#if SS_SPARSE_LDL
            sparse_usolve<ND - 2>(L, z);         // z = U^-1 J^T, rows from the leaves to the root, structural zeros skipped
#else
#pragma unroll
            for (int i = 1; i < ND; i++) {
                const float4* Li = reinterpret_cast<const float4*>(L + i * HS);
                float acc = z[i];
#pragma unroll
                for (int j = 0; j < (i + 3) / 4; j++) {
                    float4 l = Li[j];
                    if (4 * j < i) acc = fmaf(-l.x, z[4 * j], acc);
                    if (4 * j + 1 < i) acc = fmaf(-l.y, z[4 * j + 1], acc);
                    if (4 * j + 2 < i) acc = fmaf(-l.z, z[4 * j + 2], acc);
                    if (4 * j + 3 < i) acc = fmaf(-l.w, z[4 * j + 3], acc);
                }
                z[i] = acc;
            }
#endif
            if (npass == 2) {
#pragma unroll
                for (int k = 0; k < ND; k++) dsides = fmaf(z[k] * W.dinv[k], z[k], dsides);
                float4* Zr = reinterpret_cast<float4*>(Z + r * ZS);
                if (pass == 0) {
#pragma unroll
                    for (int j = 0; j < ZS / 4; j++) Zr[j] = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < ZS / 4; j++) {
                        float4 v = Zr[j];
                        z[4 * j] = v.x - z[4 * j]; z[4 * j + 1] = v.y - z[4 * j + 1]; z[4 * j + 2] = v.z - z[4 * j + 2]; z[4 * j + 3] = v.w - z[4 * j + 3];
                    }
                }
            }
        }
        // the row's z goes to shared memory: the Delassus loop below reads it back chunk by chunk (so that z[] and the
        // 32 Delassus entries never hold registers at the same time) and the final velocity update reads it with lane = dof
        float diag = 0.f;
#pragma unroll
        for (int k = 0; k < ND; k++) diag = fmaf(z[k] * W.dinv[k], z[k], diag);
        {
            float4* Zr = reinterpret_cast<float4*>(Z + r * ZS);
#pragma unroll
            for (int j = 0; j < ZS / 4; j++) Zr[j] = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
        }
        __syncwarp();
        // row constants (btMultiBodyConstraintSolver::setupMultiBodyContactConstraint / btMultiBodyJointLimitConstraint).
        // Stones are soft (contactStiffness / contactDamping -> cfm, erp); link-link contacts are hard (cfm 0, default contact ERP).
        float invD = 0.f, rhs = 0.f, cfmt = 0.f, lo = 0.f, hi = 0.f;
        const float inv_h = frcp(hdt);
        if (selfrow && !M.self_cross_terms) diag = dsides;
        if (rtype == 0) {
            invD = frcp(diag);
            rhs = (-pen * M.erp_limit * inv_h - rel) * invD;
            hi = M.limit_max_impulse;
        } else if (rtype >= 1) {
            const float cfm = selfrow ? 0.f : M.cfm_contact, erp = selfrow ? M.self_erp : M.erp_contact;
            invD = frcp(diag + cfm);
            cfmt = cfm * invD;
            if (rtype == 1) {
                float p = pen + M.linear_slop, poserr = 0.f, velerr = -rel;
                if (p > 0.f) velerr -= p * inv_h; else poserr = -p * erp * inv_h;
                rhs = (poserr + velerr) * invD;
                hi = 1e10f;
            } else rhs = -rel * invD;
        }
        // Delassus row, scaled by this row's 1/(D + cfm):  A[r][s] = (z_r D^-1 . z_s) * invD_r ; the solver works on
        // c_r = lam_r (1 - cfm_r) + rhs_r - (J_r dv) invD_r, so the diagonal carries the extra -(1 - cfm_r).
        // The row stays in REGISTERS (every solver index is a compile-time constant).  Slots are visited in groups of
        // four -- normals {0-3, 4-7}, friction {8-11, ...}, limits {24-27, 28-31} -- and groups without a live row are
        // skipped (warp-uniform).  The dof dimension is the OUTER loop, four dofs at a time: one LDS.128 of the lane's own
        // row and of 1/D, then one broadcast LDS.128 + 4 FFMA per live slot.
        unsigned gmask = 0;
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const bool on = g < 2 ? nc > 4 * g : (g < 6 ? nc > 2 * (g - 2) : nl > 4 * (g - 6));
            gmask |= on ? 1u << g : 0u;
        }
        float Ar[32];
#pragma unroll
        for (int k = 0; k < 32; k++) Ar[k] = 0.f;
        {
            const float4* Z4 = reinterpret_cast<const float4*>(Z);
            const float4* D4 = reinterpret_cast<const float4*>(W.dinv);
#pragma unroll DEL_UNROLL
            for (int j = 0; j < ZS / 4; j++) {
                const float4 o = Z4[r * (ZS / 4) + j], d = D4[j];
                const float w0 = o.x * d.x, w1 = o.y * d.y, w2 = o.z * d.z, w3 = o.w * d.w;
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    if ((gmask >> g) & 1u) {
#pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            const float4 v = Z4[(4 * g + jj) * (ZS / 4) + j];
                            float a = Ar[4 * g + jj];
                            a = fmaf(w0, v.x, a); a = fmaf(w1, v.y, a); a = fmaf(w2, v.z, a); a = fmaf(w3, v.w, a);
                            Ar[4 * g + jj] = a;
                        }
                    }
                }
            }
            const float dself = rtype == 1 ? -cfmt : 1.f - cfmt;
#pragma unroll
            for (int k = 0; k < 32; k++) Ar[k] = Ar[k] * invD - ((k == r && rtype >= 0) ? dself : 0.f);
        }
        if (DIAG && dbg) {
            dbg[DBG_ROWS + 0] = (float)nl; dbg[DBG_ROWS + 1] = (float)nc;
            float* d = dbg + DBG_ROWS + 4 + lane * 6;
            d[0] = (float)rtype; d[1] = invD; d[2] = rhs; d[3] = rel; d[4] = diag; d[5] = lam;
        }
        // ---- projected Gauss-Seidel in constraint space (btMultiBodyConstraintSolver::solveSingleIteration order:
        //      limit rows in alternating direction, then all normals, then all friction rows).
        //      Every lane evaluates its own row update each row-step (branch-free); the owner of row rr publishes its
        //      impulse change with one shuffle and all lanes fold it into their accumulator c (5 ops + 1 shuffle, the
        //      Delassus entry is a register).  Friction bounds depend on the normal impulse, which only changes in the
        //      normal block: one shuffle per iteration.
        const bool isfric = nidx >= 0;
        const int nsrc = isfric ? nidx : lane;
        const float mu = selfrow ? M.self_friction : M.friction;
        // friction rows start dead (bounds pinned to the current impulse => no change) until their normal pushes
        float lo_c = isfric ? lam : lo, hi_c = isfric ? lam : hi;
        float c = rhs;
        // warm start + the 20 iterations, in a copy specialised for the EXACT number of contacts (compile-time row
        // indices, no per-row-pair count checks: they were 240 uniform compare + branch pairs per substep, and a contact
        // count that is odd no longer pays a dead normal step and two dead friction steps per iteration)
        switch (nc) {
            case 0: pgs_solve<DIAG, 0>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 1: pgs_solve<DIAG, 1>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 2: pgs_solve<DIAG, 2>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 3: pgs_solve<DIAG, 3>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 4: pgs_solve<DIAG, 4>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 5: pgs_solve<DIAG, 5>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 6: pgs_solve<DIAG, 6>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            case 7: pgs_solve<DIAG, 7>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
            default: pgs_solve<DIAG, 8>(Ar, c, lam, lo_c, hi_c, nl, M.n_iters, isfric, nsrc, mu, lane, push_hist); break;
        }
        if (DIAG && dbg) dbg[DBG_ROWS + 4 + lane * 6 + 5] = lam;
        // ---- velocity update: dv = L^-T D^-1 sum_r z_r lam_r ----
        __syncwarp();
        float u = 0.f;
        auto gather = [&](int s0, int s1) {
            for (int s = s0; s < s1; s++) { float ls = shfl(lam, s); if (lane < ND) u = fmaf(Z[s * ZS + lane], ls, u); }
        };
        gather(0, nc); gather(MAXC, MAXC + 2 * nc); gather(3 * MAXC, 3 * MAXC + nl);
        dv = solve_lt(W, u * W.dinv[lane], lane);
        // warm-start cache for the next substep: lanes 0..7 the contact ids, lanes 8..15 their normal impulses
        float lw = shfl(lam, (lane - MAXC) & (MAXC - 1));
        newwarm = -1.f;
        if (lane < nc) newwarm = W.con[lane][8];
        else if (lane >= MAXC) newwarm = (lane - MAXC < nc) ? lw : 0.f;
        __syncwarp();
    }
    warm = newwarm;
    float vnew = vs + dv;
    if (DIAG && clamp_hit) *clamp_hit = *clamp_hit || __any_sync(FULL, lane < ND && fabsf(vnew) > M.max_coord_vel);
    vnew = fminf(fmaxf(vnew, -M.max_coord_vel), M.max_coord_vel);
    if (lane >= ND) vnew = 0.f;
    v_dof = vnew;

    integrate_warp(M, quat, q_dof, vnew, dpos, lane);
    return cflags;
}

// ---------------------------------------------------------------------------------------------
// MT19937 (numpy legacy RandomState) -- warp-cooperative on the env's state in global memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
__device__ __forceinline__ void mt_twist_warp(uint32_t* mt, int lane) {
    // word k needs old mt[k+1] and mt[(k+397)%624] (old for k < 227, already-new afterwards):
    // three dependency-free sweeps + the last word.  Each sweep reads everything before it writes.
    const int lo[3] = { 0, 227, 454 }, hi[3] = { 227, 454, 623 };
#pragma unroll 1
    for (int ph = 0; ph < 3; ph++) {
        uint32_t tmp[8]; int cnt = 0;
        for (int k = lo[ph] + lane; k < hi[ph]; k += 32) tmp[cnt++] = mt_mix(mt[k], mt[k + 1], mt[(k + 397) % 624]);
        __syncwarp();
        cnt = 0;
        for (int k = lo[ph] + lane; k < hi[ph]; k += 32) mt[k] = tmp[cnt++];
        __syncwarp();
    }
    if (lane == 0) mt[623] = mt_mix(mt[623], mt[0], mt[396]);
    __syncwarp();
}
__device__ __forceinline__ double mt_make_double(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
// draw `count` doubles into out[] (out == nullptr discards them).  pos is warp-uniform.
// (out of line on purpose: it is cold code -- resets and stone advances -- and inlining it at every call site
// made up a third of the step kernel's instructions, which the hot substep loop then fought for I-cache with)
__device__ __noinline__ int mt_doubles_impl(uint32_t* mt, int pos, double* out, int count, int lane) {
    int i = 0;
    while (i < count) {
        if (pos >= 624) { mt_twist_warp(mt, lane); pos = 0; }
        int m = min(count - i, (624 - pos) / 2);
        if (out)
            for (int t = lane; t < m; t += 32)
                out[i + t] = mt_make_double(mt_temper(mt[pos + 2 * t]), mt_temper(mt[pos + 2 * t + 1]));
        i += m; pos += 2 * m;
        if (i < count && pos == 623) {
            uint32_t a = mt_temper(mt[623]);
            __syncwarp();
            mt_twist_warp(mt, lane);
            uint32_t b = mt_temper(mt[0]);
            pos = 1;
            if (out && lane == 0) out[i] = mt_make_double(a, b);
            i++;
        }
    }
    __syncwarp();
    return pos;
}
__device__ __forceinline__ void mt_doubles_warp(uint32_t* mt, int& pos, double* out, int count, int lane) {
    pos = mt_doubles_impl(mt, pos, out, count, lane);
}
__device__ __forceinline__ double mt_one_double_warp(uint32_t* mt, int& pos, double* scratch, int lane) {
    if (pos <= SS_MT_N - 2) {              // both words lie in the current block: every lane computes the draw in registers
        const double u = mt_make_double(mt_temper(mt[pos]), mt_temper(mt[pos + 1]));
        pos += 2;
        return u;
    }
    // the draw straddles (or follows) a twist: the general routine through the scratch buffer.  The value is read back
    // through a volatile pointer after a warp barrier: a plain read of scratch[0] here was once satisfied with a stale
    // value left by an earlier launch (resample_kernel re-drew the previous uniform; tests/test_gpu_parity.py::
    // test_update_sample_prob_resamples_flagged_envs).
    mt_doubles_warp(mt, pos, scratch, 1, lane);
    __syncwarp();
    return reinterpret_cast<volatile double*>(scratch)[0];
}

// pybullet.getEulerFromQuaternion
__device__ __forceinline__ void quat_to_rpy(const float* q, float* rpy) {
    float sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    float sarg = -2.f * (q[0] * q[2] - q[3] * q[1]);
    const float HPI = 1.57079632679489662f;
    if (sarg <= -0.99999f) { rpy[0] = 0; rpy[1] = -HPI; rpy[2] = 2 * atan2f(q[0], -q[1]); }
    else if (sarg >= 0.99999f) { rpy[0] = 0; rpy[1] = HPI; rpy[2] = 2 * atan2f(-q[0], q[1]); }
    else {
        rpy[0] = atan2f(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        rpy[1] = asinf(sarg);
        rpy[2] = atan2f(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}

}  // namespace ssb
