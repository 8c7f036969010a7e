/*
 * ss_oracle.c -- CPU ORACLE (plain C, double precision).  TEST INFRASTRUCTURE ONLY -- see ss_oracle.h.
 *
 * Restates, for one env at a time:
 *   - the env logic of mocca_envs/walker3d_envs.py:482-1411 and mocca_envs/mike.py:25-921
 *     (each function cites the lines it follows),
 *   - the robot wrappers mocca_envs/robots.py:285-439, 581-618, 844-881,
 *   - the vec-env worker semantics common/envs_utils.py:25-40, 59-153, 645-649,
 *   - numpy's legacy RandomState (MT19937) so that terrain streams are seed-exact,
 *   - one Bullet multibody substep (third-party PyBullet, not in the reference tree; restated from its
 *     published algorithm: link-frame Featherstone ABA as in btMultiBody::
 *     computeAccelerationsArticulatedBodyAlgorithmMultiDof, unit-impulse responses as in
 *     calcAccelerationDeltasMultiDof, rows and PGS as in btMultiBodyConstraintSolver).  PARITY UNPINNED
 *     against PyBullet for this last part (per substep; behaviourally pinned by the reference's shipped
 *     policies, tests/test_policy_behaviour.py).
 *
 * The GPU library computes the same quantities by a different route (world-frame composite-rigid-body
 * inertia + LDL^T, Delassus-space PGS) in fp32, so agreement between the two is a real check.
 */
#include "ss_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NB SS_NB
#define NJ SS_NJ
#define ND SS_NDOF
#define MAXC SS_MAX_CONTACTS
#define MAXL SS_MAX_LIMIT_ROWS
#define MAXR (MAXL + 3 * MAXC)
#define DEG (M_PI / 180.0)

/* ------------------------------------------------------------------------------------------------
 * MT19937 + numpy legacy RandomState draws (random_sample = rk_double; uniform = lo + (hi-lo)*u)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; uint32_t pos; } mt_t;

static void mt_twist(mt_t* s) {
    uint32_t* mt = s->mt;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->pos = 0;
}
static uint32_t mt_u32(mt_t* s) {
    if (s->pos >= 624) mt_twist(s);
    uint32_t y = s->mt[s->pos++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
static double mt_double(mt_t* s) {
    uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

void sso_mt_doubles(uint32_t* mt625, double* out, int32_t n) {
    mt_t s; memcpy(s.mt, mt625, 624 * 4); s.pos = mt625[624];
    for (int i = 0; i < n; i++) out[i] = mt_double(&s);
    memcpy(mt625, s.mt, 624 * 4); mt625[624] = s.pos;
}

/* ------------------------------------------------------------------------------------------------
 * small linear algebra
 * ---------------------------------------------------------------------------------------------- */
static void m3mul(const double* A, const double* B, double* C) {
    double T[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, T, sizeof T);
}
static void m3v(const double* A, const double* x, double* y) {
    double t0 = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
    double t1 = A[3] * x[0] + A[4] * x[1] + A[5] * x[2];
    double t2 = A[6] * x[0] + A[7] * x[1] + A[8] * x[2];
    y[0] = t0; y[1] = t1; y[2] = t2;
}
static void m3tv(const double* A, const double* x, double* y) {
    double t0 = A[0] * x[0] + A[3] * x[1] + A[6] * x[2];
    double t1 = A[1] * x[0] + A[4] * x[1] + A[7] * x[2];
    double t2 = A[2] * x[0] + A[5] * x[1] + A[8] * x[2];
    y[0] = t0; y[1] = t1; y[2] = t2;
}
static void cross(const double* a, const double* b, double* c) {
    double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
    c[0] = t0; c[1] = t1; c[2] = t2;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void quat_to_mat(const double* q, double* R) { /* q = x,y,z,w */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void axis_angle_mat(const double* a, double th, double* R) {
    double c = cos(th), s = sin(th), t = 1 - c, x = a[0], y = a[1], z = a[2];
    R[0] = t * x * x + c; R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
    R[3] = t * x * y + s * z; R[4] = t * y * y + c; R[5] = t * y * z - s * x;
    R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}
/* getQuaternionFromEuler([roll,pitch,yaw]) as a matrix: Rz(yaw) Ry(pitch) Rx(roll) */
static void euler_to_mat(double roll, double pitch, double yaw, double* R) {
    double cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch), cy = cos(yaw), sy = sin(yaw);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}
/* pybullet.getEulerFromQuaternion (bullet_utils.py:71-72) */
static void quat_to_rpy(const double* q, double* rpy) {
    double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    double sarg = -2 * (q[0] * q[2] - q[3] * q[1]);
    if (sarg <= -0.99999) { rpy[0] = 0; rpy[1] = -0.5 * M_PI; rpy[2] = 2 * atan2(q[0], -q[1]); }
    else if (sarg >= 0.99999) { rpy[0] = 0; rpy[1] = 0.5 * M_PI; rpy[2] = 2 * atan2(-q[0], q[1]); }
    else {
        rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        rpy[1] = asin(sarg);
        rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}

typedef double v6[6];
typedef double m6[6][6];
static void m6v(m6 A, const double* x, double* y) {
    double t[6];
    for (int i = 0; i < 6; i++) { t[i] = 0; for (int j = 0; j < 6; j++) t[i] += A[i][j] * x[j]; }
    memcpy(y, t, sizeof t);
}
static void m6tv(m6 A, const double* x, double* y) {
    double t[6];
    for (int i = 0; i < 6; i++) { t[i] = 0; for (int j = 0; j < 6; j++) t[i] += A[j][i] * x[j]; }
    memcpy(y, t, sizeof t);
}
static double dot6(const double* a, const double* b) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }
/* motion cross v x m, force cross v x* f  (spatial vectors: angular first) */
static void crm(const double* v, const double* m, double* o) {
    double a[3], b[3], c[3];
    cross(v, m, a); cross(v, m + 3, b); cross(v + 3, m, c);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static void crf(const double* v, const double* f, double* o) {
    double a[3], b[3], c[3];
    cross(v, f, a); cross(v + 3, f + 3, b); cross(v, f + 3, c);
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
/* solve symmetric positive definite 6x6 */
static void solve6(m6 A, const double* b, double* x) {
    double M[6][7];
    for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) M[i][j] = A[i][j]; M[i][6] = b[i]; }
    for (int k = 0; k < 6; k++) {
        int p = k; for (int i = k + 1; i < 6; i++) if (fabs(M[i][k]) > fabs(M[p][k])) p = i;
        if (p != k) for (int j = 0; j < 7; j++) { double t = M[k][j]; M[k][j] = M[p][j]; M[p][j] = t; }
        for (int i = k + 1; i < 6; i++) { double f = M[i][k] / M[k][k]; for (int j = k; j < 7; j++) M[i][j] -= f * M[k][j]; }
    }
    for (int i = 5; i >= 0; i--) { double s = M[i][6]; for (int j = i + 1; j < 6; j++) s -= M[i][j] * x[j]; x[i] = s / M[i][i]; }
}

/* ------------------------------------------------------------------------------------------------
 * env state
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    double pos[3], quat[4], qj[NJ];
    double w[3], v[3], qdj[NJ];            /* base angular / linear velocity (world), hinge rates */
    double warm[SS_NWARM];
    double* terrain; double* plank;
    double potential, walk_target[3], initial_z;
    double feet_contact[2], feet_xyz[2][3], cache_contact[2];
    double ep_ret;
    int next_step, reached_count, mirrored, elapsed, ep_len, update_terrain, phantom_idx, started;
    mt_t rng, grng;
    double trace_q[4][SS_NQ], trace_qd[4][ND];
    int tslot;                                     /* sampling-table version this env draws from */
    int32_t ev_cur[SS_NEV], trace_ev[4][SS_NEV];   /* discrete decisions of each substep: limit rows (lo / hi hinge masks), contact ids */
    /* scratch shared between calc_* functions of one step */
    double body_rpy[3], joint_speeds[NJ], joint_angles[NJ], robot_state[50];
    int joints_at_limit, target_reached, done;
    double foot_dist[2], distance_to_target, targets[10];
    double progress, posture_penalty, energy_penalty, joints_penalty, tall_bonus, step_bonus, target_bonus;
} env_t;

struct sso_env {
    ss_model M; ss_params P; int n; env_t* e; double* cdf; int n_cells;      /* cdf: ring of SS_TABLE_RING versions */
    int table_head; long table_count;
};

static int g_threads = 0;
int sso_set_threads(int32_t n) { g_threads = n; return 0; }

/* ------------------------------------------------------------------------------------------------
 * kinematics
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    double Rw[NB][9], ow[NB][3], aw[NB][3];   /* world rotation, frame origin, hinge axis */
    double E[NB][9];                           /* child coords -> parent coords (R0 * rot(axis,q)) */
} kin_t;

static void fk(const ss_model* M, const env_t* e, kin_t* K) {
    quat_to_mat(e->quat, K->Rw[0]);
    memcpy(K->ow[0], e->pos, 24);
    K->aw[0][0] = K->aw[0][1] = K->aw[0][2] = 0;
    for (int i = 1; i < NB; i++) {
        int p = M->parent[i];
        double Rj[9];
        axis_angle_mat(M->axis[i], e->qj[i - 1], Rj);
        m3mul(M->R0[i], Rj, K->E[i]);
        m3mul(K->Rw[p], K->E[i], K->Rw[i]);
        double t[3]; m3v(K->Rw[p], M->anchor[i], t);
        for (int k = 0; k < 3; k++) K->ow[i][k] = K->ow[p][k] + t[k];
        m3v(K->Rw[i], M->axis[i], K->aw[i]);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Featherstone ABA in link coordinates (body frame origin = hinge anchor; spatial vectors angular-first)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    m6 X[NB];        /* motion transform parent coords -> child coords                       */
    m6 I[NB];        /* rigid-body spatial inertia, body frame                               */
    m6 IA[NB];       /* articulated inertia (after the inward pass)                          */
    v6 U[NB]; double Dinv[NB];
} aba_t;

static void aba_prepare(const ss_model* M, const kin_t* K, aba_t* A) {
    for (int i = 0; i < NB; i++) {
        /* spatial inertia about the frame origin */
        const double* c = M->com[i]; double m = M->mass[i];
        const double* in = M->inertia[i];
        double Ic[9] = { in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2] };
        double cc = dot3(c, c);
        memset(A->I[i], 0, sizeof(m6));
        for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++)
            A->I[i][r][s] = Ic[3 * r + s] + m * ((r == s ? cc : 0.0) - c[r] * c[s]);
        double cx[3][3] = { { 0, -c[2], c[1] }, { c[2], 0, -c[0] }, { -c[1], c[0], 0 } };
        for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) {
            A->I[i][r][3 + s] = m * cx[r][s];
            A->I[i][3 + r][s] = -m * cx[r][s];
        }
        for (int r = 0; r < 3; r++) A->I[i][3 + r][3 + r] = m;
        memcpy(A->IA[i], A->I[i], sizeof(m6));
        if (i == 0) continue;
        /* X = [E 0; -E rx, E],  E = parent->child rotation, r = anchor in parent coords */
        double Et[9]; for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) Et[3 * r + s] = K->E[i][3 * s + r];
        const double* r_ = M->anchor[i];
        double rx[9] = { 0, -r_[2], r_[1], r_[2], 0, -r_[0], -r_[1], r_[0], 0 };
        double Erx[9]; m3mul(Et, rx, Erx);
        memset(A->X[i], 0, sizeof(m6));
        for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) {
            A->X[i][r][s] = Et[3 * r + s]; A->X[i][3 + r][3 + s] = Et[3 * r + s];
            A->X[i][3 + r][s] = -Erx[3 * r + s];
        }
    }
    for (int i = NB - 1; i >= 1; i--) {
        int p = M->parent[i];
        double S[6] = { M->axis[i][0], M->axis[i][1], M->axis[i][2], 0, 0, 0 };
        m6v(A->IA[i], S, A->U[i]);
        double D = dot6(S, A->U[i]);
        A->Dinv[i] = 1.0 / D;
        m6 Ia, T;
        for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) Ia[r][s] = A->IA[i][r][s] - A->U[i][r] * A->U[i][s] * A->Dinv[i];
        /* IA[p] += X^T Ia X */
        for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) { double a = 0; for (int k = 0; k < 6; k++) a += Ia[r][k] * A->X[i][k][s]; T[r][s] = a; }
        for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) { double a = 0; for (int k = 0; k < 6; k++) a += A->X[i][k][r] * T[k][s]; A->IA[p][r][s] += a; }
    }
}

/* vector passes: given hinge torques tau[21], initial bias forces pA[NB] (body coords) and velocity
 * product accelerations c[NB] (NULL = zero), return the base spatial acceleration a0 (base coords) and
 * hinge accelerations qdd[21]. */
static void aba_solve(const ss_model* M, aba_t* A, const double* tau, v6* pA, v6* c, double* a0, double* qdd) {
    double u[NB]; v6 a[NB];
    for (int i = NB - 1; i >= 1; i--) {
        int p = M->parent[i];
        double S[6] = { M->axis[i][0], M->axis[i][1], M->axis[i][2], 0, 0, 0 };
        u[i] = tau[i - 1] - dot6(S, pA[i]);
        double pa[6];
        memcpy(pa, pA[i], sizeof pa);
        if (c) {
            double Ic[6]; m6v(A->IA[i], c[i], Ic);
            double Uc = dot6(A->U[i], c[i]);
            for (int k = 0; k < 6; k++) pa[k] += Ic[k] - A->U[i][k] * Uc * A->Dinv[i];
        }
        for (int k = 0; k < 6; k++) pa[k] += A->U[i][k] * u[i] * A->Dinv[i];
        double t[6]; m6tv(A->X[i], pa, t);
        for (int k = 0; k < 6; k++) pA[p][k] += t[k];
    }
    double nb[6]; for (int k = 0; k < 6; k++) nb[k] = -pA[0][k];
    solve6(A->IA[0], nb, a[0]);
    memcpy(a0, a[0], sizeof(v6));
    for (int i = 1; i < NB; i++) {
        int p = M->parent[i];
        double ap[6]; m6v(A->X[i], a[p], ap);
        if (c) for (int k = 0; k < 6; k++) ap[k] += c[i][k];
        qdd[i - 1] = (u[i] - dot6(A->U[i], ap)) * A->Dinv[i];
        memcpy(a[i], ap, sizeof ap);
        for (int k = 0; k < 3; k++) a[i][k] += M->axis[i][k] * qdd[i - 1];
    }
}

/* generalised acceleration (27: base angular world, base linear world, hinges) under hinge torques tau,
 * gravity g, at the env's current state */
static void forward_dynamics(const ss_model* M, const kin_t* K, aba_t* A, const env_t* e, const double* tau,
                             double g, double* acc) {
    v6 vel[NB], c[NB], pA[NB];
    m3tv(K->Rw[0], e->w, vel[0]); m3tv(K->Rw[0], e->v, vel[0] + 3);
    memset(c, 0, sizeof c);
    for (int i = 0; i < NB; i++) {
        if (i > 0) {
            int p = M->parent[i];
            m6v(A->X[i], vel[p], vel[i]);
            double vj[6] = { M->axis[i][0] * e->qdj[i - 1], M->axis[i][1] * e->qdj[i - 1], M->axis[i][2] * e->qdj[i - 1], 0, 0, 0 };
            for (int k = 0; k < 6; k++) vel[i][k] += vj[k];
            crm(vel[i], vj, c[i]);
        }
        double Iv[6]; m6v(A->I[i], vel[i], Iv);
        crf(vel[i], Iv, pA[i]);
        /* gravity as an external force at the centre of mass */
        double fw[3] = { 0, 0, -M->mass[i] * g }, fb[3], nb[3];
        m3tv(K->Rw[i], fw, fb); cross(M->com[i], fb, nb);
        for (int k = 0; k < 3; k++) { pA[i][k] -= nb[k]; pA[i][3 + k] -= fb[k]; }
    }
    double a0[6];
    aba_solve(M, A, tau, pA, c, a0, acc + 6);
    /* classical acceleration of the base origin = spatial + w x v (base coords), then to world */
    double wxv[3]; cross(vel[0], vel[0] + 3, wxv);
    double lin[3] = { a0[3] + wxv[0], a0[4] + wxv[1], a0[5] + wxv[2] };
    m3v(K->Rw[0], a0, acc); m3v(K->Rw[0], lin, acc + 3);
}

/* velocity change (27) produced by a unit impulse of a constraint row: either hinge j (dir = +-1) or a
 * world force `dir` at world point x on body k -- Bullet's calcAccelerationDeltasMultiDof */
static void unit_response(const ss_model* M, const kin_t* K, aba_t* A, int hinge, double hdir, int body,
                          const double* x, const double* dir, double* dv) {
    v6 pA[NB]; double tau[NJ];
    memset(pA, 0, sizeof pA); memset(tau, 0, sizeof tau);
    if (hinge >= 0) tau[hinge] = hdir;
    else {
        double r[3] = { x[0] - K->ow[body][0], x[1] - K->ow[body][1], x[2] - K->ow[body][2] }, nw[3], nb[3], fb[3];
        cross(r, dir, nw); m3tv(K->Rw[body], nw, nb); m3tv(K->Rw[body], dir, fb);
        for (int k = 0; k < 3; k++) { pA[body][k] = -nb[k]; pA[body][3 + k] = -fb[k]; }
    }
    double a0[6];
    aba_solve(M, A, tau, pA, NULL, a0, dv + 6);
    m3v(K->Rw[0], a0, dv); m3v(K->Rw[0], a0 + 3, dv + 3);
}

/* explicit Jacobian row (27) of a world direction at world point x on body k */
static void contact_jacobian(const ss_model* M, const kin_t* K, int body, const double* x, const double* dir, double* J) {
    memset(J, 0, ND * sizeof(double));
    double r[3] = { x[0] - K->ow[0][0], x[1] - K->ow[0][1], x[2] - K->ow[0][2] }, t[3];
    cross(r, dir, t);
    for (int k = 0; k < 3; k++) { J[k] = t[k]; J[3 + k] = dir[k]; }
    for (int i = body; i >= 1; i = M->parent[i]) {
        double ri[3] = { x[0] - K->ow[i][0], x[1] - K->ow[i][1], x[2] - K->ow[i][2] };
        cross(ri, dir, t);
        J[5 + i] = dot3(K->aw[i], t);
    }
}

/* ------------------------------------------------------------------------------------------------
 * collision.  Stone candidates: (geom, plank slot, segment end) -> id geom*8 + slot*2 + end, for every geom whose
 * link passes the stones' collision filter (collision_mode 0: the foot capsules only).  Self-collision candidates
 * (collision_mode 2): one closest-point pair per geom pair of M->pair_a/b -> id SS_NGEOM*8 + pair.  The SS_MAX_CONTACTS
 * deepest candidates (ties by id) become contacts, ordered by id.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int active, id, body, body_b, code; double depth, n[3], x[3], xb[3]; } cand_t;
/* cand_t.code: which branch of the contact geometry produced the point (parity probe, sso_substep_events): bit 1 the
 * segment missed the plank's footprint (whole segment kept), bit 2 the point lies inside the box (face normal), bits 3-4 that
 * face's axis.  Bit 0 is added by the row set-up: btPlaneSpace1's |n.z| > 0.7071 branch (the friction basis). */
#define MAXCAND (SS_NGEOM * 8 + SS_MAX_PAIRS)

static void plank_box(const ss_params* P, const double* pl, double* Rs, double* cb) {
    /* bullet_objects.py:75-80 + plank.urdf: the body is placed by its COM frame (q12) */
    euler_to_mat(pl[4], pl[5], pl[3], Rs);
    double off[3] = { 0, 0, P->stone_com_z + P->stone_box_z }, t[3];
    m3v(Rs, off, t);
    cb[0] = pl[0] + t[0]; cb[1] = pl[1] + t[1]; cb[2] = pl[2] - P->stone_com_z + t[2];
}

/* capsule segment [p0,p1] (world) of radius r against the plank box (Rs, cb): end 0 / 1 of the part of the segment
 * above the plank's footprint; returns 0 if this end yields no candidate */
static int capsule_box(const ss_params* P, const double* Rs, const double* cb, const double* w0, const double* w1, double r,
                       int end, double breaking, cand_t* c) {
    double p0[3], p1[3], a[3], bb[3];
    for (int i = 0; i < 3; i++) { p0[i] = w0[i] - cb[i]; p1[i] = w1[i] - cb[i]; }
    m3tv(Rs, p0, a); m3tv(Rs, p1, bb);
    /* clip the segment to the plank's footprint (local x,y slabs); keep the full segment if it misses */
    double t0 = 0, t1 = 1; int empty = 0;
    for (int ax = 0; ax < 2; ax++) {
        double d = bb[ax] - a[ax], h = P->stone_half[ax];
        if (fabs(d) < 1e-12) { if (fabs(a[ax]) > h) empty = 1; }
        else {
            double ta = (-h - a[ax]) / d, tb = (h - a[ax]) / d;
            if (ta > tb) { double s = ta; ta = tb; tb = s; }
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
        }
    }
    c->code = 0;
    if (empty || t0 > t1) { t0 = 0; t1 = 1; c->code |= 2; }
    if (end == 1 && t1 - t0 < 1e-6) return 0;
    double tt = end ? t1 : t0, s[3], q[3], nl[3], dist;
    for (int i = 0; i < 3; i++) s[i] = a[i] + tt * (bb[i] - a[i]);
    int inside = 1;
    for (int i = 0; i < 3; i++) {
        double h = P->stone_half[i];
        q[i] = s[i] < -h ? -h : (s[i] > h ? h : s[i]);
        if (q[i] != s[i]) inside = 0;
    }
    if (!inside) {
        double d[3] = { s[0] - q[0], s[1] - q[1], s[2] - q[2] };
        dist = sqrt(dot3(d, d));
        for (int i = 0; i < 3; i++) nl[i] = d[i] / dist;
    } else {
        int ax = 0; double best = 1e30;
        for (int i = 0; i < 3; i++) { double m = P->stone_half[i] - fabs(s[i]); if (m < best) { best = m; ax = i; } }
        nl[0] = nl[1] = nl[2] = 0; nl[ax] = s[ax] >= 0 ? 1.0 : -1.0;
        dist = -best;
        c->code |= 4 | (ax << 3);
    }
    c->depth = dist - r;
    if (!(c->depth < breaking)) return 0;
    m3v(Rs, nl, c->n);
    double sw[3]; m3v(Rs, s, sw);
    for (int i = 0; i < 3; i++) { c->x[i] = sw[i] + cb[i] - r * c->n[i]; c->xb[i] = c->x[i]; }
    return 1;
}

/* closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9) */
static void segment_segment(const double* p1, const double* q1, const double* p2, const double* q2, double* c1, double* c2) {
    double d1[3], d2[3], r[3];
    for (int i = 0; i < 3; i++) { d1[i] = q1[i] - p1[i]; d2[i] = q2[i] - p2[i]; r[i] = p1[i] - p2[i]; }
    double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), s, t;
    const double EPS = 1e-12;
    if (a <= EPS && e <= EPS) { s = t = 0; }
    else if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
    else {
        double c = dot3(d1, r);
        if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        else {
            double b = dot3(d1, d2), denom = a * e - b * b;
            if (denom > EPS * a * e) { s = (b * f - c * e) / denom; s = s < 0 ? 0 : (s > 1 ? 1 : s); } else s = 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
            else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        }
    }
    for (int i = 0; i < 3; i++) { c1[i] = p1[i] + d1[i] * s; c2[i] = p2[i] + d2[i] * t; }
}

/* returns the number of contacts written to out[] (<= SS_MAX_CONTACTS, ordered by id) */
static int detect(const ss_model* M, const ss_params* P, const env_t* e, const kin_t* K, cand_t* out) {
    static _Thread_local cand_t cand[MAXCAND];
    int n = 0;
    double gw0[SS_NGEOM][3], gw1[SS_NGEOM][3];
    for (int g = 0; g < M->n_geom; g++) {
        int b = M->geom_body[g]; double t[3];
        m3v(K->Rw[b], M->geom_p0[g], t); for (int i = 0; i < 3; i++) gw0[g][i] = K->ow[b][i] + t[i];
        m3v(K->Rw[b], M->geom_p1[g], t); for (int i = 0; i < 3; i++) gw1[g][i] = K->ow[b][i] + t[i];
    }
    for (int slot = 0; slot < 4; slot++) {
        int k = -1;
        for (int kk = e->next_step - 2; kk <= e->next_step + 1; kk++) if (kk >= 0 && (kk & 3) == slot) k = kk;
        if (k < 0 || k > P->n_steps - 1) continue;
        double Rs[9], cb[3];
        plank_box(P, e->plank + 6 * (k % P->rendered), Rs, cb);
        for (int g = 0; g < M->n_geom; g++) {
            if (!M->geom_stone[g] || (P->collision_mode == 0 && M->geom_foot[g] < 0)) continue;
            double breaking = M->geom_breaking[g] < P->stone_breaking ? M->geom_breaking[g] : P->stone_breaking;
            for (int end = 0; end < (M->geom_sphere[g] ? 1 : 2); end++) {
                cand_t c; memset(&c, 0, sizeof c);
                if (!capsule_box(P, Rs, cb, gw0[g], gw1[g], M->geom_radius[g], end, breaking, &c)) continue;
                c.active = 1; c.id = g * 8 + slot * 2 + end; c.body = M->geom_body[g]; c.body_b = -1;
                cand[n++] = c;
            }
        }
    }
    if (P->collision_mode >= 2)
        for (int p = 0; p < M->n_pair; p++) {
            int ga = M->pair_a[p], gb = M->pair_b[p];
            double ca[3], cb2[3], d[3];
            segment_segment(gw0[ga], gw1[ga], gw0[gb], gw1[gb], ca, cb2);
            for (int i = 0; i < 3; i++) d[i] = ca[i] - cb2[i];
            double dist = sqrt(dot3(d, d)), ra = M->geom_radius[ga], rb = M->geom_radius[gb];
            double breaking = M->geom_breaking[ga] < M->geom_breaking[gb] ? M->geom_breaking[ga] : M->geom_breaking[gb];
            if (!(dist - ra - rb < breaking)) continue;
            cand_t c; memset(&c, 0, sizeof c);
            c.active = 1; c.id = SS_NGEOM * 8 + p; c.body = M->geom_body[ga]; c.body_b = M->geom_body[gb];
            c.depth = dist - ra - rb;
            if (dist > 1e-9) for (int i = 0; i < 3; i++) c.n[i] = d[i] / dist; else { c.n[0] = c.n[1] = 0; c.n[2] = 1; }
            for (int i = 0; i < 3; i++) { c.x[i] = ca[i] - ra * c.n[i]; c.xb[i] = cb2[i] + rb * c.n[i]; }
            cand[n++] = c;
        }
    /* keep the SS_MAX_CONTACTS deepest (ties by id); candidates are already in id order within each class and the
     * stone ids (< SS_NGEOM*8) precede the pair ids, but the stone loop runs slot-major: sort the survivors by id */
    int m = 0;
    for (int l = 0; l < n; l++) {
        int rank = 0;
        for (int q = 0; q < n; q++)
            if (cand[q].depth < cand[l].depth || (cand[q].depth == cand[l].depth && cand[q].id < cand[l].id)) rank++;
        if (rank < MAXC) out[m++] = cand[l];
    }
    for (int i = 1; i < m; i++) { cand_t c = out[i]; int j = i - 1; while (j >= 0 && out[j].id > c.id) { out[j + 1] = out[j]; j--; } out[j + 1] = c; }
    return m;
}

/* does a contact touch foot f (right 0 / left 1)?  calc_feet_state counts contacts with ANY body, self-collisions
 * included */
static int contact_on_foot(const ss_model* M, const cand_t* c, int f) {
    int fb = M->foot_body[f];
    return c->body == fb || c->body_b == fb;
}

/* btPlaneSpace1 */
static void plane_space(const double* n, double* p, double* q) {
    if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
        double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

/* ------------------------------------------------------------------------------------------------
 * one physics substep (Bullet btMultiBodyDynamicsWorld::internalSingleStepSimulation for this scene)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double J[ND], W[ND], invD, rhs, cfm, lo, hi, lam, mu; int normal; } row_t;

static int clamp_vel(const ss_params* P, double* gv) {      /* returns 1 if the clamp engaged */
    int hit = 0;
    for (int k = 0; k < ND; k++) {
        if (gv[k] > P->max_coord_vel) { gv[k] = P->max_coord_vel; hit = 1; }
        if (gv[k] < -P->max_coord_vel) { gv[k] = -P->max_coord_vel; hit = 1; }
    }
    return hit;
}
static void pack_vel(const env_t* e, double* gv) { memcpy(gv, e->w, 24); memcpy(gv + 3, e->v, 24); memcpy(gv + 6, e->qdj, NJ * 8); }
static void unpack_vel(env_t* e, const double* gv) { memcpy(e->w, gv, 24); memcpy(e->v, gv + 3, 24); memcpy(e->qdj, gv + 6, NJ * 8); }

static void integrate(const ss_params* P, env_t* e) {
    double h = P->dt;
    for (int k = 0; k < 3; k++) e->pos[k] += e->v[k] * h;
    /* exponential map of the world angular velocity (btMultiBody::stepPositionsMultiDof) */
    double ang = sqrt(dot3(e->w, e->w)), ax[3];
    if (ang * h > 0.25 * M_PI) ang = 0.25 * M_PI / h;      /* ANGULAR_MOTION_THRESHOLD = 0.5*HALF_PI */
    double f = ang < 0.001 ? 0.5 * h - h * h * h * 0.020833333333 * ang * ang : sin(0.5 * ang * h) / ang;
    for (int k = 0; k < 3; k++) ax[k] = e->w[k] * f;
    double dq[4] = { ax[0], ax[1], ax[2], cos(0.5 * ang * h) }, *q = e->quat, r[4];
    r[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
    r[1] = dq[3] * q[1] - dq[0] * q[2] + dq[1] * q[3] + dq[2] * q[0];
    r[2] = dq[3] * q[2] + dq[0] * q[1] - dq[1] * q[0] + dq[2] * q[3];
    r[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
    double nn = 1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    for (int k = 0; k < 4; k++) q[k] = r[k] * nn;
    for (int j = 0; j < NJ; j++) e->qj[j] += e->qdj[j] * h;
}

static void substep(const ss_model* M, const ss_params* P, env_t* e, const double* tau) {
    kin_t K; aba_t A; cand_t con[MAXC];
    double h = P->dt;
    fk(M, e, &K);
    int nc = detect(M, P, e, &K, con);               /* collision detection at the current pose */
    e->cache_contact[0] = e->cache_contact[1] = 0;
    for (int c = 0; c < nc; c++) for (int f = 0; f < 2; f++) if (contact_on_foot(M, &con[c], f)) e->cache_contact[f] = 1;
    aba_prepare(M, &K, &A);
    double acc[ND], gv[ND];
    forward_dynamics(M, &K, &A, e, tau, P->gravity, acc);
    pack_vel(e, gv);
    for (int k = 0; k < ND; k++) gv[k] += h * acc[k];
    int clamped = clamp_vel(P, gv);

    /* ---- constraint rows: violated hinge limits, then contact normals, then 2 friction rows each ---- */
    row_t rows[MAXR]; int nl = 0;
    for (int j = 0; j < NJ && nl < MAXL; j++) {
        for (int side = 0; side < 2 && nl < MAXL; side++) {
            double pen = side ? M->q_hi[j] - e->qj[j] : e->qj[j] - M->q_lo[j];
            if (pen > 0) continue;
            row_t* r = &rows[nl++];
            memset(r, 0, sizeof *r);
            double dir = side ? -1.0 : 1.0;
            r->J[6 + j] = dir;
            unit_response(M, &K, &A, j, dir, 0, NULL, NULL, r->W);
            double d = 0, rel = 0; for (int k = 0; k < ND; k++) { d += r->J[k] * r->W[k]; rel += r->J[k] * gv[k]; }
            r->invD = 1.0 / d;
            r->rhs = (-pen * P->erp_limit / h - rel) * r->invD;
            r->cfm = 0; r->lo = 0; r->hi = P->limit_max_impulse; r->normal = -1;
        }
    }
    /* stones are soft (contactStiffness / contactDamping, bullet_objects.py:61-69); link-link contacts are hard:
     * erp = the default contact ERP (bullet_utils.py:324), cfm = 0 */
    double kd = h * P->contact_stiffness + P->contact_damping;
    double cfm_stone = 1.0 / kd / h, erp_stone = h * P->contact_stiffness / kd;
    for (int c = 0; c < nc; c++) {
        cand_t* cd = &con[c];
        int self = cd->body_b >= 0;
        double cfm = self ? 0.0 : cfm_stone, erp = self ? P->self_erp : erp_stone, mu = self ? P->self_friction : P->friction;
        double t1[3], t2[3]; plane_space(cd->n, t1, t2);
        const double* dirs[3] = { cd->n, t1, t2 };
        for (int a = 0; a < 3; a++) {
            row_t* r = &rows[a == 0 ? nl + c : nl + nc + 2 * c + (a - 1)];
            memset(r, 0, sizeof *r);
            contact_jacobian(M, &K, cd->body, cd->x, dirs[a], r->J);
            unit_response(M, &K, &A, -1, 0, cd->body, cd->x, dirs[a], r->W);
            double d = 0; for (int k = 0; k < ND; k++) d += r->J[k] * r->W[k];
            if (self) {
                /* body B of the same multibody gets the opposite impulse at its own surface point */
                double Jb[ND], Wb[ND];
                contact_jacobian(M, &K, cd->body_b, cd->xb, dirs[a], Jb);
                unit_response(M, &K, &A, -1, 0, cd->body_b, cd->xb, dirs[a], Wb);
                if (P->self_cross_terms) {
                    d = 0; for (int k = 0; k < ND; k++) d += (r->J[k] - Jb[k]) * (r->W[k] - Wb[k]);
                } else {
                    /* btMultiBodyConstraintSolver::setupMultiBodyContactConstraint: denom0 (A) + denom1 (B), each side
                     * on its own -- no cross terms even when A and B are links of one multibody */
                    for (int k = 0; k < ND; k++) d += Jb[k] * Wb[k];
                }
                for (int k = 0; k < ND; k++) { r->J[k] -= Jb[k]; r->W[k] -= Wb[k]; }
            }
            double rel = 0; for (int k = 0; k < ND; k++) rel += r->J[k] * gv[k];
            r->invD = 1.0 / (d + cfm);
            r->cfm = cfm * r->invD;
            if (a == 0) {
                double pen = cd->depth + P->linear_slop, poserr = 0, velerr = -rel;
                if (pen > 0) velerr -= pen / h; else poserr = -pen * erp / h;
                r->rhs = (poserr + velerr) * r->invD;
                r->lo = 0; r->hi = 1e10; r->normal = -1;
                for (int w = 0; w < MAXC; w++) if ((int)e->warm[w] == cd->id) r->lam = e->warm[MAXC + w] * P->warmstart;
            } else {
                r->rhs = -rel * r->invD;
                r->mu = mu; r->normal = nl + c;
            }
        }
    }
    int nr = nl + 3 * nc;
    /* the substep's discrete decisions (parity probe: sso_substep_events) */
    e->ev_cur[0] = e->ev_cur[1] = 0;
    for (int r = 0; r < nl; r++) for (int j = 0; j < NJ; j++) if (rows[r].J[6 + j] != 0) e->ev_cur[rows[r].J[6 + j] > 0 ? 0 : 1] |= 1 << j;
    for (int c = 0; c < MAXC; c++) e->ev_cur[2 + c] = c < nc ? (con[c].id | ((con[c].code | (fabs(con[c].n[2]) > 0.7071067811865475244008443621048490 ? 1 : 0)) << 16)) : -1;
    if (clamped) e->ev_cur[1] |= 1 << 30;           /* the max_coord_vel safety clamp engaged (before the solve) */
    for (int c = 0; c < MAXC; c++) e->ev_cur[10 + c] = 0;
    if (nr > 0) {
        double dv[ND]; memset(dv, 0, sizeof dv);
        for (int r = 0; r < nr; r++) if (rows[r].lam != 0) for (int k = 0; k < ND; k++) dv[k] += rows[r].W[k] * rows[r].lam;
        for (int it = 0; it < P->n_iters; it++) {
            for (int s = 0; s < nr; s++) {
                int ri = s;
                if (s < nl) ri = (it & 1) ? s : nl - 1 - s;     /* btMultiBodyConstraintSolver::solveSingleIteration */
                row_t* r = &rows[ri];
                if (r->normal >= 0) {
                    double tot = rows[r->normal].lam;
                    if (tot > 0 && it < 31) e->ev_cur[10 + (r->normal - nl)] |= 1 << it;     /* solver history (parity probe) */
                    if (!(tot > 0)) continue;
                    r->lo = -r->mu * tot; r->hi = r->mu * tot;
                }
                double jdv = 0; for (int k = 0; k < ND; k++) jdv += r->J[k] * dv[k];
                double dl = r->rhs - r->lam * r->cfm - jdv * r->invD;
                double sum = r->lam + dl;
                if (sum < r->lo) { dl = r->lo - r->lam; sum = r->lo; }
                else if (sum > r->hi) { dl = r->hi - r->lam; sum = r->hi; }
                r->lam = sum;
                for (int k = 0; k < ND; k++) dv[k] += r->W[k] * dl;
            }
        }
        for (int k = 0; k < ND; k++) gv[k] += dv[k];
        if (clamp_vel(P, gv)) e->ev_cur[1] |= 1 << 30;    /* ... or after it */
    }
    for (int w = 0; w < MAXC; w++) { e->warm[w] = -1; e->warm[MAXC + w] = 0; }
    for (int c = 0; c < nc; c++) { e->warm[c] = con[c].id; e->warm[MAXC + c] = rows[nl + c].lam; }
    unpack_vel(e, gv);
    integrate(P, e);
}

/* ------------------------------------------------------------------------------------------------
 * env logic
 * ---------------------------------------------------------------------------------------------- */
static double base_phi(const ss_params* P, const env_t* e, int k) {
    /* walker3d_envs.py:824-827 / mike.py:345-348 */
    double b = k == 0 ? -10.0 : (k == P->n_steps - 1 ? 10.0 : ((k & 1) ? 20.0 : -20.0));
    double sgn = e->mirrored ? 1.0 : -1.0;
    return DEG * b * sgn;
}

static double sgn(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

/* walker3d_envs.py:1245-1274 / mike.py:755-784 */
static void set_next_next_step_location(const ss_params* P, env_t* e, double pitch, double yaw, double dr, double xt, double yt) {
    const double* nx = e->terrain + 6 * e->next_step;
    int bci = (e->next_step + 1) % P->n_steps;
    double bphi = base_phi(P, e, bci), byaw = nx[3];
    double dx = dr * sin(pitch) * cos(yaw + bphi);
    dx = sgn(dx) * fmin(fmax(fabs(dx), P->step_radius * 2.5), P->r_max);
    double dy = dr * sin(pitch) * sin(yaw + bphi);
    double x = nx[0] + cos(byaw) * dx - sin(byaw) * dy;
    double y = nx[1] + sin(byaw) * dx + cos(byaw) * dy;
    double z = nx[2] + dr * cos(pitch);
    double* t = e->terrain + 6 * bci;
    t[0] = x; t[1] = y; t[2] = z; t[3] = yaw + byaw; t[4] = xt; t[5] = yt;
}

/* walker3d_envs.py:624-755 (W3D) / mike.py:167-276 (Mike) */
static void generate_step_placements(const ss_params* P, env_t* e) {
    int n = P->n_steps;
    double *dr = malloc(8 * n * 8), *dphi = dr + n, *dth = dr + 2 * n, *xt = dr + 3 * n, *yt = dr + 4 * n,
           *dphi_copy = dr + 5 * n, *cphi = dr + 6 * n, *tmp = dr + 7 * n;
    int mike = P->terrain_mode == 1;
    double dr_lo = 0.65, dr_hi = mike ? 0.65 : 0.8, y0 = mike ? -10 * DEG : 0.0, t0 = mike ? -20 * DEG : 0.0;
    for (int i = 0; i < n; i++) dr[i] = dr_lo + (dr_hi - dr_lo) * mt_double(&e->rng);
    for (int i = 0; i < n; i++) dphi[i] = y0 + (y0 - y0) * mt_double(&e->rng);
    for (int i = 0; i < n; i++) dth[i] = 90 * DEG + (90 * DEG - 90 * DEG) * mt_double(&e->rng);
    dr[0] = 0.0; dphi[0] = 0.0; dth[0] = M_PI / 2;
    dr[1] = 0.8; dphi[1] = 0.0; dth[1] = M_PI / 2;
    dr[2] = 0.75; dphi[2] = 0.0; dth[2] = M_PI / 2;
    for (int i = 0; i < n; i++) xt[i] = t0 + (t0 - t0) * mt_double(&e->rng);
    for (int i = 0; i < n; i++) yt[i] = 0.0 + 0.0 * mt_double(&e->rng);
    for (int i = 0; i < 3; i++) { xt[i] = 0; yt[i] = 0; }
    memcpy(dphi_copy, dphi, n * 8);
    double acc = 0; for (int i = 0; i < n; i++) { acc += dphi[i]; cphi[i] = acc; }
    double ax = 0, ay = 0, az = 0;
    double min_z = P->step_radius * sin(P->tilt_limit_deg * DEG) + 0.01;
    for (int i = 0; i < n; i++) {
        double bp = base_phi(P, e, i);
        double x_ = dr[i] * sin(dth[i]) * cos(cphi[i] + bp);
        if (i >= 2) x_ = sgn(x_) * fmin(fmax(fabs(x_), P->step_radius * 2.5), P->r_max);
        double y_ = dr[i] * sin(dth[i]) * sin(cphi[i] + bp);
        double z_ = dr[i] * cos(dth[i]);
        ax += x_; ay += y_; az += z_;
        double* t = e->terrain + 6 * i;
        t[0] = ax; t[1] = ay; t[2] = fmax(az + 0, min_z); t[3] = cphi[i]; t[4] = xt[i]; t[5] = yt[i];
    }
    (void)tmp;
    for (int i = 2; i < n - 1; i++) {
        const double* nx = e->terrain + 6 * i;
        int bci = (i + 1) % n;
        double bphi = base_phi(P, e, bci), byaw = nx[3];
        double pitch = dth[bci], yaw = dphi_copy[bci];
        if (!mike) {
            pitch = (i % 2 == 0) ? 120 * DEG : 60 * DEG;
            double u = mt_double(&e->grng);              /* the GLOBAL numpy stream, walker3d_envs.py:728 (q7) */
            if (u < 0.2) pitch = 90 * DEG;
            if (P->terrain_mode == 2) pitch = 90 * DEG;  /* flat switch (north-star "flat-terrain stage") */
        }
        double dx = dr[bci] * sin(pitch) * cos(yaw + bphi);
        dx = sgn(dx) * fmin(fmax(fabs(dx), P->step_radius * 2.5), P->r_max);
        double dy = dr[bci] * sin(pitch) * sin(yaw + bphi);
        double x = nx[0] + (cos(byaw) * dx + -sin(byaw) * dy);
        double y = nx[1] + (sin(byaw) * dx + cos(byaw) * dy);
        double z = nx[2] + dr[bci] * cos(pitch);
        double* t = e->terrain + 6 * bci;
        t[0] = x; t[1] = y; t[2] = z; t[3] = yaw + byaw;
        if (mike) t[4] = -20 * DEG;
    }
    free(dr);
}

/* robots.py:304-365 WalkerBase.calc_state (contact flags are NOT refreshed here: q1) */
static void calc_state(const ss_model* M, env_t* e) {
    e->joints_at_limit = 0;
    for (int j = 0; j < NJ; j++) {
        double mid = 0.5 * (M->q_lo[j] + M->q_hi[j]);
        e->joint_angles[j] = 2 * (e->qj[j] - mid) / (M->q_hi[j] - M->q_lo[j]);      /* bullet_utils.py:191-194 */
        e->joint_speeds[j] = 0.1 * e->qdj[j];
        if (fabs(e->joint_angles[j]) > 0.99) e->joints_at_limit++;
    }
    quat_to_rpy(e->quat, e->body_rpy);
    double yaw = e->body_rpy[2], c = cos(-yaw), s = sin(-yaw);
    double vx = c * e->v[0] - s * e->v[1], vy = s * e->v[0] + c * e->v[1], vz = e->v[2];
    double* st = e->robot_state;
    st[0] = e->pos[2] - e->initial_z; st[1] = vx; st[2] = vy; st[3] = vz; st[4] = e->body_rpy[0]; st[5] = e->body_rpy[1];
    for (int j = 0; j < NJ; j++) { st[6 + j] = e->joint_angles[j]; st[27 + j] = e->joint_speeds[j]; }
    st[48] = e->feet_contact[0]; st[49] = e->feet_contact[1];
    for (int k = 0; k < 50; k++) st[k] = st[k] < -5 ? -5 : (st[k] > 5 ? 5 : st[k]);   /* NaN passes through, as np.clip */
}

/* walker3d_envs.py:910-924 */
static void calc_potential(const ss_params* P, env_t* e) {
    double dx = e->walk_target[0] - e->pos[0], dy = e->walk_target[1] - e->pos[1];
    e->distance_to_target = sqrt(dx * dx + dy * dy);
    e->potential = -e->distance_to_target / P->control_dt;
}

/* walker3d_envs.py:798-810 */
static void update_steps(const ss_params* P, env_t* e) {
    int thr = P->rendered / 2;
    if (e->next_step >= thr) {
        int oldest = ((e->next_step - thr - 1) % P->rendered + P->rendered) % P->rendered;
        int next = e->next_step - thr - 1 + P->rendered;
        if (next > P->n_steps - 1) next = P->n_steps - 1;
        memcpy(e->plank + 6 * oldest, e->terrain + 6 * next, 48);
    }
}

/* walker3d_envs.py:961-999; feet positions are the foot links' COM (bullet_utils.py:94) */
static void calc_feet_state(const ss_model* M, const ss_params* P, env_t* e) {
    kin_t K; fk(M, e, &K);
    const double* p = e->terrain + 6 * e->next_step;
    e->target_reached = 0;
    for (int i = 0; i < 2; i++) {
        int b = M->foot_body[i]; double t[3];
        m3v(K.Rw[b], M->foot_ref[i], t);
        for (int k = 0; k < 3; k++) e->feet_xyz[i][k] = K.ow[b][k] + t[k];
        e->feet_contact[i] = e->cache_contact[i] != 0 ? 1.0 : 0.0;
        double d[3] = { e->feet_xyz[i][0] - p[0], e->feet_xyz[i][1] - p[1], e->feet_xyz[i][2] - p[2] };
        e->foot_dist[i] = sqrt(d[0] * d[0] + d[1] * d[1]);
        if (e->feet_contact[i] != 0 && (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < 0.3) e->target_reached = 1;
    }
    if (e->target_reached) {
        e->reached_count++;
        if (e->reached_count >= P->stop_frames) { e->next_step++; e->reached_count = 0; update_steps(P, e); }
        if (e->next_step >= P->n_steps) e->next_step--;
    }
}

/* walker3d_envs.py:1046-1077 */
static void delta_to_k_targets(const ss_params* P, env_t* e) {
    for (int k = 0; k < 2; k++) {
        int idx = e->next_step + k; if (idx > P->n_steps - 1) idx = P->n_steps - 1;
        const double* t = e->terrain + 6 * idx;
        if (k == 1) memcpy(e->walk_target, t, 24);                     /* q8 */
        double dx = t[0] - e->pos[0], dy = t[1] - e->pos[1], dz = t[2] - e->pos[2];
        double ang = atan2(dy, dx) - e->body_rpy[2], dist = sqrt(dx * dx + dy * dy);
        double* o = e->targets + 5 * k;
        o[0] = sin(ang) * dist; o[1] = cos(ang) * dist; o[2] = dz; o[3] = t[4]; o[4] = t[5];
    }
}

/* mike.py:660-692: one categorical draw over the 11^5 grid = one uniform + searchsorted(side='right') */
static void sample_next_next_step(const struct sso_env* S, env_t* e) {
    const ss_params* P = &S->P;
    double u = mt_double(&e->rng);
    int lo = 0, hi = S->n_cells;
    const double* cdf = S->cdf + (size_t)e->tslot * S->n_cells;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
    int idx = lo < S->n_cells ? lo : S->n_cells - 1, g = SS_GRID;
    int i4 = idx % g, i3 = (idx / g) % g, i2 = (idx / (g * g)) % g, i1 = (idx / (g * g * g)) % g, i0 = idx / (g * g * g * g);
    set_next_next_step_location(P, e, P->pitch_samples[i1] + M_PI / 2, P->yaw_samples[i0], P->r_samples[i2],
                                P->xtilt_samples[i3], P->ytilt_samples[i4]);
}

/* walker3d_envs.py:1079-1093 / mike.py:589-603 */
static void update_terrain_info(const struct sso_env* S, env_t* e) {
    const ss_params* P = &S->P;
    int nn = e->next_step + 1;
    if (P->use_sampler) sample_next_next_step(S, e);
    memcpy(e->plank + 6 * (nn % P->rendered), e->terrain + 6 * (nn % P->n_steps), 48);
    delta_to_k_targets(P, e);
}

/* walker3d_envs.py:926-959 */
static void calc_base_reward(const ss_params* P, env_t* e, const float* a) {
    double old = e->potential;
    calc_potential(P, e);
    e->progress = e->potential - old;
    e->posture_penalty = 0;
    if (!(-0.2 < e->body_rpy[1] && e->body_rpy[1] < 0.4)) e->posture_penalty = fabs(e->body_rpy[1]);
    if (!(-0.4 < e->body_rpy[0] && e->body_rpy[0] < 0.4)) e->posture_penalty += fabs(e->body_rpy[0]);
    double m1 = 0, m2 = 0;
    for (int j = 0; j < NJ; j++) { m1 += fabs((double)a[j] * e->joint_speeds[j]); m2 += (double)a[j] * (double)a[j]; }   /* unclipped action: q5 */
    e->energy_penalty = P->electricity_cost * (m1 / NJ) + P->stall_torque_cost * (m2 / NJ);
    e->joints_penalty = P->joints_at_limit_cost * e->joints_at_limit;
    double height = e->pos[2] - fmin(e->feet_xyz[0][2], e->feet_xyz[1][2]);
    e->tall_bonus = height > P->done_height ? 2.0 : -1.0;
    e->done = e->done || e->tall_bonus < 0;
}

/* walker3d_envs.py:1001-1024 / mike.py:522-534 */
static void calc_step_reward(const ss_params* P, env_t* e) {
    e->step_bonus = 0;
    if (e->target_reached && e->reached_count == 1) {
        e->step_bonus = 50 * exp(-fmin(e->foot_dist[0], e->foot_dist[1]) / 0.25);
        if (P->n_phantoms > 0 && e->next_step % 3 == 0)              /* walker3d_envs.py:1007-1016 (visual only) */
            e->phantom_idx = (e->phantom_idx + 1) % 20;
    }
    e->target_bonus = 0;
    if (e->next_step == P->n_steps - 1 && e->distance_to_target < 0.15) e->target_bonus = 2.0;
}

/* walker3d_envs.py:1026-1044 */
static void calc_env_state(const struct sso_env* S, env_t* e, const float* a) {
    const ss_params* P = &S->P;
    for (int k = 0; k < 50; k++) if (!isfinite(e->robot_state[k])) e->done = 1;
    int cur = e->next_step;
    calc_feet_state(&S->M, P, e);
    calc_base_reward(P, e, a);
    calc_step_reward(P, e);
    delta_to_k_targets(P, e);
    e->update_terrain = cur != e->next_step;
    if (cur != e->next_step) { update_terrain_info(S, e); calc_potential(P, e); }
}

static void write_obs(const env_t* e, float* obs) {
    for (int k = 0; k < 50; k++) obs[k] = (float)e->robot_state[k];
    for (int k = 0; k < 10; k++) obs[50 + k] = (float)e->targets[k];
    obs[0] = (float)(e->pos[2] - fmin(e->feet_xyz[0][2], e->feet_xyz[1][2]));       /* q6 */
}

/* Walker3DStepperEnv.reset walker3d_envs.py:812-867 / mike.py:333-388 + Walker3D.reset robots.py:581-618 */
static void env_reset(const struct sso_env* S, env_t* e, float* obs) {
    const ss_model* M = &S->M; const ss_params* P = &S->P;
    e->done = 0; e->reached_count = 0;
    /* robots.py:583 / :846: `self.np_random.rand() < p`.  robot.np_random is the RandomState EnvBase.__init__
     * created with seed(None) (env_base.py:26,65); env.seed(seed+rank) later REPLACES env.np_random
     * (env_base.py:131-133, envs_utils.py:28) but never re-links the robot's, so this coin (and the phantoms'
     * draws, walker3d_envs.py:569,861) come from a separate, entropy-seeded stream -- the "aux" stream here.
     * W3D's test is `< 0` (never true) and unobservable, so no draw is made for it. */
    e->mirrored = 0;
    if (P->mirror_prob > 0) { double u = mt_double(&e->grng); e->mirrored = u < P->mirror_prob; }
    memcpy(e->qj, M->init_q, sizeof e->qj);
    if (e->mirrored) {
        static const int R[9] = { 3, 4, 5, 6, 7, 13, 14, 15, 16 }, L[9] = { 8, 9, 10, 11, 12, 17, 18, 19, 20 };
        for (int k = 0; k < 9; k++) { e->qj[R[k]] = M->init_q[L[k]]; e->qj[L[k]] = M->init_q[R[k]]; }
        e->qj[0] *= -1; e->qj[2] *= -1;
    }
    memset(e->qdj, 0, sizeof e->qdj);
    memcpy(e->pos, M->init_pos, 24);
    e->quat[0] = e->quat[1] = e->quat[2] = 0; e->quat[3] = 1;
    /* robots.py:612-616: resetBasePositionAndOrientation then resetBaseVelocity(linearVelocity=vel).  The
     * second call leaves the angular velocity untouched, but Bullet's init-pose command behind
     * the first one zeroes BOTH base velocities whenever it sets position+orientation, so the net effect is
     * w = v = 0 after every reset. */
    e->v[0] = e->v[1] = e->v[2] = 0;
    e->w[0] = e->w[1] = e->w[2] = 0;
    for (int l = 0; l < MAXC; l++)                       /* the cache survives resets (q4); a NaN must not */
        if (!isfinite(e->warm[l]) || !isfinite(e->warm[MAXC + l])) { e->warm[l] = -1; e->warm[MAXC + l] = 0; }
    /* WalkerBase.reset robots.py:433-439 */
    e->feet_contact[0] = e->feet_contact[1] = 0;
    memset(e->feet_xyz, 0, sizeof e->feet_xyz);
    e->initial_z = e->pos[2];
    calc_state(M, e);
    /* stale-state calc_feet_state: q2 */
    calc_feet_state(M, P, e);
    generate_step_placements(P, e);
    for (int i = 0; i < P->rendered; i++) memcpy(e->plank + 6 * i, e->terrain + 6 * i, 48);
    e->next_step = 1;
    delta_to_k_targets(P, e);
    calc_potential(P, e);
    write_obs(e, obs);
    e->phantom_idx = 0;          /* walker3d_envs.py:859-862: phantom.reset draws from the robot stream, see above */
    /* update_terrain is NOT touched: the reference's reset() (walker3d_envs.py:812-867, mike.py:333-388) leaves the flag of
     * the terminal step standing, so a following update_sample_prob / create_temp_states acts on the freshly reset env when
     * the episode ended on a stone advance (golden: mike_ghost_sampleprob_at_done) */
    e->elapsed = 0; e->ep_ret = 0; e->ep_len = 0; e->started = 1;
}

/* actuator_mode 0: the torque the env applies, tau_max * clip(a), held over the control step (robots.py:289-293);
 * 1: the reference's abandoned PD block (robots.py:486-493), evaluated once per control step like it;
 * 2: the same PD law re-evaluated at every physics substep ("per-substep PD joint torques", BASELINE north star).
 * The MJCF joint damping torque is computed once per control step in every mode (PyBullet applies it per stepSimulation). */
static void apply_action(const ss_model* M, const ss_params* P, const env_t* e, const float* a, double* tau, const double* damp_qd) {
    for (int j = 0; j < NJ; j++) {
        double x = a[j] < -1 ? -1 : (a[j] > 1 ? 1 : (double)a[j]);
        if (P->actuator_mode == 0) tau[j] = M->tau_max[j] * x;            /* robots.py:289-293 */
        else {                                                            /* robots.py:486-493 (opt-in PD) */
            double mid = 0.5 * (M->q_lo[j] + M->q_hi[j]), rel = 2 * (e->qj[j] - mid) / (M->q_hi[j] - M->q_lo[j]);
            tau[j] = M->pd_p[j] * (x - rel) - M->pd_d[j] * e->qdj[j];
        }
        tau[j] -= M->joint_damping[j] * damp_qd[j];   /* PyBullet applies MJCF joint damping once per stepSimulation */
    }
}

static void env_step(const struct sso_env* S, env_t* e, const float* a, float* obs, float* rew, uint8_t* done,
                     uint8_t* timeout, float* ep_ret, int32_t* ep_len, int physics) {
    const ss_model* M = &S->M; const ss_params* P = &S->P;
    if (physics == 1) {
        double tau[NJ], qd0[NJ];
        memcpy(qd0, e->qdj, sizeof qd0);
        apply_action(M, P, e, a, tau, qd0);
        for (int s = 0; s < P->n_substeps; s++) {
            if (P->actuator_mode == 2 && s > 0) apply_action(M, P, e, a, tau, qd0);
            substep(M, P, e, tau);
            if (s < 4) {
                memcpy(e->trace_q[s], e->pos, 24); memcpy(e->trace_q[s] + 3, e->quat, 32); memcpy(e->trace_q[s] + 7, e->qj, NJ * 8);
                pack_vel(e, e->trace_qd[s]);
                memcpy(e->trace_ev[s], e->ev_cur, sizeof e->ev_cur);
            }
        }
    }
    calc_state(M, e);
    calc_env_state(S, e, a);
    double r = e->progress - e->energy_penalty;
    r += e->step_bonus + e->target_bonus - 0.0;
    r += e->tall_bonus - e->posture_penalty - e->joints_penalty;
    write_obs(e, obs);
    *rew = (float)r;
    if (!physics) { *done = (uint8_t)e->done; return; }
    /* gym TimeLimit + TimeLimitMask envs_utils.py:59-69 + Monitor :130-153 */
    e->elapsed++;
    int d = e->done;
    if (e->elapsed >= P->max_episode_steps) d = 1;
    *timeout = (uint8_t)(d && e->elapsed == P->max_episode_steps);
    e->ep_ret += r; e->ep_len++;
    *done = (uint8_t)d;
    *ep_ret = (float)(round(e->ep_ret * 1e6) / 1e6); *ep_len = e->ep_len;
    if (d) env_reset(S, e, obs);                                          /* envs_utils.py:646-649 */
}

/* ------------------------------------------------------------------------------------------------
 * public API
 * ---------------------------------------------------------------------------------------------- */
int sso_create(const ss_model* model, const ss_params* params, int32_t n, const uint32_t* mt_env,
               const uint32_t* mt_aux, sso_env** out) {
    if (!model || !params || n <= 0 || !mt_env || !out) return SS_ERR_ARG;
    sso_env* S = calloc(1, sizeof *S);
    S->M = *model; S->P = *params; S->n = n;
    S->e = calloc(n, sizeof(env_t));
    S->n_cells = SS_GRID_CELLS;
    S->cdf = malloc(sizeof(double) * S->n_cells * SS_TABLE_RING);
    S->table_head = 0; S->table_count = 1;
    for (int i = 0; i < S->n_cells; i++) S->cdf[i] = (double)(i + 1) / S->n_cells;
    for (int i = 0; i < n; i++) {
        env_t* e = &S->e[i];
        e->terrain = calloc(6 * params->n_steps, 8);
        e->plank = calloc(6 * params->rendered, 8);
        memcpy(e->rng.mt, mt_env + (size_t)i * SS_MT_WORDS, 624 * 4); e->rng.pos = mt_env[(size_t)i * SS_MT_WORDS + 624];
        if (mt_aux) { memcpy(e->grng.mt, mt_aux + (size_t)i * SS_MT_WORDS, 624 * 4); e->grng.pos = mt_aux[(size_t)i * SS_MT_WORDS + 624]; }
        e->quat[3] = 1;
        for (int w = 0; w < MAXC; w++) e->warm[w] = -1;
        memcpy(e->pos, model->init_pos, 24);
        memcpy(e->qj, model->init_q, sizeof e->qj);
    }
    *out = S;
    return 0;
}

void sso_destroy(sso_env* S) {
    if (!S) return;
    for (int i = 0; i < S->n; i++) { free(S->e[i].terrain); free(S->e[i].plank); }
    free(S->e); free(S->cdf); free(S);
}

int sso_reset_all(sso_env* S, float* obs) {
    for (int i = 0; i < S->n; i++) env_reset(S, &S->e[i], obs + 60 * (size_t)i);
    return 0;
}

int sso_step(sso_env* S, const float* act, float* obs, float* rew, uint8_t* done, uint8_t* timeout,
             float* ep_ret, int32_t* ep_len) {
    int nt = g_threads;
#ifdef _OPENMP
    if (nt <= 0) nt = omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(dynamic, 4)
#endif
    for (int i = 0; i < S->n; i++)
        env_step(S, &S->e[i], act + 21 * (size_t)i, obs + 60 * (size_t)i, rew + i, done + i, timeout + i, ep_ret + i, ep_len + i, 1);
    (void)nt;
    return 0;
}

/* env_step with the physics replaced by an injected post-physics state (golden tests): q [n,28], qd [n,27],
 * contact [n,2] = per-foot contact flags of the last substep's collision pass */
int sso_step_injected(sso_env* S, const float* act, const double* q, const double* qd, const double* contact, float* obs,
                      float* rew, uint8_t* done, uint8_t* timeout, float* ep_ret, int32_t* ep_len) {
    for (int i = 0; i < S->n; i++) {
        env_t* e = &S->e[i];
        memcpy(e->pos, q + (size_t)i * SS_NQ, 24); memcpy(e->quat, q + (size_t)i * SS_NQ + 3, 32); memcpy(e->qj, q + (size_t)i * SS_NQ + 7, NJ * 8);
        unpack_vel(e, qd + (size_t)i * ND);
        e->cache_contact[0] = contact[2 * i]; e->cache_contact[1] = contact[2 * i + 1];
        env_step(S, e, act + 21 * (size_t)i, obs + 60 * (size_t)i, rew + i, done + i, timeout + i, ep_ret + i, ep_len + i, 2);
    }
    return 0;
}

/* feet link COM positions (what getLinkState()[0] reports, bullet_utils.py:94) for a given q [28] */
int sso_feet_fk(sso_env* S, const double* q, double* out6) {
    env_t e; memset(&e, 0, sizeof e);
    memcpy(e.pos, q, 24); memcpy(e.quat, q + 3, 32); memcpy(e.qj, q + 7, NJ * 8);
    kin_t K; fk(&S->M, &e, &K);
    for (int i = 0; i < 2; i++) {
        int b = S->M.foot_body[i]; double t[3];
        m3v(K.Rw[b], S->M.foot_ref[i], t);
        for (int k = 0; k < 3; k++) out6[3 * i + k] = K.ow[b][k] + t[k];
    }
    return 0;
}

int sso_env_post(sso_env* S, const float* act, float* obs, float* rew, uint8_t* done) {
    uint8_t to; float er; int32_t el;
    for (int i = 0; i < S->n; i++)
        env_step(S, &S->e[i], act + 21 * (size_t)i, obs + 60 * (size_t)i, rew + i, done + i, &to, &er, &el, 0);
    return 0;
}

/* create_temp_states walker3d_envs.py:1189-1243 / mike.py:699-753 (variant 1: yaw x pitch at r = 0.7; 2: yaw x pitch x r;
 * 5: yaw x pitch x r x x_tilt x y_tilt).  For every env whose update_terrain flag is set: the next-next stone is moved
 * to every grid cell in turn (set_next_next_step_location(pi/2 - pitch, yaw, r, x_tilt, y_tilt)) and get_temp_state()
 * = concat(get_state(), delta_to_k_targets(2)) is recorded; afterwards the stone is put back with the env's
 * next_next_* parameters (`restore` = pitch, yaw, dr, x_tilt, y_tilt; NULL = the row the sampler wrote, i.e. unchanged).
 * Side effect the reference has and we keep: walk_target stays at the LAST grid cell's stone until the next
 * delta_to_k_targets.  Envs without the flag return zeros (self.temp_states).  grid = the five fake_* sample arrays. */
int sso_create_temp_states(sso_env* S, int32_t variant, const double* grid, const double* restore, float* out) {
    if (!S || !grid || !out || (variant != 1 && variant != 2 && variant != 5)) return SS_ERR_ARG;
    const ss_model* M = &S->M; const ss_params* P = &S->P;
    const int g = SS_GRID;
    const int n3 = variant >= 2 ? g : 1, n45 = variant == 5 ? g : 1;
    const size_t G = (size_t)g * g * n3 * n45 * n45;
    for (int i = 0; i < S->n; i++) {
        env_t* e = &S->e[i];
        float* o = out + (size_t)i * G * 60;
        if (!e->update_terrain) { memset(o, 0, G * 60 * sizeof(float)); continue; }
        int bci = (e->next_step + 1) % P->n_steps;
        double saved[6]; memcpy(saved, e->terrain + 6 * bci, 48);
        for (int iy = 0; iy < g; iy++)
            for (int ip = 0; ip < g; ip++)
                for (int ir = 0; ir < n3; ir++)
                    for (int ix = 0; ix < n45; ix++)
                        for (int it = 0; it < n45; it++) {
                            double r = variant == 1 ? 0.7 : grid[2 * g + ir];
                            double xt = variant == 5 ? grid[3 * g + ix] : 0.0, yt = variant == 5 ? grid[4 * g + it] : 0.0;
                            set_next_next_step_location(P, e, M_PI / 2 - grid[g + ip], grid[iy], r, xt, yt);
                            calc_state(M, e);                                  /* get_state: mike.py:390-394 */
                            delta_to_k_targets(P, e);
                            write_obs(e, o);
                            o += 60;
                        }
        if (restore) set_next_next_step_location(P, e, restore[0], restore[1], restore[2], restore[3], restore[4]);
        else memcpy(e->terrain + 6 * bci, saved, 48);
    }
    return 0;
}

/* Every call installs a new table VERSION in a ring of SS_TABLE_RING.  apply_now == 0 (update_curriculum /
 * update_specialist, called on every env: envs_utils.py:580-586): all envs take it.  apply_now != 0 (update_sample_prob,
 * walker3d_envs.py:1317-1320 / mike.py:827-830): `if self.update_terrain: self.yaw_pitch_r_tilt_prob = ...;
 * self.update_terrain_info()` -- only the envs that advanced a stone in their last step take the table and re-sample; the
 * others keep the one they have.  When the ring wraps, envs still on the version being overwritten move to the oldest
 * surviving one. */
int sso_set_sample_cdf(sso_env* S, const double* cdf, int32_t n_cells, int32_t apply_now) {
    if (n_cells != SS_GRID_CELLS) return SS_ERR_ARG;
    const int slot = (S->table_head + 1) % SS_TABLE_RING;
    if (S->table_count >= SS_TABLE_RING)
        for (int i = 0; i < S->n; i++) if (S->e[i].tslot == slot) S->e[i].tslot = (slot + 1) % SS_TABLE_RING;
    memcpy(S->cdf + (size_t)slot * n_cells, cdf, sizeof(double) * n_cells);
    S->table_head = slot; S->table_count++;
    for (int i = 0; i < S->n; i++) {
        env_t* e = &S->e[i];
        if (!apply_now) e->tslot = slot;
        else if (e->update_terrain) { e->tslot = slot; update_terrain_info(S, e); }
    }
    return 0;
}

int sso_get_state(sso_env* S, ss_state_view* v) {
    int ns = S->P.n_steps, nr = S->P.rendered;
    for (int i = 0; i < S->n; i++) {
        env_t* e = &S->e[i];
        if (v->q) { double* q = v->q + (size_t)i * SS_NQ; memcpy(q, e->pos, 24); memcpy(q + 3, e->quat, 32); memcpy(q + 7, e->qj, NJ * 8); }
        if (v->qd) pack_vel(e, v->qd + (size_t)i * ND);
        if (v->warm) memcpy(v->warm + (size_t)i * SS_NWARM, e->warm, sizeof e->warm);
        if (v->terrain) memcpy(v->terrain + (size_t)i * ns * 6, e->terrain, ns * 48);
        if (v->plank) memcpy(v->plank + (size_t)i * nr * 6, e->plank, nr * 48);
        if (v->scalars) {
            double* s = v->scalars + (size_t)i * SS_NSCAL;
            s[SS_S_POTENTIAL] = e->potential; memcpy(s + SS_S_WALK_X, e->walk_target, 24); s[SS_S_INITIAL_Z] = e->initial_z;
            s[SS_S_FEET_CONTACT0] = e->feet_contact[0]; s[SS_S_FEET_CONTACT1] = e->feet_contact[1];
            memcpy(s + SS_S_FEET_X0, e->feet_xyz, 48); s[SS_S_EP_RET] = e->ep_ret;
            s[SS_S_CACHE_CONTACT0] = e->cache_contact[0]; s[SS_S_CACHE_CONTACT1] = e->cache_contact[1];
        }
        if (v->ints) {
            int32_t* t = v->ints + (size_t)i * SS_NINT;
            t[SS_I_NEXT_STEP] = e->next_step; t[SS_I_REACHED_COUNT] = e->reached_count; t[SS_I_MIRRORED] = e->mirrored;
            t[SS_I_ELAPSED] = e->elapsed; t[SS_I_EP_LEN] = e->ep_len; t[SS_I_UPDATE_TERRAIN] = e->update_terrain;
            t[SS_I_PHANTOM_IDX] = e->phantom_idx; t[SS_I_STARTED] = e->started;
        }
        if (v->mt) {
            uint32_t* m = v->mt + (size_t)i * 2 * SS_MT_WORDS;
            memcpy(m, e->rng.mt, 624 * 4); m[624] = e->rng.pos;
            memcpy(m + SS_MT_WORDS, e->grng.mt, 624 * 4); m[SS_MT_WORDS + 624] = e->grng.pos;
        }
        if (v->table_slot) v->table_slot[i] = e->tslot;
    }
    return 0;
}

int sso_table_ring_state(const sso_env* S, int32_t* head, int64_t* count) { *head = S->table_head; *count = S->table_count; return 0; }

int sso_set_state(sso_env* S, const ss_state_view* v) {
    int ns = S->P.n_steps, nr = S->P.rendered;
    for (int i = 0; i < S->n; i++) {
        env_t* e = &S->e[i];
        if (v->q) { const double* q = v->q + (size_t)i * SS_NQ; memcpy(e->pos, q, 24); memcpy(e->quat, q + 3, 32); memcpy(e->qj, q + 7, NJ * 8); }
        if (v->qd) unpack_vel(e, v->qd + (size_t)i * ND);
        if (v->warm) memcpy(e->warm, v->warm + (size_t)i * SS_NWARM, sizeof e->warm);
        if (v->terrain) memcpy(e->terrain, v->terrain + (size_t)i * ns * 6, ns * 48);
        if (v->plank) memcpy(e->plank, v->plank + (size_t)i * nr * 6, nr * 48);
        if (v->scalars) {
            const double* s = v->scalars + (size_t)i * SS_NSCAL;
            e->potential = s[SS_S_POTENTIAL]; memcpy(e->walk_target, s + SS_S_WALK_X, 24); e->initial_z = s[SS_S_INITIAL_Z];
            e->feet_contact[0] = s[SS_S_FEET_CONTACT0]; e->feet_contact[1] = s[SS_S_FEET_CONTACT1];
            memcpy(e->feet_xyz, s + SS_S_FEET_X0, 48); e->ep_ret = s[SS_S_EP_RET];
            e->cache_contact[0] = s[SS_S_CACHE_CONTACT0]; e->cache_contact[1] = s[SS_S_CACHE_CONTACT1];
        }
        if (v->ints) {
            const int32_t* t = v->ints + (size_t)i * SS_NINT;
            e->next_step = t[SS_I_NEXT_STEP]; e->reached_count = t[SS_I_REACHED_COUNT]; e->mirrored = t[SS_I_MIRRORED];
            e->elapsed = t[SS_I_ELAPSED]; e->ep_len = t[SS_I_EP_LEN]; e->update_terrain = t[SS_I_UPDATE_TERRAIN];
            e->phantom_idx = t[SS_I_PHANTOM_IDX]; e->started = t[SS_I_STARTED];
        }
        if (v->mt) {
            const uint32_t* m = v->mt + (size_t)i * 2 * SS_MT_WORDS;
            memcpy(e->rng.mt, m, 624 * 4); e->rng.pos = m[624];
            memcpy(e->grng.mt, m + SS_MT_WORDS, 624 * 4); e->grng.pos = m[SS_MT_WORDS + 624];
        }
        if (v->table_slot) e->tslot = v->table_slot[i] >= 0 && v->table_slot[i] < SS_TABLE_RING ? v->table_slot[i] : 0;
    }
    return 0;
}

int sso_substep_trace(sso_env* S, double* q, double* qd) {
    for (int s = 0; s < 4; s++) for (int i = 0; i < S->n; i++) {
        memcpy(q + ((size_t)s * S->n + i) * SS_NQ, S->e[i].trace_q[s], SS_NQ * 8);
        memcpy(qd + ((size_t)s * S->n + i) * ND, S->e[i].trace_qd[s], ND * 8);
    }
    return 0;
}

int sso_substep_events(sso_env* S, int32_t* ev) {
    for (int s = 0; s < 4; s++) for (int i = 0; i < S->n; i++) memcpy(ev + ((size_t)s * S->n + i) * SS_NEV, S->e[i].trace_ev[s], SS_NEV * 4);
    return 0;
}

int sso_aba_substep(sso_env* S, double* q, double* qd, const double* tau, int32_t n) {
    for (int i = 0; i < n; i++) {
        env_t e; memset(&e, 0, sizeof e);
        memcpy(e.pos, q + (size_t)i * SS_NQ, 24); memcpy(e.quat, q + (size_t)i * SS_NQ + 3, 32); memcpy(e.qj, q + (size_t)i * SS_NQ + 7, NJ * 8);
        unpack_vel(&e, qd + (size_t)i * ND);
        kin_t K; aba_t A; double acc[ND], gv[ND];
        fk(&S->M, &e, &K); aba_prepare(&S->M, &K, &A);
        forward_dynamics(&S->M, &K, &A, &e, tau + (size_t)i * NJ, S->P.gravity, acc);
        pack_vel(&e, gv);
        for (int k = 0; k < ND; k++) gv[k] += S->P.dt * acc[k];
        clamp_vel(&S->P, gv);
        unpack_vel(&e, gv);
        integrate(&S->P, &e);
        memcpy(q + (size_t)i * SS_NQ, e.pos, 24); memcpy(q + (size_t)i * SS_NQ + 3, e.quat, 32); memcpy(q + (size_t)i * SS_NQ + 7, e.qj, NJ * 8);
        pack_vel(&e, qd + (size_t)i * ND);
    }
    return 0;
}

int sso_mass_matrix(sso_env* S, int32_t i, double* H, double* C) {
    env_t* e = &S->e[i];
    kin_t K; aba_t A;
    fk(&S->M, e, &K); aba_prepare(&S->M, &K, &A);
    /* columns of H^-1 from unit generalised forces, then invert numerically */
    static double Hinv[ND][ND];
    for (int d = 0; d < ND; d++) {
        double dv[ND];
        if (d >= 6) unit_response(&S->M, &K, &A, d - 6, 1.0, 0, NULL, NULL, dv);
        else {
            /* unit generalised base force: torque about / force at the base origin, world axes */
            v6 pA[NB]; double tau[NJ], a0[6]; memset(pA, 0, sizeof pA); memset(tau, 0, sizeof tau);
            double ew[3] = { 0, 0, 0 }, eb[3]; ew[d % 3] = 1; m3tv(K.Rw[0], ew, eb);
            for (int k = 0; k < 3; k++) pA[0][(d < 3 ? 0 : 3) + k] = -eb[k];
            aba_solve(&S->M, &A, tau, pA, NULL, a0, dv + 6);
            m3v(K.Rw[0], a0, dv); m3v(K.Rw[0], a0 + 3, dv + 3);
        }
        for (int k = 0; k < ND; k++) Hinv[k][d] = dv[k];
    }
    /* Gauss-Jordan inverse */
    static double Aug[ND][2 * ND];
    for (int r = 0; r < ND; r++) for (int c = 0; c < ND; c++) { Aug[r][c] = Hinv[r][c]; Aug[r][ND + c] = r == c; }
    for (int k = 0; k < ND; k++) {
        int p = k; for (int r = k + 1; r < ND; r++) if (fabs(Aug[r][k]) > fabs(Aug[p][k])) p = r;
        for (int c = 0; c < 2 * ND; c++) { double t = Aug[k][c]; Aug[k][c] = Aug[p][c]; Aug[p][c] = t; }
        double f = 1.0 / Aug[k][k]; for (int c = 0; c < 2 * ND; c++) Aug[k][c] *= f;
        for (int r = 0; r < ND; r++) if (r != k) { double g = Aug[r][k]; for (int c = 0; c < 2 * ND; c++) Aug[r][c] -= g * Aug[k][c]; }
    }
    for (int r = 0; r < ND; r++) for (int c = 0; c < ND; c++) H[r * ND + c] = Aug[r][ND + c];
    /* C = -H * acc(tau = 0) */
    double tau0[NJ], acc[ND]; memset(tau0, 0, sizeof tau0);
    forward_dynamics(&S->M, &K, &A, e, tau0, S->P.gravity, acc);
    for (int r = 0; r < ND; r++) { double s = 0; for (int c = 0; c < ND; c++) s += H[r * ND + c] * acc[c]; C[r] = -s; }
    return 0;
}

/* mechanical invariants of a state (physics sanity tests): out[0] total mass, [1..3] centre of mass, [4..6] linear
 * momentum, [7..9] angular momentum about the centre of mass, [10] kinetic energy, [11] potential energy (m g z).
 * Computed body by body from world-frame rigid-body velocities -- independent of the ABA / joint-space code. */
int sso_invariants(sso_env* S, const double* q, const double* qd, double* out) {
    const ss_model* M = &S->M;
    env_t e; memset(&e, 0, sizeof e);
    memcpy(e.pos, q, 24); memcpy(e.quat, q + 3, 32); memcpy(e.qj, q + 7, NJ * 8);
    unpack_vel(&e, qd);
    kin_t K; fk(M, &e, &K);
    double w[NB][3], vo[NB][3];                 /* world angular velocity and velocity of each frame origin */
    memcpy(w[0], e.w, 24); memcpy(vo[0], e.v, 24);
    for (int i = 1; i < NB; i++) {
        int p = M->parent[i];
        double r[3] = { K.ow[i][0] - K.ow[p][0], K.ow[i][1] - K.ow[p][1], K.ow[i][2] - K.ow[p][2] }, t[3];
        cross(w[p], r, t);
        for (int k = 0; k < 3; k++) { vo[i][k] = vo[p][k] + t[k]; w[i][k] = w[p][k] + K.aw[i][k] * e.qdj[i - 1]; }
    }
    double mt = 0, c[3] = { 0, 0, 0 }, P[3] = { 0, 0, 0 }, ke = 0, pe = 0;
    double cw[NB][3], vc[NB][3], Lb[NB][3];
    for (int i = 0; i < NB; i++) {
        double t[3], u[3];
        m3v(K.Rw[i], M->com[i], t);
        for (int k = 0; k < 3; k++) cw[i][k] = K.ow[i][k] + t[k];
        cross(w[i], t, u);
        for (int k = 0; k < 3; k++) vc[i][k] = vo[i][k] + u[k];
        double wb[3], Iw[3]; m3tv(K.Rw[i], w[i], wb);
        const double* in = M->inertia[i];
        double Ib[3] = { in[0] * wb[0] + in[3] * wb[1] + in[4] * wb[2], in[3] * wb[0] + in[1] * wb[1] + in[5] * wb[2],
                         in[4] * wb[0] + in[5] * wb[1] + in[2] * wb[2] };
        m3v(K.Rw[i], Ib, Iw); memcpy(Lb[i], Iw, 24);
        double m = M->mass[i];
        mt += m; for (int k = 0; k < 3; k++) { c[k] += m * cw[i][k]; P[k] += m * vc[i][k]; }
        ke += 0.5 * m * dot3(vc[i], vc[i]) + 0.5 * dot3(w[i], Iw);
        pe += m * S->P.gravity * cw[i][2];
    }
    for (int k = 0; k < 3; k++) c[k] /= mt;
    double L[3] = { 0, 0, 0 };
    for (int i = 0; i < NB; i++) {
        double r[3] = { cw[i][0] - c[0], cw[i][1] - c[1], cw[i][2] - c[2] }, mv[3] = { M->mass[i] * vc[i][0], M->mass[i] * vc[i][1], M->mass[i] * vc[i][2] }, t[3];
        cross(r, mv, t);
        for (int k = 0; k < 3; k++) L[k] += t[k] + Lb[i][k];
    }
    out[0] = mt; memcpy(out + 1, c, 24); memcpy(out + 4, P, 24); memcpy(out + 7, L, 24); out[10] = ke; out[11] = pe;
    return 0;
}

int sso_contacts(sso_env* S, int32_t i, double* out) {
    kin_t K; cand_t con[MAXC];
    fk(&S->M, &S->e[i], &K);
    int nc = detect(&S->M, &S->P, &S->e[i], &K, con);
    memset(out, 0, sizeof(double) * MAXC * 16);
    for (int c = 0; c < nc; c++) {
        double* o = out + 16 * c;
        o[0] = 1; o[1] = con[c].depth; memcpy(o + 2, con[c].n, 24); memcpy(o + 5, con[c].x, 24);
        o[8] = con[c].id; o[9] = con[c].body; o[10] = con[c].body_b; memcpy(o + 11, con[c].xb, 24);
    }
    return 0;
}

/* algorithms/storage.py:59-71 */
int sso_gae(const float* rewards, const float* value_preds, const float* masks, const float* bad_masks,
            float* returns, int32_t T, int32_t N, float gamma, float lam) {
    for (int i = 0; i < N; i++) {
        float gae = 0;
        for (int t = T - 1; t >= 0; t--) {
            float delta = rewards[(size_t)t * N + i] + gamma * value_preds[(size_t)(t + 1) * N + i] * masks[(size_t)(t + 1) * N + i] - value_preds[(size_t)t * N + i];
            gae = delta + gamma * lam * masks[(size_t)(t + 1) * N + i] * gae;
            gae = gae * bad_masks[(size_t)(t + 1) * N + i];
            returns[(size_t)t * N + i] = gae + value_preds[(size_t)t * N + i];
        }
    }
    return 0;
}
