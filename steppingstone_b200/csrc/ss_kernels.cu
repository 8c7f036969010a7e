// ss_kernels.cu -- kernels + C-ABI of libss_b200.so (sm_100a).  See include/ss_env.h for the contract and
// ss_device.cuh for the warp-level physics.  One warp steps one env through a whole control step:
//   apply_action -> 4 x (dynamics, collision, rows, PGS, integrate) -> calc_state -> calc_env_state ->
//   reward / done / Monitor / TimeLimit -> auto-reset (terrain generator on the env's MT19937 stream).
// There is no CPU fallback: every entry point needs a CUDA device and fails with SS_ERR_CUDA otherwise.
#include "ss_device.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <chrono>

namespace ssb {

// CTA shape of the env kernels.  The substep loop is ~9k straight-line instructions (144 KB): warps that drift apart
// thrash the instruction caches (ncu: stall_no_instruction was the top stall).  The warps of a CTA are re-aligned by one
// __syncthreads() per substep so that they share their instruction fetches (without it the kernel is 1.6x slower).
// ONE CTA of 20 warps per SM (96 registers, 210 KB shared memory): all resident warps of an SM move in step
// (2 CTAs of 10 warps: 0.826 ms; 1 CTA of 20: 0.802 ms at N=8192; profiles/README.md).
#ifndef SS_WPB
#define SS_WPB 20
#endif
#ifndef SS_SYNC
#define SS_SYNC 1
#endif
#ifndef SS_MINB
#define SS_MINB 1
#endif
constexpr int WARPS_PER_BLOCK = SS_WPB;
constexpr int BLOCK = 32 * WARPS_PER_BLOCK;

struct alignas(16) BlockSmem {
    DevModel M;
    unsigned long long bar;
    unsigned long long pad;
    WarpSmem W[WARPS_PER_BLOCK];
};

#ifndef SS_ABA_WPB
#define SS_ABA_WPB 8
#endif
#ifndef SS_ABA_MINB
#define SS_ABA_MINB 4
#endif
constexpr int ABA_WPB = SS_ABA_WPB;   // the stand-alone forward-dynamics kernel: 8-warp CTAs, 4 per SM (64 registers; 3 per SM / 80 registers is 7 % slower)
constexpr int ABA_BLOCK = 32 * ABA_WPB;
struct alignas(16) BlockSmemDyn {
    DevModel M;
    unsigned long long bar;
    unsigned long long pad;
    WarpSmemDyn W[ABA_WPB];
};

// ---------------------------------------------------------------------------------------------
// TMA bulk copy of the constant tables into shared memory (cp.async.bulk -> UBLKCP) + mbarrier wait
// ---------------------------------------------------------------------------------------------
#ifdef SS_SIMT_EMU   // tests/simt_emu builds this file for the host (test infrastructure): the bulk copy is a memcpy
template <typename BS>
__device__ __forceinline__ void stage_model_issue(BS& B, const DevModel* gm) {
    if (threadIdx.x == 0) memcpy(&B.M, gm, sizeof(DevModel));
    __syncthreads();
}
template <typename BS>
__device__ __forceinline__ void stage_model_wait(BS&) {}
#else
template <typename BS>
__device__ __forceinline__ void stage_model_issue(BS& B, const DevModel* gm) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&B.bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&B.M);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)sizeof(DevModel)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(gm), "r"((uint32_t)sizeof(DevModel)), "r"(bar) : "memory");
    }
}
template <typename BS>
__device__ __forceinline__ void stage_model_wait(BS& B) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&B.bar);
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
    }
}
#endif
template <typename BS>
__device__ __forceinline__ void stage_model(BS& B, const DevModel* gm) {
    stage_model_issue(B, gm);
    stage_model_wait(B);
}

// ---------------------------------------------------------------------------------------------
// per-env register context (all fields warp-uniform unless noted)
// ---------------------------------------------------------------------------------------------
struct Env {
    double pos[3];
    float quat[4];
    float q_dof, v_dof, warm;            // lane = dof / candidate slot
    double potential, walk[3], initial_z, ep_ret;
    float feet_contact[2], cache_contact[2];
    double feet_xyz[2][3];
    int next_step, reached, mirrored, elapsed, ep_len, update_terrain, phantom_idx, started;
    int mtpos, gmtpos;
};

// state the physics needs (kept in registers across the substeps)
__device__ __forceinline__ void env_load_phys(const DevState& S, int e, int lane, Env& E) {
    const float* sv = S.sv + (size_t)e * 64;
    float a = sv[lane], b = sv[32 + lane];
#pragma unroll
    for (int i = 0; i < 4; i++) E.quat[i] = shfl(a, i);
    E.q_dof = shfl(a, (lane + 30) & 31);            // hinge j sits at sv[4+j]; dof 6+j reads lane j+4
    if (lane < 6 || lane >= ND) E.q_dof = 0.f;
    E.v_dof = lane < ND ? b : 0.f;
    E.warm = lane < SS_NWARM ? S.warm[(size_t)e * SS_NWARM + lane] : 0.f;     // lanes 0..7 contact ids, 8..15 impulses
#pragma unroll
    for (int i = 0; i < 3; i++) E.pos[i] = S.pos[(size_t)e * 3 + i];
    E.next_step = S.ints[(size_t)e * SS_NINT + SS_I_NEXT_STEP];
}
// bookkeeping scalars (loaded after the physics so they do not occupy registers during it)
__device__ __forceinline__ void env_load_rest(const DevState& S, int e, int lane, Env& E) {
    const double* sc = S.scal + (size_t)e * SS_NSCAL;
    E.potential = sc[SS_S_POTENTIAL];
    E.walk[0] = sc[SS_S_WALK_X]; E.walk[1] = sc[SS_S_WALK_Y]; E.walk[2] = sc[SS_S_WALK_Z];
    E.initial_z = sc[SS_S_INITIAL_Z];
    E.feet_contact[0] = (float)sc[SS_S_FEET_CONTACT0]; E.feet_contact[1] = (float)sc[SS_S_FEET_CONTACT1];
#pragma unroll
    for (int i = 0; i < 6; i++) E.feet_xyz[i / 3][i % 3] = sc[SS_S_FEET_X0 + i];
    E.ep_ret = sc[SS_S_EP_RET];
    E.cache_contact[0] = (float)sc[SS_S_CACHE_CONTACT0]; E.cache_contact[1] = (float)sc[SS_S_CACHE_CONTACT1];
    const int* it = S.ints + (size_t)e * SS_NINT;
    E.next_step = it[SS_I_NEXT_STEP];      // again: not carried in a register through the physics
    E.reached = it[SS_I_REACHED_COUNT]; E.mirrored = it[SS_I_MIRRORED];
    E.elapsed = it[SS_I_ELAPSED]; E.ep_len = it[SS_I_EP_LEN]; E.update_terrain = it[SS_I_UPDATE_TERRAIN];
    E.phantom_idx = it[SS_I_PHANTOM_IDX]; E.started = it[SS_I_STARTED];
    const uint32_t* mt = S.mt + (size_t)e * 2 * SS_MT_WORDS;
    E.mtpos = (int)mt[624]; E.gmtpos = (int)mt[SS_MT_WORDS + 624];
}
__device__ __forceinline__ void env_load(const DevState& S, int e, int lane, Env& E) {
    env_load_phys(S, e, lane, E);
    env_load_rest(S, e, lane, E);
}

__device__ __forceinline__ void env_store(const DevState& S, int e, int lane, const Env& E) {
    float* sv = S.sv + (size_t)e * 64;
    float qh = shfl(E.q_dof, (lane + 2) & 31);      // lane j+4 takes dof 6+j
    float a = 0.f;
    if (lane < 4) a = lane == 0 ? E.quat[0] : (lane == 1 ? E.quat[1] : (lane == 2 ? E.quat[2] : E.quat[3]));
    else if (lane < 4 + NJ) a = qh;
    sv[lane] = a;
    sv[32 + lane] = lane < ND ? E.v_dof : 0.f;
    if (lane < SS_NWARM) S.warm[(size_t)e * SS_NWARM + lane] = E.warm;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) S.pos[(size_t)e * 3 + i] = E.pos[i];
        double* sc = S.scal + (size_t)e * SS_NSCAL;
        sc[SS_S_POTENTIAL] = E.potential;
        sc[SS_S_WALK_X] = E.walk[0]; sc[SS_S_WALK_Y] = E.walk[1]; sc[SS_S_WALK_Z] = E.walk[2];
        sc[SS_S_INITIAL_Z] = E.initial_z;
        sc[SS_S_FEET_CONTACT0] = E.feet_contact[0]; sc[SS_S_FEET_CONTACT1] = E.feet_contact[1];
#pragma unroll
        for (int i = 0; i < 6; i++) sc[SS_S_FEET_X0 + i] = E.feet_xyz[i / 3][i % 3];
        sc[SS_S_EP_RET] = E.ep_ret;
        sc[SS_S_CACHE_CONTACT0] = E.cache_contact[0]; sc[SS_S_CACHE_CONTACT1] = E.cache_contact[1];
        int* it = S.ints + (size_t)e * SS_NINT;
        it[SS_I_NEXT_STEP] = E.next_step; it[SS_I_REACHED_COUNT] = E.reached; it[SS_I_MIRRORED] = E.mirrored;
        it[SS_I_ELAPSED] = E.elapsed; it[SS_I_EP_LEN] = E.ep_len; it[SS_I_UPDATE_TERRAIN] = E.update_terrain;
        it[SS_I_PHANTOM_IDX] = E.phantom_idx; it[SS_I_STARTED] = E.started;
        uint32_t* mt = S.mt + (size_t)e * 2 * SS_MT_WORDS;
        mt[624] = (uint32_t)E.mtpos; mt[SS_MT_WORDS + 624] = (uint32_t)E.gmtpos;
    }
}

// ---------------------------------------------------------------------------------------------
// env logic (each function cites the reference lines it follows; oracle/ss_oracle.c has the scalar twin)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double base_phi(const DevModel& M, int mirrored, int k) {
    // walker3d_envs.py:824-827 / mike.py:345-348
    double b = k == 0 ? -10.0 : (k == M.n_steps - 1 ? 10.0 : ((k & 1) ? 20.0 : -20.0));
    return SS_DEG_D * b * (mirrored ? 1.0 : -1.0);
}
__device__ __forceinline__ double sgn_d(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }
__device__ __forceinline__ double clip_dx(const DevModel& M, double dx) {
    return sgn_d(dx) * fmin(fmax(fabs(dx), M.step_radius * 2.5), M.r_max);
}

// feet link COM positions (bullet_utils.py:94) at the current pose: warp-cooperative FK
__device__ __forceinline__ void feet_positions(const DevModel& M, Env& E, int lane) {
    Xf X = fk_warp(M, E.quat, E.q_dof, lane);
#pragma unroll
    for (int f = 0; f < 2; f++) {
        float t[3]; matvec3(X.R, M.foot_ref[f], t);
        int b = M.foot_body[f];
#pragma unroll
        for (int i = 0; i < 3; i++) E.feet_xyz[f][i] = E.pos[i] + (double)shfl(X.p[i] + t[i], b);
    }
}

// robots.py:304-365 WalkerBase.calc_state.  Writes robot_state[50] into W.vec (clipped to +-5) and returns
// roll/pitch/yaw + joints_at_limit.  Contact flags are the previous step's (q1).
__device__ __forceinline__ void calc_state(const DevModel& M, WarpSmem& W, const Env& E, int lane, float* rpy, int& jal,
                                           bool& finite) {
    quat_to_rpy(E.quat, rpy);
    float ang = 0.f, spd = 0.f; bool at_limit = false;
    if (lane >= 6 && lane < ND) {
        int j = lane - 6;
        float lo = M.q_lo[j], hi = M.q_hi[j];
        ang = 2.f * (E.q_dof - 0.5f * (lo + hi)) / (hi - lo);          // bullet_utils.py:191-194
        spd = 0.1f * E.v_dof;
        at_limit = fabsf(ang) > 0.99f;
    }
    jal = __popc(__ballot_sync(FULL, at_limit));
    float vx0 = shfl(E.v_dof, 3), vy0 = shfl(E.v_dof, 4), vz0 = shfl(E.v_dof, 5);
    float s, c; sincosf(-rpy[2], &s, &c);
    float more[6] = { (float)(E.pos[2] - E.initial_z), c * vx0 - s * vy0, s * vx0 + c * vy0, vz0, rpy[0], rpy[1] };
    auto clip5 = [](float x) { return x < -5.f ? -5.f : (x > 5.f ? 5.f : x); };   // NaN passes through, as np.clip
    __syncwarp();
    if (lane < 6) W.vec[lane] = clip5(lane == 0 ? more[0] : lane == 1 ? more[1] : lane == 2 ? more[2] : lane == 3 ? more[3] : lane == 4 ? more[4] : more[5]);
    else if (lane < ND) { W.vec[lane] = clip5(ang); W.vec[21 + lane] = clip5(spd); }
    else if (lane < ND + 2) W.vec[48 + lane - ND] = clip5(lane == ND ? E.feet_contact[0] : E.feet_contact[1]);
    __syncwarp();
    bool ok = true;
    if (lane < 25) ok = isfinite(W.vec[lane]) && isfinite(W.vec[25 + lane]);
    finite = __all_sync(FULL, ok);
}

// walker3d_envs.py:910-924
__device__ __forceinline__ double calc_potential(const DevModel& M, Env& E, double& dist) {
    double dx = E.walk[0] - E.pos[0], dy = E.walk[1] - E.pos[1];
    dist = sqrt(dx * dx + dy * dy);
    return -dist / M.control_dt;
}

// walker3d_envs.py:798-810
__device__ __forceinline__ void update_steps(const DevModel& M, const DevState& S, int e, const Env& E, int lane) {
    int thr = M.rendered / 2;
    if (E.next_step >= thr) {
        int oldest = ((E.next_step - thr - 1) % M.rendered + M.rendered) % M.rendered;
        int next = min(E.next_step - thr - 1 + M.rendered, M.n_steps - 1);
        if (lane < 6) S.plank[((size_t)e * M.rendered + oldest) * 6 + lane] = S.terrain[((size_t)e * M.n_steps + next) * 6 + lane];
    }
    __syncwarp();
}

// walker3d_envs.py:961-999.  feet_xyz must be current.  Returns target_reached; foot_dist = planar distances.
__device__ __forceinline__ bool calc_feet_state(const DevModel& M, const DevState& S, int e, Env& E, int lane, double* foot_dist) {
    const double* p = S.terrain + ((size_t)e * M.n_steps + E.next_step) * 6;
    double px = p[0], py = p[1], pz = p[2];
    bool reached = false;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        E.feet_contact[i] = E.cache_contact[i] != 0.f ? 1.f : 0.f;
        double dx = E.feet_xyz[i][0] - px, dy = E.feet_xyz[i][1] - py, dz = E.feet_xyz[i][2] - pz;
        foot_dist[i] = sqrt(dx * dx + dy * dy);
        if (E.feet_contact[i] != 0.f && (dx * dx + dy * dy + dz * dz) < 0.3) reached = true;
    }
    if (reached) {
        E.reached++;
        if (E.reached >= M.stop_frames) { E.next_step++; E.reached = 0; update_steps(M, S, e, E, lane); }
        if (E.next_step >= M.n_steps) E.next_step--;
    }
    return reached;
}

// walker3d_envs.py:1046-1077; yaw is the base yaw of the last calc_state.  targets -> W.vec[50..59]
__device__ __forceinline__ void delta_to_k_targets(const DevModel& M, const DevState& S, WarpSmem& W, int e, Env& E, float yaw, int lane) {
    float out[10];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int idx = min(E.next_step + k, M.n_steps - 1);
        const double* t = S.terrain + ((size_t)e * M.n_steps + idx) * 6;
        double tx = t[0], ty = t[1], tz = t[2];
        if (k == 1) { E.walk[0] = tx; E.walk[1] = ty; E.walk[2] = tz; }          // q8
        float dx = (float)(tx - E.pos[0]), dy = (float)(ty - E.pos[1]), dz = (float)(tz - E.pos[2]);
        float ang = atan2f(dy, dx) - yaw, dist = sqrtf(dx * dx + dy * dy), s, c;
        sincosf(ang, &s, &c);
        out[5 * k] = s * dist; out[5 * k + 1] = c * dist; out[5 * k + 2] = dz; out[5 * k + 3] = (float)t[4]; out[5 * k + 4] = (float)t[5];
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 10; i++) W.vec[50 + i] = out[i];
    }
    __syncwarp();
}

// walker3d_envs.py:1245-1274 / mike.py:755-784: the row set_next_next_step_location writes, from the next stone's row nx
__device__ __forceinline__ void next_next_row(const DevModel& M, const double* nx, double bphi, double sp, double cp, double sy,
                                              double cy, double yaw, double dr, double xt, double yt, double* v) {
    double byaw = nx[3], sb, cb;
    sincos(byaw, &sb, &cb);
    double dx = clip_dx(M, dr * sp * cy), dy = dr * sp * sy;
    v[0] = nx[0] + cb * dx - sb * dy; v[1] = nx[1] + sb * dx + cb * dy; v[2] = nx[2] + dr * cp; v[3] = yaw + byaw; v[4] = xt; v[5] = yt;
    (void)bphi;
}
__device__ __forceinline__ void set_next_next_step_location(const DevModel& M, const DevState& S, int e, const Env& E, int lane,
                                                            double pitch, double yaw, double dr, double xt, double yt) {
    const double* nx = S.terrain + ((size_t)e * M.n_steps + E.next_step) * 6;
    int bci = (E.next_step + 1) % M.n_steps;
    double bphi = base_phi(M, E.mirrored, bci);
    double sp, cp, sy, cy;
    sincos(pitch, &sp, &cp); sincos(yaw + bphi, &sy, &cy);
    double v[6];
    next_next_row(M, nx, bphi, sp, cp, sy, cy, yaw, dr, xt, yt, v);
    __syncwarp();
    if (lane == 0) {
        double* t = S.terrain + ((size_t)e * M.n_steps + bci) * 6;
#pragma unroll
        for (int i = 0; i < 6; i++) t[i] = v[i];
    }
    __syncwarp();
}

// mike.py:660-692: RandomState.choice(p, size=1, replace=False) == one uniform + searchsorted(cdf, u, 'right')
__device__ __forceinline__ void sample_next_next_step(const DevModel& M, const DevState& S, WarpSmem& W, int e, Env& E, int lane) {
    uint32_t* mt = S.mt + (size_t)e * 2 * SS_MT_WORDS;
    double u = mt_one_double_warp(mt, E.mtpos, W.u.d, lane);
    int lo = 0, hi = S.n_cells;
    const double* cdf = S.cdf + (size_t)S.tslot[e] * S.n_cells;      // this env's table version (ss_set_sample_cdf)
    while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
    int idx = lo < S.n_cells ? lo : S.n_cells - 1;
    const int g = SS_GRID;
    int i4 = idx % g, i3 = (idx / g) % g, i2 = (idx / (g * g)) % g, i1 = (idx / (g * g * g)) % g, i0 = idx / (g * g * g * g);
    set_next_next_step_location(M, S, e, E, lane, M.samples[1][i1] + SS_PI_D / 2, M.samples[0][i0], M.samples[2][i2],
                                M.samples[3][i3], M.samples[4][i4]);
}

// walker3d_envs.py:1079-1093 / mike.py:589-603
__device__ __forceinline__ void update_terrain_info(const DevModel& M, const DevState& S, WarpSmem& W, int e, Env& E, float yaw, int lane) {
    int nn = E.next_step + 1;
    if (M.use_sampler) sample_next_next_step(M, S, W, e, E, lane);
    if (lane < 6) S.plank[((size_t)e * M.rendered + nn % M.rendered) * 6 + lane] = S.terrain[((size_t)e * M.n_steps + nn % M.n_steps) * 6 + lane];
    __syncwarp();
    delta_to_k_targets(M, S, W, e, E, yaw, lane);
}

// walker3d_envs.py:624-755 (W3D) / mike.py:167-276 (Mike).  Scratch: W.u.d (>= 7 * n_steps doubles).
// Only dr is a genuine random array: the yaw / pitch / tilt draws are `uniform(a, a)` = a + 0 * u = a exactly, so they are
// consumed from the stream (the seed-exact draw order matters) but not stored.
__device__ __forceinline__ void generate_step_placements(const DevModel& M, const DevState& S, WarpSmem& W, int e, Env& E, int lane) {
    const int n = M.n_steps;
    double* dr = W.u.d; double *cphi = dr + n, *X = dr + 2 * n, *Y = dr + 3 * n, *Z = dr + 4 * n, *U = dr + 5 * n, *PHI = dr + 6 * n;
    uint32_t* mt = S.mt + (size_t)e * 2 * SS_MT_WORDS;
    double* T = S.terrain + (size_t)e * n * 6;
    const bool mike = M.terrain_mode == 1;
    const double dr_lo = 0.65, dr_hi = mike ? 0.65 : 0.8, y0 = mike ? -10 * SS_DEG_D : 0.0, t0 = mike ? -20 * SS_DEG_D : 0.0;
    auto dphi_of = [&](int i) { return i < 3 ? 0.0 : y0; };
    auto dth_of = [&](int i) { return i < 3 ? SS_PI_D / 2 : 90 * SS_DEG_D; };
    __syncwarp();
    mt_doubles_warp(mt, E.mtpos, dr, n, lane);
    mt_doubles_warp(mt, E.mtpos, nullptr, 4 * n, lane);          // dphi, dtheta, x_tilt, y_tilt draws
    for (int i = lane; i < n; i += 32) {
        dr[i] = dr_lo + (dr_hi - dr_lo) * dr[i];
        if (i < 3) dr[i] = i == 0 ? 0.0 : (i == 1 ? 0.8 : 0.75);
    }
    __syncwarp();
    if (lane == 0) { double acc = 0; for (int i = 0; i < n; i++) { acc += dphi_of(i); cphi[i] = acc; } }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        double bp = base_phi(M, E.mirrored, i), st, ct, sp, cp;
        sincos(dth_of(i), &st, &ct); sincos(cphi[i] + bp, &sp, &cp);
        double x_ = dr[i] * st * cp;
        if (i >= 2) x_ = clip_dx(M, x_);
        X[i] = x_; Y[i] = dr[i] * st * sp; Z[i] = dr[i] * ct;
    }
    __syncwarp();
    const double min_z = M.step_radius * sin(M.tilt_limit_deg * SS_DEG_D) + 0.01;
    if (lane == 0) {
        double ax = 0, ay = 0, az = 0;
        for (int i = 0; i < n; i++) {
            ax += X[i]; ay += Y[i]; az += Z[i];
            double* t = T + 6 * i;
            t[0] = ax; t[1] = ay; t[2] = fmax(az + 0, min_z); t[3] = cphi[i]; t[4] = i < 3 ? 0.0 : t0; t[5] = 0.0;
        }
    }
    __syncwarp();
    // second pass: stones 3..n-1 are re-placed relative to their predecessor (walker3d_envs.py:700-755)
    if (!mike) mt_doubles_warp(mt + SS_MT_WORDS, E.gmtpos, U + 2, n - 3, lane);      // global numpy stream (q7)
    for (int i = 2 + lane; i < n - 1; i += 32) {
        int bci = i + 1;
        double bphi = base_phi(M, E.mirrored, bci);
        double pitch = dth_of(bci), yaw = dphi_of(bci);
        if (!mike) {
            pitch = (i % 2 == 0) ? 120 * SS_DEG_D : 60 * SS_DEG_D;
            if (U[i] < 0.2) pitch = 90 * SS_DEG_D;
            if (M.terrain_mode == 2) pitch = 90 * SS_DEG_D;
        }
        double sp, cp, sy, cy;
        sincos(pitch, &sp, &cp); sincos(yaw + bphi, &sy, &cy);
        X[i] = clip_dx(M, dr[bci] * sp * cy); Y[i] = dr[bci] * sp * sy; Z[i] = dr[bci] * cp;
    }
    __syncwarp();
    if (lane == 0) { PHI[2] = cphi[2]; for (int i = 2; i < n - 1; i++) PHI[i + 1] = dphi_of(i + 1) + PHI[i]; }
    __syncwarp();
    for (int i = 2 + lane; i < n - 1; i += 32) {
        double sb, cb; sincos(PHI[i], &sb, &cb);
        double dx = X[i], dy = Y[i];
        X[i] = cb * dx + -sb * dy; Y[i] = sb * dx + cb * dy;
    }
    __syncwarp();
    if (lane == 0) {
        double x = T[12], y = T[13], z = T[14];
        for (int i = 2; i < n - 1; i++) {
            x = x + X[i]; y = y + Y[i]; z = z + Z[i];
            double* t = T + 6 * (i + 1);
            t[0] = x; t[1] = y; t[2] = z; t[3] = PHI[i + 1];
            if (mike) t[4] = -20 * SS_DEG_D;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void write_obs(const WarpSmem& W, const Env& E, float* obs, int lane) {
    float h = (float)(E.pos[2] - fmin(E.feet_xyz[0][2], E.feet_xyz[1][2]));                   // q6
    float a = W.vec[lane], b = lane < 28 ? W.vec[32 + lane] : 0.f;
    if (lane == 0) a = h;
    obs[lane] = a;
    if (lane < 28) obs[32 + lane] = b;
}

// Walker3DStepperEnv.reset walker3d_envs.py:812-867 / mike.py:333-388 + Walker3D.reset robots.py:581-618
__device__ __forceinline__ void env_reset(const DevModel& M, const DevState& S, WarpSmem& W, int e, Env& E, float* obs, int lane) {
    uint32_t* mt = S.mt + (size_t)e * 2 * SS_MT_WORDS;
    E.reached = 0;
    // robots.py:583 / :846: the robot's np_random is NOT the env's seeded stream (oracle env_reset explains): the
    // mirror coin comes from the per-env aux stream; Walker3D's `< 0` test is unobservable and draws nothing.
    E.mirrored = 0;
    if (M.mirror_prob > 0.0) { double u = mt_one_double_warp(mt + SS_MT_WORDS, E.gmtpos, W.u.d, lane); E.mirrored = u < M.mirror_prob; }
    int j = lane - 6, src = j;
    if (E.mirrored) {
        if ((j >= 3 && j <= 7) || (j >= 13 && j <= 16)) src = j + (j <= 7 ? 5 : 4);
        else if ((j >= 8 && j <= 12) || (j >= 17 && j <= 20)) src = j - (j <= 12 ? 5 : 4);
    }
    float q = 0.f;
    if (j >= 0 && j < NJ) { q = M.init_q[src]; if (E.mirrored && (j == 0 || j == 2)) q = -q; }
    E.q_dof = q; E.v_dof = 0.f;           // resetBasePositionAndOrientation zeroes both base velocities (oracle env_reset)
    {   // the warm-start cache survives resets (q4); a NaN must not
        float other = shfl(E.warm, lane ^ SS_MAX_CONTACTS);
        if (!isfinite(E.warm) || (lane < SS_NWARM && !isfinite(other))) E.warm = lane < SS_MAX_CONTACTS ? -1.f : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) E.pos[i] = M.init_pos[i];
    E.quat[0] = E.quat[1] = E.quat[2] = 0.f; E.quat[3] = 1.f;
    // WalkerBase.reset robots.py:433-439
    E.feet_contact[0] = E.feet_contact[1] = 0.f;
    E.initial_z = E.pos[2];
    float rpy[3]; int jal; bool fin;
    calc_state(M, W, E, lane, rpy, jal, fin);
    // stale-state calc_feet_state (q2)
    feet_positions(M, E, lane);
    double fd[2];
    calc_feet_state(M, S, e, E, lane, fd);
    generate_step_placements(M, S, W, e, E, lane);
    for (int i = lane; i < M.rendered * 6; i += 32) S.plank[(size_t)e * M.rendered * 6 + i] = S.terrain[(size_t)e * M.n_steps * 6 + i];
    __syncwarp();
    E.next_step = 1;
    delta_to_k_targets(M, S, W, e, E, rpy[2], lane);
    double dist; E.potential = calc_potential(M, E, dist);
    write_obs(W, E, obs, lane);
    E.phantom_idx = 0;            // walker3d_envs.py:859-862: the phantoms draw from the robot stream (visual only)
    // update_terrain keeps the terminal step's value: the reference's reset() never touches it (walker3d_envs.py:812-867,
    // mike.py:333-388), so update_sample_prob / create_temp_states right after a done-and-advance step act on the reset env
    E.elapsed = 0; E.ep_ret = 0; E.ep_len = 0; E.started = 1;
}

// Cold path, one out-of-line copy shared by the step and reset kernels: reload the env from memory, reset, store.
__device__ __noinline__ void env_reset_from_memory(const DevModel* Mp, const DevState* Sp, WarpSmem* Wp, int e, float* obs, int lane) {
    Env E; env_load(*Sp, e, lane, E);
    env_reset(*Mp, *Sp, *Wp, e, E, obs, lane);
    env_store(*Sp, e, lane, E);
}

// stone boxes of the 4 candidate slots around next_step, relative to the step-start base position
__device__ __forceinline__ void setup_stones(const DevModel& M, const DevState& S, WarpSmem& W, int e, const Env& E, int lane) {
    if (lane < 4) {
        int k = -1;
        for (int kk = E.next_step - 2; kk <= E.next_step + 1; kk++) if (kk >= 0 && (kk & 3) == lane) k = kk;
        float* st = W.stone[lane];
        if (k < 0 || k > M.n_steps - 1) st[12] = 0.f;
        else {
            const double* pl = S.plank + ((size_t)e * M.rendered + k % M.rendered) * 6;
            // getQuaternionFromEuler([x_tilt, y_tilt, phi]) : Rz(yaw) Ry(pitch) Rx(roll); bullet_objects.py:75-80 (q12)
            double sr, cr, sp, cp, sy, cy;
            sincos(pl[4], &sr, &cr); sincos(pl[5], &sp, &cp); sincos(pl[3], &sy, &cy);
            double R[9] = { cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                            -sp, cp * sr, cp * cr };
            double off = M.stone_box_off;
            st[0] = (float)(pl[0] + R[2] * off - E.pos[0]);
            st[1] = (float)(pl[1] + R[5] * off - E.pos[1]);
            st[2] = (float)(pl[2] - (double)M.stone_com_z + R[8] * off - E.pos[2]);
#pragma unroll
            for (int i = 0; i < 9; i++) st[3 + i] = (float)R[i];
            st[12] = 1.f;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char smem_raw[];

// DIAG = per-substep trace / debug dump requested (parity tests): a second instantiation, so that the production kernel
// carries neither the code nor the registers of the probes.
// S is __grid_constant__: its address is handed to the out-of-line reset path, and taking the address of a plain by-value
// parameter makes every thread copy the whole struct to its local-memory frame in the prologue (8 STL.128 per thread
// = 33 MB of DRAM writes per 8192-env launch once the dirty lines are evicted: profiles/r01l traffic).
template <bool DIAG>
__global__ void __launch_bounds__(BLOCK, SS_MINB) env_step_kernel(const DevModel* gm, const __grid_constant__ DevState S, const float* __restrict__ act,
                                                         float* __restrict__ obs, float* __restrict__ rew,
                                                         uint8_t* __restrict__ done, uint8_t* __restrict__ timeout,
                                                         float* __restrict__ ep_ret, int32_t* __restrict__ ep_len, int physics,
                                                         int e_begin, int e_end, double* __restrict__ rew64) {
    // envs [e_begin, e_end) of the batch.  rew64 != null (host-buffer entry point): every output pointer is PINNED HOST
    // memory the kernel writes straight into (zero-copy over PCIe, posted stores) -- the observation row, and the scalars in
    // the caller's own types (the reward as float64, like the reference's stacked python floats).
    BlockSmem& B = *reinterpret_cast<BlockSmem*>(smem_raw);
    stage_model_issue(B, gm);                  // the env's state row is fetched while the tables are in flight
    const DevModel& M = B.M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e_raw = e_begin + blockIdx.x * WARPS_PER_BLOCK + warp;
#if SS_SYNC
    const bool ghost = e_raw >= e_end;          // keeps hitting the block barriers, stores nothing
    // the k-th warp takes the k-th env in order of contact count (sort_envs_kernel): the warps of a CTA then carry
    // similar solver loads and wait less for each other at the per-substep barrier.  Which warp steps which env has no
    // effect on the results.
    const int e = S.perm[ghost ? e_end - 1 : e_raw];
#else
    if (e_raw >= e_end) return;
    const bool ghost = false;
    const int e = e_raw;
#endif
    WarpSmem& W = B.W[warp];
    Env E; env_load_phys(S, e, lane, E);
    float a_in = 0.f;
    if (lane >= 6 && lane < ND) a_in = act[(size_t)e * NJ + lane - 6];
    // robots.py:290 `assert np.isfinite(a).all()`: the host-buffer entry point refuses such a batch; on the device path
    // there is nobody to raise to, so the env treats the whole action as zero (a NaN would otherwise clamp to -1 = full
    // negative torque and poison reward, GAE and the update) and raises the sticky flag ss_error_flags reports.
    if (__any_sync(FULL, !isfinite(a_in))) { a_in = 0.f; if (lane == 0 && !ghost) atomicOr(S.err, 1); }
    stage_model_wait(B);

    // robots.py:289-293 apply_action (torque held over the control step) + MJCF joint damping torque.
    // actuator_mode 1: the reference's abandoned PD block (robots.py:486-493), once per control step like it;
    // 2: the same law re-evaluated at every substep ("per-substep PD joint torques", BASELINE north star).
    float tau = 0.f;
    if (lane >= 6 && lane < ND) {
        int j = lane - 6;
        float x = fminf(fmaxf(a_in, -1.f), 1.f);
        if (M.actuator_mode == 0) tau = M.tau_max[j] * x;
        else {                                                                     // robots.py:486-493 (opt-in PD)
            float lo = M.q_lo[j], hi = M.q_hi[j];
            tau = M.pd_p[j] * (x - 2.f * (E.q_dof - 0.5f * (lo + hi)) / (hi - lo)) - M.pd_d[j] * E.v_dof;
        }
        const float damp0 = M.damping[j] * E.v_dof;
        tau -= damp0;
        if (M.actuator_mode == 2) { W.pd[0][lane] = x; W.pd[1][lane] = damp0; }
    }
    int cflags = -1;
    if (physics == 1) {      // 0: env logic only (ss_env_post); 2: the wrappers too, physics replaced by the injected state (ss_step_injected)
        setup_stones(M, S, W, e, E, lane);
        // the double-precision base position is not needed inside the substeps (they integrate the displacement dpos):
        // park it in shared memory so that it does not hold six registers through the solver
        if (lane < 3) W.parked[lane] = lane == 0 ? E.pos[0] : (lane == 1 ? E.pos[1] : E.pos[2]);     // (no dynamic index: E stays in registers)
        float dpos[3] = { 0.f, 0.f, 0.f };
        for (int s = 0; s < M.n_substeps; s++) {
#if SS_SYNC
            __syncthreads();                     // re-align the CTA's warps so that they share instruction fetches
#endif
            if (M.actuator_mode == 2 && s > 0 && lane >= 6 && lane < ND) {
                const int j = lane - 6;
                const float lo = M.q_lo[j], hi = M.q_hi[j];
                tau = M.pd_p[j] * (W.pd[0][lane] - 2.f * (E.q_dof - 0.5f * (lo + hi)) / (hi - lo)) - M.pd_d[j] * E.v_dof - W.pd[1][lane];
            }
            float* dbg = (DIAG && S.dbg && s == M.n_substeps - 1) ? S.dbg + (size_t)e * DBG_FLOATS : nullptr;
            unsigned ev_lo = 0, ev_hi = 0;
            if (DIAG && S.trace_ev) {            // the limit rows this substep is about to create (same test as substep_warp)
                int lv = 0;
                if (lane >= 6 && lane < ND) lv = (E.q_dof - M.q_lo[lane - 6] <= 0.f) ? 1 : ((M.q_hi[lane - 6] - E.q_dof <= 0.f) ? 2 : 0);
                const unsigned any = __ballot_sync(FULL, lv != 0);
                const bool kept = lv != 0 && __popc(any & ((1u << lane) - 1u)) < MAXL;
                ev_lo = __ballot_sync(FULL, kept && lv == 1) >> 6; ev_hi = __ballot_sync(FULL, kept && lv == 2) >> 6;
            }
            bool clamp_hit = false; unsigned push_hist = 0;
            cflags = substep_warp<DIAG>(M, W, E.quat, E.q_dof, E.v_dof, tau, dpos, E.warm, lane, dbg, DIAG ? &clamp_hit : nullptr,
                                        DIAG ? &push_hist : nullptr);
            if (DIAG && S.trace_ev && s < 4 && !ghost) {
                int* ev = S.trace_ev + ((size_t)s * S.n + e) * SS_NEV;
                if (lane < MAXC) {                 // contact ids of this substep (-1 = none) | geometry-branch code << 16
                    const int id = (int)E.warm;
                    ev[2 + lane] = id < 0 ? -1 : (id | (((int)W.con[lane][14] | (fabsf(W.con[lane][2]) > 0.70710678f ? 1 : 0)) << 16));
                    ev[10 + lane] = id < 0 ? 0 : (int)push_hist;     // iterations in which this contact's normal impulse was > 0
                }
                if (lane == 0) { ev[0] = (int)ev_lo; ev[1] = (int)(ev_hi | (clamp_hit ? 1u << 30 : 0u)); }   // bit 30: max_coord_vel clamp engaged
            }
            if (DIAG && S.trace_q && s < 4 && !ghost) {
                double* tq = S.trace_q + ((size_t)s * S.n + e) * SS_NQ;
                double* tv = S.trace_qd + ((size_t)s * S.n + e) * ND;
                if (lane < 3) tq[lane] = E.pos[lane] + (double)(lane == 0 ? dpos[0] : lane == 1 ? dpos[1] : dpos[2]);
                if (lane < 4) tq[3 + lane] = lane == 0 ? E.quat[0] : lane == 1 ? E.quat[1] : lane == 2 ? E.quat[2] : E.quat[3];
                if (lane >= 6 && lane < ND) tq[1 + lane] = E.q_dof;
                if (lane < ND) tv[lane] = E.v_dof;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; i++) E.pos[i] = W.parked[i] + (double)dpos[i];
    }
    if (ghost) return;
    env_load_rest(S, e, lane, E);
    if (cflags >= 0) { E.cache_contact[0] = (cflags & 1) ? 1.f : 0.f; E.cache_contact[1] = (cflags & 2) ? 1.f : 0.f; }

    // ---- robot.calc_state + calc_env_state (walker3d_envs.py:875-903, 1026-1044) ----
    float rpy[3]; int jal; bool finite;
    calc_state(M, W, E, lane, rpy, jal, finite);
    bool dn = !finite;
    const int cur = E.next_step;
    feet_positions(M, E, lane);
    double fd[2];
    bool reached = calc_feet_state(M, S, e, E, lane, fd);
    // calc_base_reward walker3d_envs.py:926-959
    double old_pot = E.potential, dist;
    E.potential = calc_potential(M, E, dist);
    float progress = (float)(E.potential - old_pot);
    float posture = 0.f;
    if (!(-0.2f < rpy[1] && rpy[1] < 0.4f)) posture = fabsf(rpy[1]);
    if (!(-0.4f < rpy[0] && rpy[0] < 0.4f)) posture += fabsf(rpy[0]);
    float m1 = 0.f, m2 = 0.f, a_again = 0.f;
    if (lane >= 6 && lane < ND) {                                                               // unclipped action (q5)
        a_again = act[(size_t)e * NJ + lane - 6];      // re-read (L1/L2 hit): keeps a register free through the physics
    }
    if (__any_sync(FULL, !isfinite(a_again))) a_again = 0.f;           // non-finite action: stepped as zero (see the top)
    m1 = fabsf(a_again * 0.1f * E.v_dof); m2 = a_again * a_again;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { m1 += __shfl_xor_sync(FULL, m1, off); m2 += __shfl_xor_sync(FULL, m2, off); }
    float energy = M.electricity_cost * (m1 / NJ) + M.stall_torque_cost * (m2 / NJ);
    float joints_pen = M.joints_at_limit_cost * (float)jal;
    float height = (float)(E.pos[2] - fmin(E.feet_xyz[0][2], E.feet_xyz[1][2]));
    float tall = height > M.done_height ? 2.f : -1.f;
    dn = dn || tall < 0.f;
    // calc_step_reward walker3d_envs.py:1001-1024
    float step_bonus = 0.f;
    if (reached && E.reached == 1) {
        step_bonus = 50.f * expf(-(float)fmin(fd[0], fd[1]) / 0.25f);
        if (M.n_phantoms > 0 && E.next_step % 3 == 0) E.phantom_idx = (E.phantom_idx + 1) % 20;   // walker3d_envs.py:1007-1016
    }
    float target_bonus = (E.next_step == M.n_steps - 1 && dist < 0.15) ? 2.f : 0.f;
    delta_to_k_targets(M, S, W, e, E, rpy[2], lane);
    E.update_terrain = cur != E.next_step;
    if (cur != E.next_step) { update_terrain_info(M, S, W, e, E, rpy[2], lane); double d2; E.potential = calc_potential(M, E, d2); }
    float r = progress - energy;
    r += step_bonus + target_bonus - 0.f;
    r += tall - posture - joints_pen;
    float* o = obs + (size_t)e * SS_OBS;
    write_obs(W, E, o, lane);
    if (!physics) {
        if (lane == 0) { rew[e] = r; done[e] = dn ? 1 : 0; }
        env_store(S, e, lane, E);
        return;
    }
    // gym TimeLimit + TimeLimitMask envs_utils.py:59-69 + Monitor :130-153
    E.elapsed++;
    bool d = dn || E.elapsed >= M.max_episode_steps;
    E.ep_ret += (double)r; E.ep_len++;
    if (lane == 0) {
        const int to = (d && E.elapsed == M.max_episode_steps) ? 1 : 0;
        const float er = (float)(rint(E.ep_ret * 1e6) / 1e6);
        if (rew64) rew64[e] = (double)r; else rew[e] = r;
        done[e] = d ? 1 : 0; timeout[e] = to; ep_ret[e] = er; ep_len[e] = E.ep_len;
    }
    env_store(S, e, lane, E);
    if (d) { __syncwarp(); env_reset_from_memory(&M, &S, &W, e, o, lane); }         // envs_utils.py:646-649
}

__global__ void __launch_bounds__(BLOCK) env_reset_kernel(const DevModel* gm, const __grid_constant__ DevState S, float* __restrict__ obs) {
    BlockSmem& B = *reinterpret_cast<BlockSmem*>(smem_raw);
    stage_model(B, gm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * WARPS_PER_BLOCK + warp;
    if (e >= S.n) return;
    env_reset_from_memory(&B.M, &S, &B.W[warp], e, obs + (size_t)e * SS_OBS, lane);
}

// update_sample_prob's immediate re-sample (walker3d_envs.py:1317-1320): envs whose update_terrain flag is set
__global__ void __launch_bounds__(BLOCK) resample_kernel(const DevModel* gm, DevState S, int slot) {
    BlockSmem& B = *reinterpret_cast<BlockSmem*>(smem_raw);
    stage_model(B, gm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * WARPS_PER_BLOCK + warp;
    if (e >= S.n) return;
    Env E; env_load(S, e, lane, E);
    if (E.update_terrain) {                  // mike.py:827-830: only an env that advanced a stone takes the table it is sent
        if (lane == 0) S.tslot[e] = slot;
        __syncwarp();
        float rpy[3]; quat_to_rpy(E.quat, rpy);
        update_terrain_info(B.M, S, B.W[warp], e, E, rpy[2], lane);
        env_store(S, e, lane, E);
    }
}

// ---------------------------------------------------------------------------------------------
// create_temp_states (walker3d_envs.py:1189-1243 / mike.py:699-753): value-guided sampling probes.
//   temp_prep_kernel  warp per env: the env's current observation row (get_state + both targets), the context the
//                     per-cell kernel needs, and the reference's side effects (walk_target left at the LAST grid
//                     cell; the next-next stone "restored" from next_next_* -- a no-op for Mike, a flattening for W3D)
//   temp_fill_kernel  block per 256 grid cells of one env: only observation columns 55..59 (the second target)
//                     depend on the cell; rows are assembled in shared memory and streamed out as float4.
//                     HBM-write bound: 240 B per (env, cell).
// ---------------------------------------------------------------------------------------------
struct TempGrid { double a[5][SS_GRID]; double restore[5]; int variant, has_restore, n3, n45; };
constexpr int TS_CTX = 12;         // doubles per env: flag, pos xyz, yaw, nx (x y z phi), bphi, modifies-target-1, pad
constexpr int TS_CELLS = 256;      // cells per block of temp_fill_kernel

__global__ void __launch_bounds__(BLOCK) temp_prep_kernel(const DevModel* gm, DevState S, TempGrid G, float* __restrict__ base,
                                                          double* __restrict__ ctx) {
    BlockSmem& B = *reinterpret_cast<BlockSmem*>(smem_raw);
    stage_model(B, gm);
    const DevModel& M = B.M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * WARPS_PER_BLOCK + warp;
    if (e >= S.n) return;
    WarpSmem& W = B.W[warp];
    Env E; env_load(S, e, lane, E);
    double* c = ctx + (size_t)e * TS_CTX;
    if (!E.update_terrain) { if (lane == 0) c[0] = 0.0; return; }
    float rpy[3]; int jal; bool fin;
    calc_state(M, W, E, lane, rpy, jal, fin);                    // get_state: robot.calc_state() on the current pose
    delta_to_k_targets(M, S, W, e, E, rpy[2], lane);
    write_obs(W, E, base + (size_t)e * 64, lane);
    const int bci = (E.next_step + 1) % M.n_steps, idx1 = min(E.next_step + 1, M.n_steps - 1);
    const double* nx = S.terrain + ((size_t)e * M.n_steps + E.next_step) * 6;
    const double bphi = base_phi(M, E.mirrored, bci);
    if (lane == 0) {
        c[0] = 1.0; c[1] = E.pos[0]; c[2] = E.pos[1]; c[3] = E.pos[2]; c[4] = (double)rpy[2];
        c[5] = nx[0]; c[6] = nx[1]; c[7] = nx[2]; c[8] = nx[3]; c[9] = bphi; c[10] = idx1 == bci ? 1.0 : 0.0;
    }
    // side effects: the last get_temp_state leaves walk_target at the last cell's stone (if that stone is target 1)
    if (idx1 == bci) {
        const int g = SS_GRID - 1;
        double r = G.variant == 1 ? 0.7 : G.a[2][g], xt = G.variant == 5 ? G.a[3][g] : 0.0, yt = G.variant == 5 ? G.a[4][g] : 0.0;
        double sp, cp, sy, cy, v[6];
        sincos(SS_PI_D / 2 - G.a[1][g], &sp, &cp); sincos(G.a[0][g] + bphi, &sy, &cy);
        next_next_row(M, nx, bphi, sp, cp, sy, cy, G.a[0][g], r, xt, yt, v);
        E.walk[0] = v[0]; E.walk[1] = v[1]; E.walk[2] = v[2];
    }
    if (G.has_restore) set_next_next_step_location(M, S, e, E, lane, G.restore[0], G.restore[1], G.restore[2], G.restore[3], G.restore[4]);
    env_store(S, e, lane, E);
}

__global__ void __launch_bounds__(TS_CELLS) temp_fill_kernel(const DevModel* gm, TempGrid G, const float* __restrict__ base,
                                                             const double* __restrict__ ctx, float* __restrict__ out, int cells) {
    __shared__ __align__(16) float row[64];
    __shared__ float cellv[TS_CELLS][5];
    __shared__ double tab[4][SS_GRID];       // sin/cos(pi/2 - pitch), sin/cos(yaw + base_phi)
    const int e = blockIdx.y, c0 = blockIdx.x * TS_CELLS, t = threadIdx.x;
    const double* c = ctx + (size_t)e * TS_CTX;
    const bool on = c[0] != 0.0, mod = c[10] != 0.0;
    const int ncell = min(TS_CELLS, cells - c0);
    float4* o4 = reinterpret_cast<float4*>(out + ((size_t)e * cells + c0) * SS_OBS);
    const int nv = ncell * (SS_OBS / 4);
    if (!on) {                                // self.temp_states: zeros
        for (int v = t; v < nv; v += TS_CELLS) __stcs(o4 + v, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    if (t < 64) row[t] = t < SS_OBS ? base[(size_t)e * 64 + t] : 0.f;
    if (t >= 64 && t < 64 + SS_GRID) { int i = t - 64; sincos(SS_PI_D / 2 - G.a[1][i], &tab[0][i], &tab[1][i]); }
    if (t >= 96 && t < 96 + SS_GRID) { int i = t - 96; sincos(G.a[0][i] + c[9], &tab[2][i], &tab[3][i]); }
    __syncthreads();
    if (t < ncell) {
        float* cv = cellv[t];
        if (!mod) {
#pragma unroll
            for (int i = 0; i < 5; i++) cv[i] = row[55 + i];
        } else {
            int idx = c0 + t;
            int it = idx % G.n45; idx /= G.n45;
            int ix = idx % G.n45; idx /= G.n45;
            int ir = idx % G.n3; idx /= G.n3;
            int ip = idx % SS_GRID, iy = idx / SS_GRID;
            double r = G.variant == 1 ? 0.7 : G.a[2][ir], xt = G.variant == 5 ? G.a[3][ix] : 0.0, yt = G.variant == 5 ? G.a[4][it] : 0.0;
            const double sp = tab[0][ip], cp = tab[1][ip], sy = tab[2][iy], cy = tab[3][iy];
            double sb, cb; sincos(c[8], &sb, &cb);
            const DevModel& M = *gm;
            double dx = clip_dx(M, r * sp * cy), dy = r * sp * sy;
            double x = c[5] + cb * dx - sb * dy, y = c[6] + sb * dx + cb * dy, z = c[7] + r * cp;
            float fx = (float)(x - c[1]), fy = (float)(y - c[2]), fz = (float)(z - c[3]);
            float ang = atan2f(fy, fx) - (float)c[4], dist = sqrtf(fx * fx + fy * fy), s, co;
            sincosf(ang, &s, &co);
            cv[0] = s * dist; cv[1] = co * dist; cv[2] = fz; cv[3] = (float)xt; cv[4] = (float)yt;
        }
    }
    __syncthreads();
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int v = t; v < nv; v += TS_CELLS) {
        int cell = v / 15, k = v - cell * 15;
        float4 val;
        if (k < 13) val = r4[k];
        else if (k == 13) val = make_float4(row[52], row[53], row[54], cellv[cell][0]);
        else val = make_float4(cellv[cell][1], cellv[cell][2], cellv[cell][3], cellv[cell][4]);
        __stcs(o4 + v, val);
    }
}

// sampling-table versions: from < 0: every env takes version `to` (update_curriculum / update_specialist); else the envs on
// version `from` move to `to` (the ring is about to overwrite `from`)
__global__ void table_assign_kernel(int* __restrict__ tslot, int n, int from, int to) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n && (from < 0 || tslot[e] == from)) tslot[e] = to;
}

// Order the envs of [e_begin, e_end) by the number of contacts they carried in their last substep (the warm-start list),
// envs with a link-link contact first, then most contacts first (the last, partial wave of the step kernel then gets the
// light envs): a counting sort with 18 bins.
//   sort_hist_kernel     key per env + bin totals (block histogram in shared memory, one global atomic per bin and block)
//   sort_scatter_kernel  perm[base[key] + position] = env; positions from warp-aggregated atomics (the order inside a
//                        bin is arbitrary -- it only decides which warp steps which env, never a result)
// bins: int[64] per sort slot, zeroed by a cudaMemsetAsync in front: [0..17] totals, [32..49] cursors.
constexpr int SORT_THREADS = 256;
constexpr int SORT_BINS = 2 * (SS_MAX_CONTACTS + 1);      // key = contacts + 9 * (one of them is a link-link contact)
constexpr int SORT_SLOT = 64;                              // ints per sort slot: totals [0..31], cursors [32..63]
__device__ __forceinline__ int sort_key(const float* wm) {
    const float4 a = reinterpret_cast<const float4*>(wm)[0], b = reinterpret_cast<const float4*>(wm)[1];
    const int c = (a.x >= 0.f) + (a.y >= 0.f) + (a.z >= 0.f) + (a.w >= 0.f) + (b.x >= 0.f) + (b.y >= 0.f) + (b.z >= 0.f) + (b.w >= 0.f);
    // link-link contacts (ids >= SS_NGEOM * 8) cost a second substitution pass in the row build: keep those envs together
    const float m = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    return c + (m >= (float)(SS_NGEOM * 8) ? SS_MAX_CONTACTS + 1 : 0);
}
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(DevState S, unsigned char* __restrict__ keys, int* __restrict__ bins,
                                                                 int e_begin, int e_end) {
    __shared__ int h[SORT_BINS];
    if (threadIdx.x < SORT_BINS) h[threadIdx.x] = 0;
    __syncthreads();
    const int e = e_begin + blockIdx.x * SORT_THREADS + threadIdx.x;
    if (e < e_end) {
        const int c = sort_key(S.warm + (size_t)e * SS_NWARM);
        keys[e] = (unsigned char)c;
        atomicAdd(&h[c], 1);
    }
    __syncthreads();
    if (threadIdx.x < SORT_BINS && h[threadIdx.x]) atomicAdd(&bins[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(DevState S, const unsigned char* __restrict__ keys, int* __restrict__ bins,
                                                                    int e_begin, int e_end) {
    const int e = e_begin + blockIdx.x * SORT_THREADS + threadIdx.x, lane = threadIdx.x & 31;
    const bool live = e < e_end;
    const int key = live ? keys[e] : SORT_BINS + lane;                  // dead lanes: unique keys, no atomics
    int base = e_begin;
    for (int k = SORT_BINS - 1; k > key; k--) base += bins[k];          // heaviest first
    const unsigned peers = __match_any_sync(FULL, key);
    const int leader = __ffs(peers) - 1;
    int start = 0;
    if (live && lane == leader) start = atomicAdd(&bins[32 + key], __popc(peers));
    start = __shfl_sync(FULL, start, leader);
    if (live) S.perm[base + start + __popc(peers & ((1u << lane) - 1u))] = e;
}

// stand-alone forward-dynamics substep: q [N,28], qd [N,27], tau [N,21] f32 -> q, qd in place (524 B / env-substep)
__global__ void __launch_bounds__(ABA_BLOCK, SS_ABA_MINB) aba_substep_kernel(const DevModel* gm, float* __restrict__ q, float* __restrict__ qd,
                                                               const float* __restrict__ tau, int n) {
    BlockSmemDyn& B = *reinterpret_cast<BlockSmemDyn*>(smem_raw);
    stage_model(B, gm);
    const DevModel& M = B.M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * ABA_WPB + warp;
    if (e >= n) return;
    WarpSmemDyn& W = B.W[warp];
    float* qe = q + (size_t)e * SS_NQ; float* ve = qd + (size_t)e * ND;
    float a = lane < SS_NQ ? qe[lane] : 0.f;
    float quat[4], pos[3];
#pragma unroll
    for (int i = 0; i < 3; i++) pos[i] = shfl(a, i);
#pragma unroll
    for (int i = 0; i < 4; i++) quat[i] = shfl(a, 3 + i);
    float q_dof = shfl(a, (lane + 1) & 31);          // dof 6+j reads q[7+j]
    if (lane < 6 || lane >= ND) q_dof = 0.f;
    float v_dof = lane < ND ? ve[lane] : 0.f;
    float t = (lane >= 6 && lane < ND) ? tau[(size_t)e * NJ + lane - 6] : 0.f;
    float acc = dynamics_warp<false>(M, W, quat, q_dof, v_dof, t, lane, nullptr);
    float vn = v_dof + M.dt * acc;
    vn = fminf(fmaxf(vn, -M.max_coord_vel), M.max_coord_vel);
    if (lane >= ND) vn = 0.f;
    float dpos[3] = { 0.f, 0.f, 0.f };
    integrate_warp(M, quat, q_dof, vn, dpos, lane);
    float qh = shfl(q_dof, (lane + 31) & 31);
    float out = 0.f;
    if (lane < 3) out = (lane == 0 ? pos[0] + dpos[0] : lane == 1 ? pos[1] + dpos[1] : pos[2] + dpos[2]);
    else if (lane < 7) out = lane == 3 ? quat[0] : lane == 4 ? quat[1] : lane == 5 ? quat[2] : quat[3];
    else out = qh;
    if (lane < SS_NQ) qe[lane] = out;
    if (lane < ND) ve[lane] = vn;
}

// RolloutStorage.compute_returns(use_gae=True) -- algorithms/storage.py:59-71.  One thread per env, time scan.
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const float* __restrict__ masks,
                           const float* __restrict__ bad_masks, float* __restrict__ returns, int T, int N, float gamma, float lam) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float gae = 0.f, vnext = values[(size_t)T * N + i];
    for (int t = T - 1; t >= 0; t--) {
        float m = masks[(size_t)(t + 1) * N + i], bm = bad_masks[(size_t)(t + 1) * N + i], v = values[(size_t)t * N + i];
        float delta = rewards[(size_t)t * N + i] + gamma * vnext * m - v;
        gae = delta + gamma * lam * m * gae;
        gae = gae * bm;
        returns[(size_t)t * N + i] = gae + v;
        vnext = v;
    }
}

}  // namespace ssb

// ===============================================================================================
// host side: C-ABI
// ===============================================================================================
using namespace ssb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(SS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)


struct ss_env {
    int n = 0, device = 0, n_steps = 0, rendered = 0;
    DevModel hm;
    DevModel* dm = nullptr;
    DevState S{};
    double* cdf = nullptr;
    int table_head = 0; int64_t table_count = 1;     // ring of SS_TABLE_RING sampling-table versions (slot 0: the uniform table)
    int64_t launches = 0;
    size_t smem = 0;
    bool trace = false, debug = false;
    // host-buffer entry points: pinned staging the kernel reads actions from (via one H2D) and writes results into directly
    float *h_act = nullptr, *d_act = nullptr;        // pinned / device [N,21]
    unsigned char* h_out = nullptr;                  // pinned staging for outputs the caller did not give pinned memory for
    cudaStream_t own = nullptr;
    double host_t[4] = {};           // last ss_step_host: action copy, enqueue, waiting on the device, unpacking (seconds)
    float* ts_base = nullptr; double* ts_ctx = nullptr;      // create_temp_states scratch
    unsigned char* sort_keys = nullptr; int* sort_bins = nullptr;   // sort_envs scratch: [N] keys, 64 ints
};

static void fill_model(DevModel& D, const ss_model& m, const ss_params& p) {
    memset(&D, 0, sizeof D);
    for (int i = 0; i < NB; i++) {
        for (int k = 0; k < 3; k++) { D.anchor[i][k] = (float)m.anchor[i][k]; D.axis[i][k] = (float)m.axis[i][k]; D.com[i][k] = (float)m.com[i][k]; }
        for (int k = 0; k < 9; k++) D.R0[i][k] = (float)m.R0[i][k];
        for (int k = 0; k < 6; k++) D.inertia[i][k] = (float)m.inertia[i][k];
        D.mass[i] = (float)m.mass[i];
        int mask = 0;
        for (int b = i; b >= 1; b = m.parent[b]) mask |= 1 << b;
        D.anc_mask[i] = mask;
    }
    for (int j = 0; j < NJ; j++) {
        D.q_lo[j] = (float)m.q_lo[j]; D.q_hi[j] = (float)m.q_hi[j]; D.tau_max[j] = (float)m.tau_max[j];
        D.damping[j] = (float)m.joint_damping[j]; D.pd_p[j] = (float)m.pd_p[j]; D.pd_d[j] = (float)m.pd_d[j];
        D.init_q[j] = (float)m.init_q[j];
    }
    D.n_geom = m.n_geom; D.n_pair = m.n_pair; D.collision_mode = p.collision_mode; D.self_cross_terms = p.self_cross_terms;
    D.n_scand = 0;
    for (int g = 0; g < m.n_geom; g++) {
        double len2 = 0;
        for (int k = 0; k < 3; k++) {
            D.geom_p0[g][k] = (float)m.geom_p0[g][k]; D.geom_p1[g][k] = (float)m.geom_p1[g][k];
            len2 += (m.geom_p1[g][k] - m.geom_p0[g][k]) * (m.geom_p1[g][k] - m.geom_p0[g][k]);
        }
        D.geom_radius[g] = (float)m.geom_radius[g]; D.geom_half[g] = (float)(0.5 * sqrt(len2) * 1.0001 + 1e-6);
        D.geom_breaking[g] = (float)m.geom_breaking[g]; D.geom_body[g] = (unsigned char)m.geom_body[g];
    }
    // stone candidate table, pass by pass: feet (always), then -- collision_mode >= 1 -- the other capsules and the spheres
    for (int g = 0; g < m.n_geom; g++) {
        if (!m.geom_stone[g] || (p.collision_mode == 0 && m.geom_foot[g] < 0)) continue;
        for (int slot = 0; slot < 4; slot++)
            for (int end = 0; end < (m.geom_sphere[g] ? 1 : 2); end++)
                D.scand[D.n_scand++] = (unsigned short)(g | (slot << 5) | (end << 7));
    }
    for (int k = 0; k < m.n_pair; k++) { D.pair_a[k] = (unsigned char)m.pair_a[k]; D.pair_b[k] = (unsigned char)m.pair_b[k]; }
    D.stone_breaking = (float)p.stone_breaking; D.self_friction = (float)p.self_friction; D.self_erp = (float)p.self_erp;
    D.stone_radius = (float)(sqrt(p.stone_half[0] * p.stone_half[0] + p.stone_half[1] * p.stone_half[1] + p.stone_half[2] * p.stone_half[2]) * 1.0001);
    for (int f = 0; f < 2; f++) { D.foot_body[f] = m.foot_body[f]; for (int k = 0; k < 3; k++) D.foot_ref[f][k] = (float)m.foot_ref[f][k]; }
    D.contact_breaking = (float)m.contact_breaking; D.gravity = (float)p.gravity; D.dt = (float)p.dt; D.erp_limit = (float)p.erp_limit;
    double kd = p.dt * p.contact_stiffness + p.contact_damping;
    D.cfm_contact = (float)(1.0 / kd / p.dt); D.erp_contact = (float)(p.dt * p.contact_stiffness / kd);
    D.friction = (float)p.friction; D.linear_slop = (float)p.linear_slop; D.warmstart = (float)p.warmstart;
    D.max_coord_vel = (float)p.max_coord_vel; D.limit_max_impulse = (float)p.limit_max_impulse;
    for (int k = 0; k < 3; k++) D.stone_half[k] = (float)p.stone_half[k];
    D.stone_box_off = (float)(p.stone_com_z + p.stone_box_z); D.stone_com_z = (float)p.stone_com_z;
    D.n_substeps = p.n_substeps; D.n_iters = p.n_iters; D.robot_kind = p.robot_kind; D.terrain_mode = p.terrain_mode;
    D.n_steps = p.n_steps; D.rendered = p.rendered; D.stop_frames = p.stop_frames; D.max_episode_steps = p.max_episode_steps;
    D.n_phantoms = p.n_phantoms; D.use_sampler = p.use_sampler; D.actuator_mode = p.actuator_mode;
    D.done_height = (float)p.done_height; D.electricity_cost = (float)p.electricity_cost;
    D.stall_torque_cost = (float)p.stall_torque_cost; D.joints_at_limit_cost = (float)p.joints_at_limit_cost;
    for (int k = 0; k < 3; k++) D.init_pos[k] = m.init_pos[k];
    D.step_radius = p.step_radius; D.r_max = p.r_max; D.tilt_limit_deg = p.tilt_limit_deg; D.mirror_prob = p.mirror_prob;
    D.control_dt = p.control_dt;
    for (int k = 0; k < SS_GRID; k++) {
        D.samples[0][k] = p.yaw_samples[k]; D.samples[1][k] = p.pitch_samples[k]; D.samples[2][k] = p.r_samples[k];
        D.samples[3][k] = p.xtilt_samples[k]; D.samples[4][k] = p.ytilt_samples[k];
    }
}

static int grid_for(int n) { return (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; }

static int configure_kernels(size_t smem) {
    CK(cudaFuncSetAttribute(env_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(env_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(env_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(temp_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(aba_substep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlockSmemDyn)));
    return 0;
}

static int create_device_state(ss_env* E, const ss_model* model, const uint32_t* mt_env, const uint32_t* mt_aux);

extern "C" {

const char* ss_last_error(void) { return g_err.c_str(); }

int ss_create(const ss_model* model, const ss_params* params, int32_t n, int32_t device, const uint32_t* mt_env,
              const uint32_t* mt_aux, ss_env** out) {
    if (!model || !params || n <= 0 || !mt_env || !out) return fail(SS_ERR_ARG, "ss_create: bad argument");
    static const int parent_tab[NB] = SS_PARENT_TABLE;
    for (int i = 0; i < NB; i++)
        if (model->parent[i] != parent_tab[i]) return fail(SS_ERR_ARG, "ss_create: model topology differs from the compiled-in humanoid tree");
    if (params->n_steps < 4 || params->n_steps > SS_MAX_STONES || params->rendered < 2 || params->rendered > params->n_steps)
        return fail(SS_ERR_ARG, "ss_create: n_steps / rendered out of range");
    if (params->n_substeps < 1 || params->n_substeps > 4) return fail(SS_ERR_ARG, "ss_create: n_substeps must be 1..4");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(SS_ERR_CUDA, "ss_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(SS_ERR_ARG, "ss_create: device index out of range");
    CK(cudaSetDevice(device));
    ss_env* E = new ss_env();
    E->n = n; E->device = device; E->n_steps = params->n_steps; E->rendered = params->rendered;
    fill_model(E->hm, *model, *params);
    E->smem = sizeof(BlockSmem);
    if (const char* pad = getenv("SS_SMEM_PAD")) {          // experiments only: force one CTA per SM for small-CTA builds (occupancy scans)
        const size_t want = (size_t)atol(pad);
        if (want > E->smem && want <= 232448) E->smem = want;
    }
    int rc = configure_kernels(E->smem);
    if (!rc) rc = create_device_state(E, model, mt_env, mt_aux);
    if (rc) {                                    // e.g. out of memory half-way: release what was allocated, keep the message
        const std::string msg = g_err;
        ss_destroy(E);
        cudaGetLastError();
        g_err = msg;
        return rc;
    }
    *out = E;
    return SS_OK;
}

} // extern "C"

static int create_device_state(ss_env* E, const ss_model* model, const uint32_t* mt_env, const uint32_t* mt_aux) {
    const int n = E->n;
    DevState& S = E->S;
    S.n = n; S.n_cells = SS_GRID_CELLS;
    size_t N = (size_t)n;
    CK(cudaMalloc(&E->dm, sizeof(DevModel)));
    CK(cudaMemcpy(E->dm, &E->hm, sizeof(DevModel), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&S.pos, N * 3 * sizeof(double)));
    CK(cudaMalloc(&S.sv, N * 64 * sizeof(float)));
    CK(cudaMalloc(&S.warm, N * SS_NWARM * sizeof(float)));
    CK(cudaMalloc(&S.terrain, N * E->n_steps * 6 * sizeof(double)));
    CK(cudaMalloc(&S.plank, N * E->rendered * 6 * sizeof(double)));
    CK(cudaMalloc(&S.scal, N * SS_NSCAL * sizeof(double)));
    CK(cudaMalloc(&S.ints, N * SS_NINT * sizeof(int)));
    CK(cudaMalloc(&S.mt, N * 2 * SS_MT_WORDS * sizeof(uint32_t)));
    CK(cudaMalloc(&S.perm, N * sizeof(int)));
    CK(cudaMalloc(&S.err, sizeof(int))); CK(cudaMemset(S.err, 0, sizeof(int)));
    CK(cudaMalloc(&E->sort_keys, N)); CK(cudaMalloc(&E->sort_bins, SORT_SLOT * sizeof(int)));
    { std::vector<int> id(N); for (size_t i = 0; i < N; i++) id[i] = (int)i; CK(cudaMemcpy(S.perm, id.data(), N * sizeof(int), cudaMemcpyHostToDevice)); }
    CK(cudaMalloc(&E->cdf, (size_t)SS_TABLE_RING * SS_GRID_CELLS * sizeof(double)));
    S.cdf = E->cdf;
    CK(cudaMalloc(&S.tslot, N * sizeof(int))); CK(cudaMemset(S.tslot, 0, N * sizeof(int)));
    CK(cudaMemset(S.sv, 0, N * 64 * sizeof(float)));
    {
        std::vector<float> w(N * SS_NWARM, 0.f);
        for (size_t i = 0; i < N; i++) for (int k = 0; k < SS_MAX_CONTACTS; k++) w[i * SS_NWARM + k] = -1.f;
        CK(cudaMemcpy(S.warm, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(S.terrain, 0, N * E->n_steps * 6 * sizeof(double)));
    CK(cudaMemset(S.plank, 0, N * E->rendered * 6 * sizeof(double)));
    CK(cudaMemset(S.scal, 0, N * SS_NSCAL * sizeof(double)));
    CK(cudaMemset(S.ints, 0, N * SS_NINT * sizeof(int)));
    {
        std::vector<uint32_t> mt(N * 2 * SS_MT_WORDS, 0u);
        for (size_t i = 0; i < N; i++) {
            memcpy(&mt[i * 2 * SS_MT_WORDS], mt_env + i * SS_MT_WORDS, SS_MT_WORDS * 4);
            if (mt_aux) memcpy(&mt[i * 2 * SS_MT_WORDS + SS_MT_WORDS], mt_aux + i * SS_MT_WORDS, SS_MT_WORDS * 4);
            else mt[i * 2 * SS_MT_WORDS + SS_MT_WORDS + 624] = 624;
        }
        CK(cudaMemcpy(S.mt, mt.data(), mt.size() * 4, cudaMemcpyHostToDevice));
        std::vector<double> cdf(SS_GRID_CELLS);
        for (int i = 0; i < SS_GRID_CELLS; i++) cdf[i] = (double)(i + 1) / SS_GRID_CELLS;
        CK(cudaMemcpy(E->cdf, cdf.data(), cdf.size() * 8, cudaMemcpyHostToDevice));
        // initial pose (matches the oracle's sso_create): base at init_pos, identity quat, running-start hinges
        std::vector<float> sv(N * 64, 0.f); std::vector<double> pos(N * 3);
        for (size_t i = 0; i < N; i++) {
            sv[i * 64 + 3] = 1.f;
            for (int j = 0; j < NJ; j++) sv[i * 64 + 4 + j] = (float)model->init_q[j];
            for (int k = 0; k < 3; k++) pos[i * 3 + k] = model->init_pos[k];
        }
        CK(cudaMemcpy(S.sv, sv.data(), sv.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.pos, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice));
    }
    CK(cudaStreamCreateWithFlags(&E->own, cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());                     // the initial-state uploads above are pageable copies (see ss_set_state)
    return SS_OK;
}

extern "C" {

void ss_destroy(ss_env* E) {
    if (!E) return;
    cudaSetDevice(E->device);
    cudaDeviceSynchronize();
    DevState& S = E->S;
    cudaFree(E->dm); cudaFree(S.pos); cudaFree(S.sv); cudaFree(S.warm); cudaFree(S.terrain); cudaFree(S.plank);
    cudaFree(S.scal); cudaFree(S.ints); cudaFree(S.mt); cudaFree(S.perm); cudaFree(S.err); cudaFree(S.tslot); cudaFree(E->cdf); cudaFree(S.trace_q); cudaFree(S.trace_qd); cudaFree(S.trace_ev); cudaFree(S.dbg);
    cudaFreeHost(E->h_act); cudaFreeHost(E->h_out); cudaFree(E->d_act);
    cudaFree(E->ts_base); cudaFree(E->ts_ctx); cudaFree(E->sort_keys); cudaFree(E->sort_bins);
    if (E->own) cudaStreamDestroy(E->own);
    delete E;
}

int ss_reset_all(ss_env* E, float* obs_dev, void* stream) {
    if (!E || !obs_dev) return fail(SS_ERR_ARG, "ss_reset_all: bad argument");
    CK(cudaSetDevice(E->device));
    env_reset_kernel<<<grid_for(E->n), BLOCK, E->smem, (cudaStream_t)stream>>>(E->dm, E->S, obs_dev);
    E->launches++;
    CK(cudaGetLastError());
    return SS_OK;
}

static int launch_step(ss_env* E, const float* act, float* obs, float* rew, uint8_t* done, uint8_t* to, float* er, int32_t* el,
                       cudaStream_t st, int physics, int e_begin = 0, int e_end = -1, double* rew64 = nullptr, int slot = 0) {
    if (e_end < 0) e_end = E->n;
    if (physics == 1 && e_end - e_begin >= 4 * WARPS_PER_BLOCK) {
        int* bins = E->sort_bins; (void)slot;
        const int nb = (e_end - e_begin + SORT_THREADS - 1) / SORT_THREADS;
        CK(cudaMemsetAsync(bins, 0, SORT_SLOT * sizeof(int), st));
        sort_hist_kernel<<<nb, SORT_THREADS, 0, st>>>(E->S, E->sort_keys, bins, e_begin, e_end);
        sort_scatter_kernel<<<nb, SORT_THREADS, 0, st>>>(E->S, E->sort_keys, bins, e_begin, e_end);
        E->launches += 2;
    }
    if (E->S.dbg || E->S.trace_q || E->S.trace_ev)
        env_step_kernel<true><<<grid_for(e_end - e_begin), BLOCK, E->smem, st>>>(E->dm, E->S, act, obs, rew, done, to, er, el, physics,
                                                                                 e_begin, e_end, rew64);
    else
        env_step_kernel<false><<<grid_for(e_end - e_begin), BLOCK, E->smem, st>>>(E->dm, E->S, act, obs, rew, done, to, er, el, physics,
                                                                                  e_begin, e_end, rew64);
    E->launches++;
    CK(cudaGetLastError());
    return SS_OK;
}

int ss_step(ss_env* E, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, uint8_t* timeout_dev,
            float* ep_ret_dev, int32_t* ep_len_dev, void* stream) {
    if (!E || !act_dev || !obs_dev || !rew_dev || !done_dev || !timeout_dev || !ep_ret_dev || !ep_len_dev)
        return fail(SS_ERR_ARG, "ss_step: null buffer");
    CK(cudaSetDevice(E->device));
    return launch_step(E, act_dev, obs_dev, rew_dev, done_dev, timeout_dev, ep_ret_dev, ep_len_dev, (cudaStream_t)stream, 1);
}

int ss_step_injected(ss_env* E, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, uint8_t* timeout_dev,
                     float* ep_ret_dev, int32_t* ep_len_dev, void* stream) {
    if (!E || !act_dev || !obs_dev || !rew_dev || !done_dev || !timeout_dev || !ep_ret_dev || !ep_len_dev)
        return fail(SS_ERR_ARG, "ss_step_injected: null buffer");
    CK(cudaSetDevice(E->device));
    return launch_step(E, act_dev, obs_dev, rew_dev, done_dev, timeout_dev, ep_ret_dev, ep_len_dev, (cudaStream_t)stream, 2);
}

int ss_env_post(ss_env* E, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, void* stream) {
    if (!E || !act_dev || !obs_dev || !rew_dev || !done_dev) return fail(SS_ERR_ARG, "ss_env_post: null buffer");
    CK(cudaSetDevice(E->device));
    return launch_step(E, act_dev, obs_dev, rew_dev, done_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream, 0);
}

static int ensure_staging(ss_env* E) {
    if (E->h_act) return SS_OK;
    size_t N = (size_t)E->n;
    CK(cudaMallocHost(&E->h_act, N * NJ * 4));
    CK(cudaMallocHost(&E->h_out, N * (SS_OBS * 4 + 8 + 1 + 1 + 4 + 4) + 64));
    CK(cudaMalloc(&E->d_act, N * NJ * 4));
    return SS_OK;
}

/* Where the kernel writes one output array of a host-buffer call: straight into the caller's array when that is pinned host
 * memory (ss_host_alloc, cudaHostAlloc, cudaHostRegister: device-accessible under unified addressing), else into the
 * library's own pinned staging (`stage`), from which it is copied afterwards (*direct = false). */
static void* out_target(void* user, void* stage, bool* direct) {
    *direct = false;
#ifndef SS_SIMT_EMU
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, user) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        *direct = true;
        return at.devicePointer;
    }
    cudaGetLastError();
#endif
    return stage;
}

/* Host buffers in, host buffers out.  actions: checked while copied into pinned memory, one H2D DMA; ONE launch over the whole
 * batch (the sliced pipeline of round 1 paid four partial waves and 24 enqueues per step); the kernel writes every output --
 * observation rows, float64 rewards, done / timeout bytes, episode statistics -- straight into pinned host memory, the
 * caller's own arrays where those are pinned: nothing is copied back or unpacked. */
int ss_step_host(ss_env* E, const float* act, float* obs, double* rew, uint8_t* done, uint8_t* to, float* er, int32_t* el) {
    if (!E || !act || !obs || !rew || !done || !to || !er || !el) return fail(SS_ERR_ARG, "ss_step_host: null buffer");
    CK(cudaSetDevice(E->device));
    int rc = ensure_staging(E); if (rc) return rc;
    const size_t N = (size_t)E->n;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    // robots.py:290 `assert np.isfinite(a).all()`: checked on the pinned copy BEFORE anything is enqueued (a refused batch
    // leaves the envs untouched).  (u & exponent mask) + 2^23 carries into bit 31 exactly when the exponent field is all
    // ones (inf / nan): one and + add + or per float, vectorised by the host compiler.
    const size_t cnt = N * NJ;
    memcpy(E->h_act, act, cnt * 4);
    {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(E->h_act);
        uint32_t bad = 0;
        for (size_t i = 0; i < cnt; i++) bad |= (p[i] & 0x7f800000u) + 0x00800000u;
        if (bad & 0x80000000u) return fail(SS_ERR_ARG, "ss_step_host: non-finite action");
    }
    const double t1 = now();
    // staging layout: rew f64 [N] | obs f32 [N,60] | er f32 [N] | el i32 [N] | done u8 [N] | to u8 [N]
    unsigned char* st = E->h_out;
    bool d_obs, d_rew, d_done, d_to, d_er, d_el;
    double* t_rew = (double*)out_target(rew, st, &d_rew);             st += N * 8;
    float* t_obs = (float*)out_target(obs, st, &d_obs);               st += N * SS_OBS * 4;
    float* t_er = (float*)out_target(er, st, &d_er);                  st += N * 4;
    int32_t* t_el = (int32_t*)out_target(el, st, &d_el);              st += N * 4;
    uint8_t* t_done = (uint8_t*)out_target(done, st, &d_done);        st += N;
    uint8_t* t_to = (uint8_t*)out_target(to, st, &d_to);
    CK(cudaMemcpyAsync(E->d_act, E->h_act, cnt * 4, cudaMemcpyHostToDevice, E->own));
    rc = launch_step(E, E->d_act, t_obs, nullptr, t_done, t_to, t_er, t_el, E->own, 1, 0, (int)N, t_rew); if (rc) return rc;
    const double t2 = now();
    CK(cudaStreamSynchronize(E->own));
    const double t3 = now();
    if (!d_rew) memcpy(rew, t_rew, N * 8);
    if (!d_obs) memcpy(obs, t_obs, N * SS_OBS * 4);
    if (!d_er) memcpy(er, t_er, N * 4);
    if (!d_el) memcpy(el, t_el, N * 4);
    if (!d_done) memcpy(done, t_done, N);
    if (!d_to) memcpy(to, t_to, N);
    E->host_t[0] = t1 - t0; E->host_t[1] = t2 - t1; E->host_t[2] = t3 - t2; E->host_t[3] = now() - t3;
    return SS_OK;
}

int ss_reset_host(ss_env* E, float* obs) {
    if (!E || !obs) return fail(SS_ERR_ARG, "ss_reset_host: null buffer");
    CK(cudaSetDevice(E->device));
    int rc = ensure_staging(E); if (rc) return rc;
    bool direct; float* ot = (float*)out_target(obs, E->h_out + (size_t)E->n * 8, &direct);
    rc = ss_reset_all(E, ot, E->own); if (rc) return rc;
    CK(cudaStreamSynchronize(E->own));
    if (!direct) memcpy(obs, ot, (size_t)E->n * SS_OBS * 4);
    return SS_OK;
}

/* Pinned host memory for the caller's observation arrays (the kernel then writes into them directly). */
int ss_host_alloc(size_t bytes, void** out) {
    if (!out || !bytes) return fail(SS_ERR_ARG, "ss_host_alloc: bad argument");
    CK(cudaMallocHost(out, bytes));
    return SS_OK;
}
void ss_host_free(void* p) { if (p) cudaFreeHost(p); }

int ss_set_sample_cdf(ss_env* E, const double* cdf, int32_t n_cells, int32_t apply_now, void* stream) {
    if (!E || !cdf || n_cells != SS_GRID_CELLS) return fail(SS_ERR_ARG, "ss_set_sample_cdf: table must have 11^5 cells");
    CK(cudaSetDevice(E->device));
    // stream == NULL (host-buffer mode): the library's own stream, synchronised before returning.  The legacy default
    // stream would NOT do: ss_step_host works on cudaStreamNonBlocking streams, which do not wait for it, so the re-sample
    // (it writes terrain, planks and RNG state) could still be running when the next step starts.
    // (a NULL stream is also what torch hands over while its current stream is the legacy default stream)
    cudaStream_t st = stream ? (cudaStream_t)stream : E->own;
    if (!stream) CK(cudaDeviceSynchronize());       // whatever stream the caller's earlier work used (probes run on torch's)
    CK(cudaStreamSynchronize(st));
    // The table comes from pageable host memory.  A plain cudaMemcpy returns once the data is STAGED; the DMA itself is
    // ordered in the legacy default stream only, which cudaStreamNonBlocking streams do not wait for -- the re-sample
    // below (or the next ss_step_host) could read a half-written table.  Copy on the stream that uses it and wait.
    // a new table VERSION (include/ss_env.h): envs still on the ring slot about to be overwritten move to the oldest survivor
    const int slot = (E->table_head + 1) % SS_TABLE_RING, nb = (E->n + 255) / 256;
    if (E->table_count >= SS_TABLE_RING) table_assign_kernel<<<nb, 256, 0, st>>>(E->S.tslot, E->n, slot, (slot + 1) % SS_TABLE_RING);
    CK(cudaMemcpyAsync(E->cdf + (size_t)slot * n_cells, cdf, (size_t)n_cells * 8, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    E->table_head = slot; E->table_count++;
    if (apply_now) {                          // update_sample_prob: flagged envs take the version and re-sample at once
        resample_kernel<<<grid_for(E->n), BLOCK, E->smem, st>>>(E->dm, E->S, slot);
        E->launches++;
    } else {                                  // update_curriculum / update_specialist: every env takes it
        table_assign_kernel<<<nb, 256, 0, st>>>(E->S.tslot, E->n, -1, slot);
    }
    CK(cudaGetLastError());
    if (!stream) CK(cudaStreamSynchronize(st));
    return SS_OK;
}

int ss_table_ring_state(const ss_env* E, int32_t* head, int64_t* count) {
    if (!E || !head || !count) return fail(SS_ERR_ARG, "ss_table_ring_state: bad argument");
    *head = E->table_head; *count = E->table_count;
    return SS_OK;
}

int ss_update_tables(ss_env* E, const ss_model* model, const ss_params* params, void* stream) {
    if (!E || !model || !params) return fail(SS_ERR_ARG, "ss_update_tables: bad argument");
    if (params->n_steps != E->n_steps || params->rendered != E->rendered)
        return fail(SS_ERR_ARG, "ss_update_tables: n_steps / rendered are fixed at ss_create (they size the terrain arrays)");
    if (params->n_substeps < 1 || params->n_substeps > 4) return fail(SS_ERR_ARG, "ss_update_tables: n_substeps must be 1..4");
    static const int parent_tab[NB] = SS_PARENT_TABLE;
    for (int i = 0; i < NB; i++)
        if (model->parent[i] != parent_tab[i]) return fail(SS_ERR_ARG, "ss_update_tables: model topology differs from the compiled-in humanoid tree");
    CK(cudaSetDevice(E->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    CK(cudaStreamSynchronize(E->own));
    fill_model(E->hm, *model, *params);
    CK(cudaMemcpy(E->dm, &E->hm, sizeof(DevModel), cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());                     // (pageable copy: see ss_set_state)
    return SS_OK;
}

int64_t ss_temp_state_cells(int32_t variant) {
    return variant == 1 ? 121 : (variant == 2 ? 1331 : (variant == 5 ? SS_GRID_CELLS : 0));
}

int ss_create_temp_states(ss_env* E, int32_t variant, const double* grid, const double* restore, float* out_dev, void* stream) {
    if (!E || !grid || !out_dev || ss_temp_state_cells(variant) == 0) return fail(SS_ERR_ARG, "ss_create_temp_states: bad argument");
    CK(cudaSetDevice(E->device));
    if (!E->ts_base) { CK(cudaMalloc(&E->ts_base, (size_t)E->n * 64 * 4)); CK(cudaMalloc(&E->ts_ctx, (size_t)E->n * TS_CTX * 8)); }
    TempGrid G; memset(&G, 0, sizeof G);
    memcpy(G.a, grid, sizeof G.a);
    if (restore) { memcpy(G.restore, restore, sizeof G.restore); G.has_restore = 1; }
    G.variant = variant; G.n3 = variant >= 2 ? SS_GRID : 1; G.n45 = variant == 5 ? SS_GRID : 1;
    const int cells = (int)ss_temp_state_cells(variant);
    cudaStream_t st = (cudaStream_t)stream;
    temp_prep_kernel<<<grid_for(E->n), BLOCK, E->smem, st>>>(E->dm, E->S, G, E->ts_base, E->ts_ctx);
    CK(cudaGetLastError());
    dim3 grid2((cells + TS_CELLS - 1) / TS_CELLS, E->n);
    temp_fill_kernel<<<grid2, TS_CELLS, 0, st>>>(E->dm, G, E->ts_base, E->ts_ctx, out_dev, cells);
    CK(cudaGetLastError());
    E->launches += 2;
    return SS_OK;
}

int ss_get_state(ss_env* E, ss_state_view* v) {
    if (!E || !v) return fail(SS_ERR_ARG, "ss_get_state: bad argument");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    size_t N = (size_t)E->n; DevState& S = E->S;
    if (v->q || v->qd) {
        std::vector<float> sv(N * 64); std::vector<double> pos(N * 3);
        CK(cudaMemcpy(sv.data(), S.sv, sv.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pos.data(), S.pos, pos.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < N; i++) {
            if (v->q) {
                double* q = v->q + i * SS_NQ;
                for (int k = 0; k < 3; k++) q[k] = pos[i * 3 + k];
                for (int k = 0; k < 4 + NJ; k++) q[3 + k] = sv[i * 64 + k];
            }
            if (v->qd) for (int k = 0; k < ND; k++) v->qd[i * ND + k] = sv[i * 64 + 32 + k];
        }
    }
    if (v->warm) {
        std::vector<float> w(N * SS_NWARM);
        CK(cudaMemcpy(w.data(), S.warm, w.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < w.size(); i++) v->warm[i] = w[i];
    }
    if (v->terrain) CK(cudaMemcpy(v->terrain, S.terrain, N * E->n_steps * 48, cudaMemcpyDeviceToHost));
    if (v->plank) CK(cudaMemcpy(v->plank, S.plank, N * E->rendered * 48, cudaMemcpyDeviceToHost));
    if (v->scalars) CK(cudaMemcpy(v->scalars, S.scal, N * SS_NSCAL * 8, cudaMemcpyDeviceToHost));
    if (v->ints) CK(cudaMemcpy(v->ints, S.ints, N * SS_NINT * 4, cudaMemcpyDeviceToHost));
    if (v->mt) CK(cudaMemcpy(v->mt, S.mt, N * 2 * SS_MT_WORDS * 4, cudaMemcpyDeviceToHost));
    if (v->table_slot) CK(cudaMemcpy(v->table_slot, S.tslot, N * 4, cudaMemcpyDeviceToHost));
    return SS_OK;
}

int ss_set_state(ss_env* E, const ss_state_view* v) {
    if (!E || !v) return fail(SS_ERR_ARG, "ss_set_state: bad argument");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    size_t N = (size_t)E->n; DevState& S = E->S;
    if (v->q || v->qd) {
        std::vector<float> sv(N * 64); std::vector<double> pos(N * 3);
        CK(cudaMemcpy(sv.data(), S.sv, sv.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pos.data(), S.pos, pos.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < N; i++) {
            if (v->q) {
                const double* q = v->q + i * SS_NQ;
                for (int k = 0; k < 3; k++) pos[i * 3 + k] = q[k];
                for (int k = 0; k < 4 + NJ; k++) sv[i * 64 + k] = (float)q[3 + k];
            }
            if (v->qd) for (int k = 0; k < ND; k++) sv[i * 64 + 32 + k] = (float)v->qd[i * ND + k];
        }
        CK(cudaMemcpy(S.sv, sv.data(), sv.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.pos, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice));
    }
    if (v->warm) {
        std::vector<float> w(N * SS_NWARM);
        for (size_t i = 0; i < w.size(); i++) w[i] = (float)v->warm[i];
        CK(cudaMemcpy(S.warm, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    }
    if (v->terrain) CK(cudaMemcpy(S.terrain, v->terrain, N * E->n_steps * 48, cudaMemcpyHostToDevice));
    if (v->plank) CK(cudaMemcpy(S.plank, v->plank, N * E->rendered * 48, cudaMemcpyHostToDevice));
    if (v->scalars) CK(cudaMemcpy(S.scal, v->scalars, N * SS_NSCAL * 8, cudaMemcpyHostToDevice));
    if (v->ints) CK(cudaMemcpy(S.ints, v->ints, N * SS_NINT * 4, cudaMemcpyHostToDevice));
    if (v->mt) CK(cudaMemcpy(S.mt, v->mt, N * 2 * SS_MT_WORDS * 4, cudaMemcpyHostToDevice));
    if (v->table_slot) {
        for (size_t i = 0; i < N; i++) if (v->table_slot[i] < 0 || v->table_slot[i] >= SS_TABLE_RING) return fail(SS_ERR_ARG, "ss_set_state: table_slot out of range");
        CK(cudaMemcpy(S.tslot, v->table_slot, N * 4, cudaMemcpyHostToDevice));
    }
    // copies from pageable memory return once staged; their DMAs are ordered in the legacy default stream only, which the
    // library's cudaStreamNonBlocking streams (ss_step_host) do not wait for: make "synchronous" true for every stream
    CK(cudaDeviceSynchronize());
    return SS_OK;
}

int ss_enable_substep_trace(ss_env* E, int32_t on) {
    if (!E) return fail(SS_ERR_ARG, "ss_enable_substep_trace: bad argument");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    DevState& S = E->S;
    if (on && !S.trace_q) {
        CK(cudaMalloc(&S.trace_q, (size_t)4 * E->n * SS_NQ * 8)); CK(cudaMalloc(&S.trace_qd, (size_t)4 * E->n * ND * 8));
        CK(cudaMemset(S.trace_q, 0, (size_t)4 * E->n * SS_NQ * 8)); CK(cudaMemset(S.trace_qd, 0, (size_t)4 * E->n * ND * 8));
        CK(cudaMalloc(&S.trace_ev, (size_t)4 * E->n * SS_NEV * 4)); CK(cudaMemset(S.trace_ev, 0xff, (size_t)4 * E->n * SS_NEV * 4));
    } else if (!on && S.trace_q) {
        cudaFree(S.trace_q); cudaFree(S.trace_qd); cudaFree(S.trace_ev); S.trace_q = S.trace_qd = nullptr; S.trace_ev = nullptr;
    }
    CK(cudaDeviceSynchronize());                     // the memsets above run in the legacy stream (see ss_set_state)
    return SS_OK;
}

int ss_substep_trace(ss_env* E, double* q, double* qd) {
    if (!E || !q || !qd) return fail(SS_ERR_ARG, "ss_substep_trace: bad argument");
    if (!E->S.trace_q) return fail(SS_ERR_STATE, "ss_substep_trace: tracing is not enabled");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(q, E->S.trace_q, (size_t)4 * E->n * SS_NQ * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(qd, E->S.trace_qd, (size_t)4 * E->n * ND * 8, cudaMemcpyDeviceToHost));
    return SS_OK;
}

int ss_substep_events(ss_env* E, int32_t* ev) {
    if (!E || !ev) return fail(SS_ERR_ARG, "ss_substep_events: bad argument");
    if (!E->S.trace_ev) return fail(SS_ERR_STATE, "ss_substep_events: tracing is not enabled");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ev, E->S.trace_ev, (size_t)4 * E->n * SS_NEV * 4, cudaMemcpyDeviceToHost));
    return SS_OK;
}

/* diagnostics for the parity tests: per-env dump of the LAST substep's H, C, contact candidates, rows, Delassus
 * matrix (ssb::DBG_* layout).  out_host [N, ss_debug_floats()] */
int32_t ss_debug_floats(void) { return DBG_FLOATS; }
int ss_enable_debug(ss_env* E, int32_t on) {
    if (!E) return fail(SS_ERR_ARG, "ss_enable_debug: bad argument");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    if (on && !E->S.dbg) { CK(cudaMalloc(&E->S.dbg, (size_t)E->n * DBG_FLOATS * 4)); CK(cudaMemset(E->S.dbg, 0, (size_t)E->n * DBG_FLOATS * 4)); }
    else if (!on && E->S.dbg) { cudaFree(E->S.dbg); E->S.dbg = nullptr; }
    CK(cudaDeviceSynchronize());
    return SS_OK;
}
int ss_debug_read(ss_env* E, float* out_host) {
    if (!E || !out_host || !E->S.dbg) return fail(SS_ERR_STATE, "ss_debug_read: debug dump is not enabled");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_host, E->S.dbg, (size_t)E->n * DBG_FLOATS * 4, cudaMemcpyDeviceToHost));
    return SS_OK;
}

int ss_aba_substep(ss_env* E, float* q_dev, float* qd_dev, const float* tau_dev, int32_t n, void* stream) {
    if (!E || !q_dev || !qd_dev || !tau_dev || n <= 0) return fail(SS_ERR_ARG, "ss_aba_substep: bad argument");
    CK(cudaSetDevice(E->device));
    aba_substep_kernel<<<(n + ABA_WPB - 1) / ABA_WPB, ABA_BLOCK, sizeof(BlockSmemDyn), (cudaStream_t)stream>>>(E->dm, q_dev, qd_dev, tau_dev, n);
    E->launches++;
    CK(cudaGetLastError());
    return SS_OK;
}

int ss_gae(const float* rewards, const float* values, const float* masks, const float* bad_masks, float* returns, int32_t T,
           int32_t N, float gamma, float lam, void* stream) {
    if (!rewards || !values || !masks || !bad_masks || !returns || T <= 0 || N <= 0) return fail(SS_ERR_ARG, "ss_gae: bad argument");
    gae_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, masks, bad_masks, returns, T, N, gamma, lam);
    CK(cudaGetLastError());
    return SS_OK;
}

int ss_error_flags(ss_env* E, int32_t* flags, int32_t clear) {
    if (!E || !flags) return fail(SS_ERR_ARG, "ss_error_flags: bad argument");
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(flags, E->S.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (clear) { CK(cudaMemset(E->S.err, 0, sizeof(int))); CK(cudaDeviceSynchronize()); }
    return SS_OK;
}

/* seconds the last ss_step_host spent copying+checking actions, enqueueing, waiting on the device, unpacking */
int ss_host_timing(const ss_env* E, double* out4) {
    if (!E || !out4) return fail(SS_ERR_ARG, "ss_host_timing: bad argument");
    for (int i = 0; i < 4; i++) out4[i] = E->host_t[i];
    return SS_OK;
}

int32_t ss_num_envs(const ss_env* E) { return E ? E->n : 0; }
int64_t ss_launch_count(const ss_env* E) { return E ? E->launches : 0; }

}  // extern "C"
