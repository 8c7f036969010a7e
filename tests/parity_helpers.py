"""Shared parity machinery (test infrastructure): used by the GPU tests (tests/test_gpu_parity.py, through the C-ABI on a
B200) and by the CPU suite on the SIMT emulation of the same kernels (tests/test_simt_emu.py).

1. ``chain_check`` -- STRICT parity of chained substeps.  A physics substep takes discrete decisions: a joint-limit row
   exists iff ``q - lo <= 0`` (btMultiBodyJointLimitConstraint), a contact point enters the solver or not.  fp32 device
   Which branch of the contact geometry produced a point, and the solver iteration at which a contact's normal impulse
   drops to zero (Bullet then skips its friction rows, which keep their impulse), are decisions of the same kind.  fp32 device
   arithmetic and the fp64 oracle can decide differently where the state sits within rounding of such a threshold; from
   that substep on the two trajectories are different problems.  Both sides therefore record their decisions per substep
   (``ss_substep_events`` / ``sso_substep_events``) and the check is:
     * every env, at every substep up to which ALL decisions agree, is within the fp32 tolerance -- no quota;
     * every disagreement is accounted for: a limit-row flip must sit at a hinge within rounding of that limit in the
       oracle's own state, contact / solver-history flips are counted and bounded.
   Envs whose velocity clamp (``max_coord_vel``, Bullet's 100 rad/s safety net) engaged in a substep -- both sides record
   that too, bit 30 of the event word -- are outside the regime the tolerance is stated for and are excluded from that
   substep on by that explicit criterion (and counted).
2. ``replay_golden`` -- the reference-generated tapes (tests/golden/*.npz, tools/gen_golden.py) replayed through any
   backend with the oracle's probe interface (reset / step_injected / set_sample_cdf / create_temp_states / get_state).
"""
import numpy as np

# fp32 tolerance after 1..4 chained substeps from identical states.  Velocities relative to the env's velocity scale
# (max(1, |qd|_inf), SURVEY 8c: 1e-3 relative in contact); positions = h x velocity tolerance (5e-6 ~ 1e-3 / 240).
TOL_V = np.array([1e-3, 2.5e-3, 5e-3, 1e-2])
TOL_Q = np.array([5e-6, 2e-5, 8e-5, 2e-4])


def chain_check(o, trace_o, trace_g, ev_o, ev_g, q_start, label="", max_flip_frac=0.15, max_contact_flip_frac=0.08):
    """trace_* = (q [S,n,28], qd [S,n,27]); ev_* [S,n,18]; q_start [n,28] the common start state.
    Returns ``strict`` [n]: envs whose decisions agree over the whole chain and whose clamp never engaged (for those the
    caller may compare observations / rewards too)."""
    qo, qdo = trace_o; qg, qdg = trace_g
    S, n = qo.shape[0], qo.shape[1]
    lo, hi = np.array(o.model.q_lo), np.array(o.model.q_hi)
    vmax = float(o.params.max_coord_vel)
    CLAMP = 1 << 30                                                                   # bit 30 of the upper-limit word
    agree = (ev_o[:S] == ev_g[:S]).all(axis=2)
    same = np.logical_and.accumulate(agree, axis=0)                                   # decisions equal up to substep s
    vscale = np.maximum(1.0, np.abs(qdo).max(axis=2))
    clamped = np.logical_or.accumulate(((ev_o[:S, :, 1] | ev_g[:S, :, 1]) & CLAMP) != 0, axis=0)
    assert vmax > 0
    e_q = np.abs(qo - qg).max(axis=2); e_v = np.abs(qdo - qdg).max(axis=2)
    ok = (e_q < TOL_Q[:S, None] * vscale) & (e_v < TOL_V[:S, None] * vscale)
    strict = same & ~clamped
    bad = strict & ~ok
    assert not bad.any(), "%s: %d env-substeps with equal discrete decisions outside the fp32 tolerance, first (substep, env) %s: dq %.3g dv %.3g vscale %.3g" % (
        label, bad.sum(), tuple(np.argwhere(bad)[0]), e_q[bad][0], e_v[bad][0], vscale[bad][0])
    # ---- account for every flip ----
    flipped = ~same[S - 1]
    n_contact_flip = 0
    for i in np.flatnonzero(flipped):
        s0 = int(np.argmin(same[:, i]))                                               # first substep whose decisions differ
        if s0 > 0 and clamped[s0 - 1, i]:
            continue                                                                  # already outside the regime (clamp) before it flipped
        q_at = q_start[i] if s0 == 0 else qo[s0 - 1, i]                               # the oracle's state when it decided
        dl, dh = int(ev_o[s0, i, 0] ^ ev_g[s0, i, 0]), int((ev_o[s0, i, 1] ^ ev_g[s0, i, 1]) & ~CLAMP)
        if not (dl or dh):
            n_contact_flip += 1
            continue
        tol = max(TOL_Q[max(s0 - 1, 0)], 1e-6) * (vscale[max(s0 - 1, 0), i] if s0 else 1.0) * 2
        gaps = [abs(q_at[7 + j] - lo[j]) for j in range(21) if (dl >> j) & 1] + [abs(hi[j] - q_at[7 + j]) for j in range(21) if (dh >> j) & 1]
        rows = max(bin(int(ev[s0, i, 0])).count("1") + bin(int(ev[s0, i, 1]) & ~CLAMP).count("1") for ev in (ev_o, ev_g))
        # at most SS_MAX_LIMIT_ROWS (8) limit rows are kept, in hinge order: with the cap reached one genuine flip also
        # changes WHICH later hinges keep their row, so there only the closest flipped hinge must sit on its limit
        worst = min(gaps) if rows >= 8 else max(gaps)
        assert worst < tol, "%s: env %d substep %d: a limit row flipped %.3g away from its limit (masks %x/%x vs %x/%x)" % (
            label, i, s0, worst, ev_o[s0, i, 0], ev_o[s0, i, 1], ev_g[s0, i, 0], ev_g[s0, i, 1])
    print("%s: %d envs; decisions agree over the chain for %d, limit flips %d, contact flips %d, clamp engaged %d; worst in-tolerance ratios dq %.2f dv %.2f" % (
        label, n, int(same[S - 1].sum()), int(flipped.sum()) - n_contact_flip, n_contact_flip, int(clamped[S - 1].sum()),
        float((e_q / (TOL_Q[:S, None] * vscale))[strict].max()) if strict.any() else 0.0,
        float((e_v / (TOL_V[:S, None] * vscale))[strict].max()) if strict.any() else 0.0))
    assert flipped.mean() <= max_flip_frac, "%s: %d of %d envs took a different discrete decision" % (label, flipped.sum(), n)
    assert n_contact_flip <= max(1, max_contact_flip_frac * n), "%s: %d contact-set flips" % (label, n_contact_flip)
    return strict[S - 1]


def golden_params(g):
    prm = {"stop_frames": int(g["stop_frames"])}
    prm["max_episode_steps"] = int(g["time_limit"]) if int(g["time_limit"]) else 1 << 30
    return prm


def replay_golden(g, b, obs_tol=2e-6, rew_tol=1e-5, terrain_tol=1e-12, pot_tol=1e-9, ts_stride=1009):
    """Replay one golden tape (an open npz) through backend ``b`` and compare with what the REFERENCE recorded.
    obs_tol is relative to max(1, |ref|) (x 5.7e4 on Mike's locked-abdomen columns, whose normalised angle divides by a
    3.5e-5 rad range in float32).  Returns the number of steps replayed."""
    from steppingstone_b200 import seeding
    kind = int(g["kind"])
    if int(g["curriculum"]) >= 0:
        b.set_sample_cdf(seeding.choice_cdf(seeding.curriculum_table(int(g["curriculum"]))))
    obs = b.reset()[0]
    scale = np.ones(60)
    if kind == 1:
        scale[6:9] = 5.7e4

    def close(a, ref, what, t):
        err = np.abs(a.astype(np.float64) - ref.astype(np.float64)) / (scale * np.maximum(1.0, np.abs(ref)))
        k = int(err.argmax())
        assert err.max() < obs_tol, "%s differs at step %d: col %d ref %r got %r" % (what, t, k, ref[k], a[k])

    close(obs, g["obs0"], "reset obs", -1)
    terr_at = list(g["terrain_at"]); terr = g["terrain"]
    assert np.abs(b.get_state()["terrain"][0] - terr[0]).max() < terrain_tol, "terrain after the first reset"
    T = len(g["rew"]); n_ep = 0
    ts_step = list(g["ts_step"]) if "ts_step" in g.files else []
    sp_step = list(g["sp_step"]) if "sp_step" in g.files else []
    restore = seeding.temp_state_restore(kind)
    for t in range(T):
        ob, rew, done, to, er, el = b.step_injected(g["actions"][t][None], g["q"][t][None], g["qd"][t][None], g["contact"][t][None])
        assert bool(done[0]) == bool(g["done"][t]), "done differs at step %d" % t
        assert bool(to[0]) == bool(g["bad"][t]), "bad_transition differs at step %d" % t
        assert abs(float(rew[0]) - g["rew"][t]) < rew_tol * max(1.0, abs(g["rew"][t])), "reward differs at step %d: ref %r got %r" % (t, g["rew"][t], rew[0])
        close(ob[0], g["obs"][t], "obs", t)
        if t in sp_step:        # update_sample_prob (mike.py:827-830): new table, immediate re-sample iff this step advanced a stone
            b.set_sample_cdf(seeding.choice_cdf(g["sp_tables"][sp_step.index(t)].astype(np.float64)), apply_now=True)
        if t in ts_step:        # value-guided sampling probes (mike.py:699-753); side effects show in the NEXT step's reward
            k = ts_step.index(t); variant = int(g["ts_variant"][k])
            ts = b.create_temp_states(variant, g["ts_grid"], restore)[0]
            ref = g["ts_%d" % k]
            if variant == 5:
                assert np.abs(ts.astype(np.float64).sum(axis=0) - g["ts_sum_%d" % k]).max() < max(obs_tol, 2e-6) * 161051 * 5
                ts = ts[::ts_stride]
            assert ts.shape == ref.shape
            err = np.abs(ts.astype(np.float64) - ref) / (scale * np.maximum(1.0, np.abs(ref)))
            assert err.max() < obs_tol, "temp states (variant %d) differ at step %d: %r" % (variant, t, err.max())
        if done[0]:
            n_ep += 1
            assert int(el[0]) == int(g["ep_l"][t])
            assert abs(float(er[0]) - g["ep_r"][t]) < 2 * rew_tol * max(1.0, abs(g["ep_r"][t]))
        if t in terr_at:
            st = b.get_state()
            assert np.abs(st["terrain"][0] - terr[terr_at.index(t)]).max() < terrain_tol, "terrain_info differs after step %d" % t
            if not done[0]:
                assert int(st["ints"][0, 0]) == int(g["next_step"][t])
                assert int(st["ints"][0, 1]) == int(g["reached_count"][t])
                assert abs(st["scalars"][0, 0] - g["potential"][t]) < pot_tol * max(1.0, abs(g["potential"][t]))
    assert n_ep == int(g["done"].sum())
    return T
