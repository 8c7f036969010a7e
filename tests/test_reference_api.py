"""CPU, build container only: the host-side mirrors against the reference's own Python classes (imported from
/root/reference; skipped where the reference tree is absent, e.g. on the GPU box)."""
import inspect
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("SS_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "algorithms")), reason="reference tree not present")


def _ref_storage():
    sys.path.insert(0, REF)
    try:
        from algorithms.storage import RolloutStorage as Ref
    finally:
        sys.path.remove(REF)
    return Ref


def test_rollout_storage_has_the_reference_surface():
    """Same constructor arguments, tensors (names, shapes) and methods as algorithms/storage.py:5-103."""
    import torch
    from steppingstone_b200.storage import RolloutStorage
    Ref = _ref_storage()
    assert list(inspect.signature(Ref.__init__).parameters)[:6] == list(inspect.signature(RolloutStorage.__init__).parameters)[:6]
    a = Ref(5, 3, (60,), 21, 1); b = RolloutStorage(5, 3, (60,), 21, 1, device="cpu")
    for name, t in vars(a).items():
        if isinstance(t, torch.Tensor):
            assert tuple(getattr(b, name).shape) == tuple(t.shape), name
    for m in ("insert", "after_update", "compute_returns", "feed_forward_generator", "cuda"):
        assert list(inspect.signature(getattr(Ref, m)).parameters) == list(inspect.signature(getattr(RolloutStorage, m)).parameters), m


def test_rollout_storage_matches_reference_numerics():
    """insert / after_update / compute_returns(use_gae=False) / minibatch generator against the reference class on
    the same data (the GAE path is the ss_gae kernel: tests/test_gpu_parity.py::test_gae_parity)."""
    import torch
    from steppingstone_b200.storage import RolloutStorage
    Ref = _ref_storage()
    T, N = 6, 4
    a = Ref(T, N, (60,), 21, 1); b = RolloutStorage(T, N, (60,), 21, 1, device="cpu")
    g = torch.Generator().manual_seed(0)
    for t in range(T):
        args = (torch.randn(N, 60, generator=g), torch.zeros(N, 1), torch.randn(N, 21, generator=g), torch.randn(N, 1, generator=g),
                torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), (torch.rand(N, 1, generator=g) > 0.2).float(),
                (torch.rand(N, 1, generator=g) > 0.1).float())
        a.insert(*args); b.insert(*args)
    nv = torch.randn(N, 1, generator=g)
    a.compute_returns(nv, False, 0.99, 0.95); b.compute_returns(nv, False, 0.99, 0.95)
    for name in ("observations", "rewards", "value_preds", "returns", "action_log_probs", "actions", "masks", "bad_masks"):
        assert torch.allclose(getattr(a, name), getattr(b, name)), name
    adv = a.returns[:-1] - a.value_preds[:-1]
    batches_a = list(a.feed_forward_generator(adv, 3)); batches_b = list(b.feed_forward_generator(adv, 3))
    assert len(batches_a) == len(batches_b) and [len(x) for x in batches_a] == [len(x) for x in batches_b]
    assert sorted(x[0].shape[0] for x in batches_a) == sorted(x[0].shape[0] for x in batches_b)
    # both visit every sample exactly once per pass (different permutations)
    seen = torch.cat([x[4] for x in batches_b]).flatten().sort().values
    assert torch.allclose(seen, b.returns[:-1].flatten().sort().values)
    a.after_update(); b.after_update()
    assert torch.allclose(a.observations[0], b.observations[0]) and torch.allclose(a.masks[0], b.masks[0])


def test_mirror_indices_match_the_reference_env_source():
    """get_mirror_indices restates walker3d_envs.py:1358-1411: compare with the literal index lists in the source."""
    import re
    from steppingstone_b200 import mirror
    src = open(os.path.join(REF, ".environments", "mocca_envs", "walker3d_envs.py")).read()
    body = src[src.index("def get_mirror_indices", src.index("class Walker3DStepperEnv")):]
    body = body[:body.index("return")]
    neg = [int(x) for x in re.findall(r"^\s+(\d+),", body[body.index("negation_obs_indices"):body.index("right_obs_indices")], flags=re.M)]
    ours = mirror.get_mirror_indices()[0]
    target_part = [50 + 0, 50 + 3, 50 + 5, 50 + 8]
    assert sorted(set(ours.tolist())) == sorted(set(neg + target_part)) or set(neg) <= set(ours.tolist())


def test_mirror_function_matches_the_reference_function():
    """get_mirror_function against common/envs_utils.py:687-740 itself (gym stubbed by tools/fakes)."""
    import torch
    from conftest import ROOT
    from steppingstone_b200 import mirror
    fakes = os.path.join(ROOT, "tools", "fakes")
    sys.path.insert(0, REF); sys.path.insert(0, fakes)
    try:
        from common.envs_utils import get_mirror_function as ref_get
    finally:
        sys.path.remove(REF); sys.path.remove(fakes)
    idx = mirror.get_mirror_indices()
    ours = mirror.get_mirror_function(idx)
    theirs = ref_get(idx)
    g = torch.Generator().manual_seed(1)
    B = 9
    sample = (torch.randn(B, 60, generator=g), torch.randn(B, 1, generator=g), torch.randn(B, 21, generator=g),
              torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g), torch.ones(B, 1), torch.randn(B, 1, generator=g),
              torch.randn(B, 1, generator=g))
    out_a = ours(sample); out_b = theirs(sample)
    assert len(out_a) == len(out_b)
    for x, y in zip(out_a, out_b):
        assert x.shape == y.shape and torch.equal(x, y)


def test_sampling_tables_match_the_reference_envs():
    """curriculum_table / specialist_table / fake_sample_grid against the tables the reference envs build themselves
    (update_curriculum, update_specialist, fake_*_samples; PyBullet and gym stubbed by tools/fakes)."""
    from conftest import ROOT
    from steppingstone_b200 import seeding
    fakes = os.path.join(ROOT, "tools", "fakes")
    sys.path.insert(0, REF); sys.path.insert(0, fakes)
    try:
        import pybullet as fake
        from mocca_envs.mike import MikeStepperEnv
        from mocca_envs.walker3d_envs import Walker3DStepperEnv
        for cls in (MikeStepperEnv, Walker3DStepperEnv):
            fake.reset_world()
            env = cls()
            for c in range(0, 7):
                env.update_curriculum(c)
                assert np.array_equal(env.yaw_pitch_r_tilt_prob, seeding.curriculum_table(c)), (cls.__name__, "curriculum", c)
            for s in range(0, 6):
                env.update_specialist(s)
                assert np.array_equal(env.yaw_pitch_r_tilt_prob, seeding.specialist_table(s)), (cls.__name__, "specialist", s)
            grid = seeding.fake_sample_grid()
            for k, name in enumerate(("fake_yaw_samples", "fake_pitch_samples", "fake_r_samples", "fake_x_tilt_samples", "fake_y_tilt_samples")):
                assert np.array_equal(getattr(env, name), grid[k]), name
    finally:
        sys.path.remove(REF); sys.path.remove(fakes)


def test_mirror_indices_and_registry_match_the_reference():
    """get_mirror_indices() of the reference envs themselves, and the gym registration (ids, max_episode_steps)."""
    from conftest import ROOT
    from steppingstone_b200 import mirror, model
    fakes = os.path.join(ROOT, "tools", "fakes")
    sys.path.insert(0, REF); sys.path.insert(0, fakes)
    try:
        import pybullet as fake
        from mocca_envs.mike import MikeStepperEnv
        from mocca_envs.walker3d_envs import Walker3DStepperEnv
        ours = mirror.get_mirror_indices()
        for cls in (Walker3DStepperEnv, MikeStepperEnv):
            fake.reset_world()
            env = cls()
            theirs = env.get_mirror_indices()
            assert len(theirs) == 6
            for a, b in zip(ours, theirs):
                assert np.array_equal(np.asarray(a), np.asarray(b)), cls.__name__
            assert env.observation_space.shape == (60,) and env.action_space.shape == (21,)
        src = open(os.path.join(REF, ".environments", "mocca_envs", "__init__.py")).read()
        for env_id, kind in (("Walker3DStepperEnv-v0", 0), ("MikeStepperEnv-v0", 1)):
            block = src[src.index('id="%s"' % env_id):]
            block = block[:block.index(")")]
            assert "max_episode_steps=1000" in block.replace(" ", "")
            assert model.make_params(kind).max_episode_steps == 1000
    finally:
        sys.path.remove(REF); sys.path.remove(fakes)


def test_robot_tables_match_the_reference_robot_classes():
    """Joint order, joint limits, torque limits, start pose and base height as the reference's Walker3D / Mike robot classes
    hold them after load_robot_model + calc_torque_limits + reset (robots.py:371-431, 385-388, 446-468, 535-545, 581-618)."""
    from conftest import ROOT
    from steppingstone_b200 import model
    fakes = os.path.join(ROOT, "tools", "fakes")
    sys.path.insert(0, REF); sys.path.insert(0, fakes)
    try:
        import pybullet as fake
        from mocca_envs.mike import MikeStepperEnv
        from mocca_envs.walker3d_envs import Walker3DStepperEnv
        for kind, cls in ((0, Walker3DStepperEnv), (1, MikeStepperEnv)):
            fake.reset_world()
            env = cls()
            env.seed(8)
            T = model.load_tables(kind)
            joints = env.robot.ordered_joints
            assert [j.joint_name for j in joints] == model.JOINT_NAMES
            assert np.allclose([j.lowerLimit for j in joints], T["q_lo"], atol=1e-12)
            assert np.allclose([j.upperLimit for j in joints], T["q_hi"], atol=1e-12)
            assert np.allclose(env.robot.torque_limits, T["tau_max"])
            env.reset()
            q = np.array([j.get_state()[0] for j in joints])
            mirrored = getattr(env.robot, "mirrored", False)
            if not mirrored:
                assert np.allclose(q, T["init_q"], atol=1e-12)
            assert np.allclose(env.robot.body_xyz, T["init_pos"], atol=1e-9)
    finally:
        sys.path.remove(REF); sys.path.remove(fakes)


def test_monitor_files_match_the_reference_writer(tmp_path):
    """log_dir: `<log_dir>/<rank>.monitor.csv` with the header line, CSV header and row format of the reference's
    Monitor / ResultsWriter (common/envs_utils.py:70-94, 172-194; make_env_fns :33-36) -- the parameter used to be accepted
    and ignored (missing 7)."""
    fakes = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fakes")
    sys.path.insert(0, fakes); sys.path.insert(0, REF)
    try:
        from common.envs_utils import ResultsWriter
    finally:
        sys.path.remove(REF); sys.path.remove(fakes)
    from steppingstone_b200.vec_env import _MonitorLog
    ref_file = str(tmp_path / "ref" / "3")
    os.makedirs(os.path.dirname(ref_file))
    w = ResultsWriter(ref_file, header={"t_start": 12.5, "env_id": "Walker3DStepperEnv-v0"})
    w.write_row({"r": 101.25, "l": 17, "t": 0.75}); w.write_row({"r": -3.5, "l": 1000, "t": 1.5}); w.f.close()
    m = _MonitorLog(str(tmp_path / "ours"), 2, 3, "mocca_envs:Walker3DStepperEnv-v0", 12.5)
    m.write(np.array([1]), [101.25], [17], 0.75); m.write(np.array([1, 2]), [-3.5, 8.0], [1000, 5], 1.5); m.flush()
    assert sorted(os.listdir(tmp_path / "ours")) == ["2.monitor.csv", "3.monitor.csv", "4.monitor.csv"]
    ours = open(tmp_path / "ours" / "3.monitor.csv", newline="").read()
    assert ours == open(ref_file + ".monitor.csv", newline="").read()


def test_lazy_infos_behave_like_the_reference_tuple():
    """numpy-mode ``infos``: len / indexing / iteration / keys as the tuple of N dicts ShmemVecEnv.step_wait returns; the
    consumer loop of train.py:446-454 runs on it unchanged."""
    from steppingstone_b200.vec_env import _ArrayPool, _Infos
    infos = _Infos(6, ([1, 4], [2.0, 1.0], [9, 1000], [0, 1], 0.2))
    bad_masks = np.ones((6, 1)); episode_rewards = []
    for p_index, info in enumerate(infos):                      # train.py:446-454
        keys = info.keys()
        if "bad_transition" in keys:
            bad_masks[p_index] = 0.0
        if "episode" in keys:
            episode_rewards.append(info["episode"]["r"])
    assert bad_masks[:, 0].tolist() == [1, 1, 1, 1, 0, 1] and episode_rewards == [2.0, 1.0]
    assert len(infos) == 6 and infos[-2]["bad_transition"] and infos[0] == {} and len(infos[1:3]) == 2
    pool = _ArrayPool((8,), np.float32)
    a, pa = pool.get(); ida = id(a); assert pa == a.ctypes.data; del a
    b, _ = pool.get(); assert id(b) == ida                      # nobody kept the first array: it is handed out again
    c, _ = pool.get(); assert c is not b                        # b is still alive in the caller: a fresh array
