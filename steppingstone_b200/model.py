"""Model compiler: MJCF humanoid + stone URDF -> constant tables (``ss_model`` / ``ss_params``).

What the reference does at load time (and PyBullet's importers do for it) is restated here as
plain arithmetic so that the CPU oracle and the CUDA kernels consume the *same* numbers:

* ``WalkerBase.load_robot_model`` / ``parse_joints_and_links``  (mocca_envs/robots.py:371-431)
  -> joint order, limits, links;  ``Walker3D``/``Mike`` tables (robots.py:446-468, 702-724, 786-791).
* ``loadMJCF`` (PyBullet MJCF importer; not in the reference tree -- restated from its published
  behaviour): mass = 1000 kg/m^3 x geom volume; the link's inertial frame sits at the midpoint of the
  link's last ``fromto`` geom (body origin when the link has none); the inertia is the box inertia of the
  axis-aligned bounds of the link's collision shapes in that frame (btCompoundShape::calculateLocalInertia);
  ``armature`` is ignored; ``damping`` becomes a joint damping torque.
* ``Plank`` (mocca_envs/bullet_objects.py:44-93, data/objects/steps/plank.urdf).

The compiled tables are committed as JSON under ``steppingstone_b200/data`` (numbers only) so nothing at
run time reads the reference tree; ``tools/compile_models.py`` regenerates them.
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

NB, NJ, NDOF, NQ, NCOL, GRID = 22, 21, 27, 28, 4, 11
NGEOM, MAX_PAIRS = 22, 160
PARENT = [-1, 0, 1, 2, 3, 4, 5, 6, 7, 3, 9, 10, 11, 12, 0, 14, 15, 16, 0, 18, 19, 20]
JOINT_NAMES = [
    "abdomen_z", "abdomen_y", "abdomen_x",
    "right_hip_x", "right_hip_z", "right_hip_y", "right_knee", "right_ankle",
    "left_hip_x", "left_hip_z", "left_hip_y", "left_knee", "left_ankle",
    "right_shoulder_x", "right_shoulder_z", "right_shoulder_y", "right_elbow",
    "left_shoulder_x", "left_shoulder_z", "left_shoulder_y", "left_elbow",
]
# mocca_envs/robots.py:446-468 (Walker3D) and :702-724 (Mike)
TAU_W3D = [60, 80, 60, 80, 60, 100, 90, 60, 80, 60, 100, 90, 60, 60, 60, 50, 60, 60, 60, 50, 60]
TAU_MIKE = [0, 0, 0, 80, 60, 100, 90, 60, 80, 60, 100, 90, 60, 30, 30, 25, 30, 30, 30, 25, 30]
PD_P = [60, 80, 60, 80, 60, 100, 90, 60, 80, 60, 100, 90, 60, 60, 60, 50, 60, 60, 60, 50, 60]  # robots.py:482
DEG = math.pi / 180.0
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class SSModel(ctypes.Structure):
    """ctypes mirror of ``ss_model`` (include/ss_model.h)."""
    _fields_ = [
        ("parent", ctypes.c_int32 * NB), ("depth", ctypes.c_int32 * NB),
        ("anchor", ctypes.c_double * 3 * NB), ("R0", ctypes.c_double * 9 * NB),
        ("axis", ctypes.c_double * 3 * NB), ("mass", ctypes.c_double * NB),
        ("com", ctypes.c_double * 3 * NB), ("inertia", ctypes.c_double * 6 * NB),
        ("q_lo", ctypes.c_double * NJ), ("q_hi", ctypes.c_double * NJ),
        ("tau_max", ctypes.c_double * NJ), ("joint_damping", ctypes.c_double * NJ),
        ("pd_p", ctypes.c_double * NJ), ("pd_d", ctypes.c_double * NJ),
        ("col_body", ctypes.c_int32 * NCOL), ("col_foot", ctypes.c_int32 * NCOL),
        ("col_radius", ctypes.c_double * NCOL),
        ("col_p0", ctypes.c_double * 3 * NCOL), ("col_p1", ctypes.c_double * 3 * NCOL),
        ("foot_body", ctypes.c_int32 * 2), ("foot_ref", ctypes.c_double * 3 * 2),
        ("contact_breaking", ctypes.c_double),
        ("init_q", ctypes.c_double * NJ), ("init_pos", ctypes.c_double * 3),
        ("total_mass", ctypes.c_double),
        ("n_geom", ctypes.c_int32), ("n_pair", ctypes.c_int32),
        ("geom_body", ctypes.c_int32 * NGEOM), ("geom_foot", ctypes.c_int32 * NGEOM),
        ("geom_stone", ctypes.c_int32 * NGEOM), ("geom_sphere", ctypes.c_int32 * NGEOM),
        ("geom_radius", ctypes.c_double * NGEOM), ("geom_p0", ctypes.c_double * 3 * NGEOM),
        ("geom_p1", ctypes.c_double * 3 * NGEOM), ("geom_breaking", ctypes.c_double * NGEOM),
        ("pair_a", ctypes.c_int32 * MAX_PAIRS), ("pair_b", ctypes.c_int32 * MAX_PAIRS),
    ]


class SSParams(ctypes.Structure):
    """ctypes mirror of ``ss_params`` (include/ss_model.h)."""
    _fields_ = [
        ("gravity", ctypes.c_double), ("dt", ctypes.c_double),
        ("n_substeps", ctypes.c_int32), ("n_iters", ctypes.c_int32),
        ("erp_limit", ctypes.c_double), ("contact_stiffness", ctypes.c_double),
        ("contact_damping", ctypes.c_double), ("friction", ctypes.c_double),
        ("linear_slop", ctypes.c_double), ("warmstart", ctypes.c_double),
        ("lin_damping", ctypes.c_double), ("ang_damping", ctypes.c_double),
        ("max_coord_vel", ctypes.c_double), ("limit_max_impulse", ctypes.c_double),
        ("stone_half", ctypes.c_double * 3), ("stone_box_z", ctypes.c_double),
        ("stone_com_z", ctypes.c_double), ("broadphase_margin", ctypes.c_double),
        ("robot_kind", ctypes.c_int32), ("terrain_mode", ctypes.c_int32),
        ("n_steps", ctypes.c_int32), ("rendered", ctypes.c_int32),
        ("stop_frames", ctypes.c_int32), ("max_episode_steps", ctypes.c_int32),
        ("n_phantoms", ctypes.c_int32), ("use_sampler", ctypes.c_int32),
        ("actuator_mode", ctypes.c_int32), ("collision_mode", ctypes.c_int32),
        ("done_height", ctypes.c_double), ("electricity_cost", ctypes.c_double),
        ("stall_torque_cost", ctypes.c_double), ("joints_at_limit_cost", ctypes.c_double),
        ("step_radius", ctypes.c_double), ("r_max", ctypes.c_double),
        ("tilt_limit_deg", ctypes.c_double), ("mirror_prob", ctypes.c_double),
        ("control_dt", ctypes.c_double),
        ("yaw_samples", ctypes.c_double * GRID), ("pitch_samples", ctypes.c_double * GRID),
        ("r_samples", ctypes.c_double * GRID), ("xtilt_samples", ctypes.c_double * GRID),
        ("ytilt_samples", ctypes.c_double * GRID),
        ("self_friction", ctypes.c_double), ("self_erp", ctypes.c_double), ("stone_breaking", ctypes.c_double),
        ("self_cross_terms", ctypes.c_int32), ("reserved1", ctypes.c_int32),
    ]


# --------------------------------------------------------------------------------------------
# MJCF parsing
# --------------------------------------------------------------------------------------------
def _vec(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None:
        assert v.size == n, (s, n)
    return v


def _quat_wxyz_to_mat(q):
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def _shortest_arc_z_to(d):
    """Rotation taking +z onto unit vector d (how the importer orients a fromto capsule)."""
    z = np.array([0.0, 0.0, 1.0])
    c = float(np.dot(z, d))
    if c > 1.0 - 1e-12:
        return np.eye(3)
    if c < -1.0 + 1e-12:
        return np.diag([1.0, -1.0, -1.0])
    k = np.cross(z, d)
    s = np.linalg.norm(k)
    k = k / s
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + s * K + (1 - c) * (K @ K)


class _Geom:
    def __init__(self, kind, radius, p0, p1, name, contype=3, conaffinity=3):
        self.kind, self.radius, self.p0, self.p1, self.name = kind, radius, p0, p1, name
        self.contype, self.conaffinity = contype, conaffinity

    def volume(self):
        r = self.radius
        v = 4.0 / 3.0 * math.pi * r ** 3
        if self.kind == "capsule":
            v += math.pi * r * r * float(np.linalg.norm(self.p1 - self.p0))
        return v

    def centre(self):
        return 0.5 * (self.p0 + self.p1)

    def aabb(self, origin):
        """Axis-aligned bounds of the collision shape as Bullet reports them (sphere: +-r; capsule: the
        bounds of its oriented bounding box (r, r, r + L/2))."""
        c = self.centre() - origin
        r = self.radius
        if self.kind == "sphere":
            ext = np.array([r, r, r])
        else:
            d = self.p1 - self.p0
            L = float(np.linalg.norm(d))
            Rm = _shortest_arc_z_to(d / L)
            ext = np.abs(Rm) @ np.array([r, r, r + 0.5 * L])
        return c - ext, c + ext


class _Link:
    """One MJCF <body>: frame (pos, rot) in the parent body frame, hinges, geoms, fixed children."""
    def __init__(self, name, pos, rot):
        self.name, self.pos, self.rot = name, pos, rot
        self.joints, self.geoms, self.children = [], [], []


def _parse_body(el, gdef=(3, 3)):
    pos = _vec(el.get("pos", "0 0 0"), 3)
    rot = _quat_wxyz_to_mat(_vec(el.get("quat"), 4)) if el.get("quat") else np.eye(3)
    link = _Link(el.get("name"), pos, rot)
    for ch in el:
        if ch.tag == "joint":
            assert ch.get("type", "hinge") == "hinge"
            rng = _vec(ch.get("range"), 2) * DEG          # <compiler angle="degree">
            ax = _vec(ch.get("axis"), 3)
            link.joints.append(dict(name=ch.get("name"), axis=ax / np.linalg.norm(ax),
                                    pos=_vec(ch.get("pos", "0 0 0"), 3), lo=rng[0], hi=rng[1]))
        elif ch.tag == "geom":
            size = _vec(ch.get("size"))
            ct, ca = int(ch.get("contype", gdef[0])), int(ch.get("conaffinity", gdef[1]))
            if ch.get("type") == "sphere":
                p = _vec(ch.get("pos", "0 0 0"), 3)
                link.geoms.append(_Geom("sphere", size[0], p, p.copy(), ch.get("name"), ct, ca))
            elif ch.get("type") == "capsule":
                ft = _vec(ch.get("fromto"), 6)
                link.geoms.append(_Geom("capsule", size[0], ft[:3], ft[3:], ch.get("name"), ct, ca))
            else:
                raise ValueError("unsupported geom type %r" % ch.get("type"))
        elif ch.tag == "body":
            link.children.append(_parse_body(ch, gdef))
    return link


def _link_inertial(link, density=1000.0, mass_override=None):
    """(mass, com, I3x3 about com) of one MJCF body the way the MJCF importer + URDF2Bullet build it."""
    mass = density * sum(g.volume() for g in link.geoms)
    if mass_override is not None:
        mass = mass_override
    shift = np.zeros(3)
    for g in link.geoms:                   # the last fromto geom wins
        if g.kind == "capsule":
            shift = g.centre()
    if not link.geoms:
        return 0.0, shift, np.zeros((3, 3))
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for g in link.geoms:
        a, b = g.aabb(shift)
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
    l = hi - lo
    I = mass / 12.0 * np.array([l[1] ** 2 + l[2] ** 2, l[0] ** 2 + l[2] ** 2, l[0] ** 2 + l[1] ** 2])
    return mass, shift, np.diag(I)


def _merge(m1, c1, I1, m2, c2, I2):
    """Composite of two rigid bodies given in the same frame (I about each own COM)."""
    m = m1 + m2
    if m == 0.0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    c = (m1 * c1 + m2 * c2) / m

    def shift(I, mm, d):
        return I + mm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    return m, c, shift(I1, m1, c1 - c) + shift(I2, m2, c2 - c)


def compile_mjcf(xml_path, robot_kind):
    """Parse ``walker3d.xml`` / ``mike.xml`` into a dict of numpy tables (see ``ss_model``)."""
    root = ET.parse(xml_path).getroot()
    assert root.find("compiler").get("angle") == "degree"
    jdef = root.find("default").find("joint")
    damping = float(jdef.get("damping"))
    gdef_el = root.find("default").find("geom")
    gdef = (int(gdef_el.get("contype", 1)), int(gdef_el.get("conaffinity", 1)))
    base_el = root.find("worldbody").find("body")
    base = _parse_body(base_el, gdef)

    T = dict(
        parent=np.array(PARENT, dtype=np.int32), depth=np.zeros(NB, np.int32),
        anchor=np.zeros((NB, 3)), R0=np.tile(np.eye(3).reshape(1, 9), (NB, 1)),
        axis=np.zeros((NB, 3)), mass=np.zeros(NB), com=np.zeros((NB, 3)),
        inertia=np.zeros((NB, 6)), q_lo=np.zeros(NJ), q_hi=np.zeros(NJ),
        joint_damping=np.full(NJ, damping),
    )
    I3 = [np.zeros((3, 3)) for _ in range(NB)]
    names = []
    geoms_of_body = {}
    # every collision geom: (MJCF body name, our body index, p0, p1 in OUR body frame, geom); and the MJCF body tree
    all_geoms = []
    mj_parent = {}
    link_disc = {}

    def note_link(link, body_idx, offset, rot, parent_name):
        mj_parent[link.name] = parent_name
        for g in link.geoms:
            all_geoms.append(dict(link=link.name, body=body_idx, p0=offset + rot @ g.p0, p1=offset + rot @ g.p1,
                                  radius=g.radius, kind=g.kind, name=g.name))
        if link.geoms:
            # btCollisionShape::getAngularMotionDisc of the link's compound shape (about its inertial frame)
            _, shift, _ = _link_inertial(link)
            lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
            for g in link.geoms:
                a, b = g.aabb(shift)
                lo, hi = np.minimum(lo, a), np.maximum(hi, b)
            link_disc[link.name] = 0.5 * np.linalg.norm(hi - lo) + np.linalg.norm(0.5 * (hi + lo))
            link_filter[link.name] = (link.geoms[-1].contype, link.geoms[-1].conaffinity)   # the last geom wins
    link_filter = {}
    foot_ref = {}
    mass_override = {"waist": 8.0} if robot_kind == 1 else {}     # robots.py:786-791

    def add_fixed(body_idx, link, offset, rot, parent_name):
        """Merge a jointless child (head, torso, hand) into body ``body_idx``; offset/rot place the child's
        frame in that body's frame."""
        m, c, I = _link_inertial(link, mass_override=mass_override.get(link.name))
        c_b = offset + rot @ c
        I_b = rot @ I @ rot.T
        T["mass"][body_idx], T["com"][body_idx], I3[body_idx] = _merge(
            T["mass"][body_idx], T["com"][body_idx], I3[body_idx], m, c_b, I_b)
        note_link(link, body_idx, offset, rot, parent_name)
        for ch in link.children:
            assert not ch.joints
            add_fixed(body_idx, ch, offset + rot @ ch.pos, rot @ ch.rot, link.name)

    def walk(link, parent_idx, parent_frame_origin_in_parent_mjcf, parent_name=None):
        """``parent_frame_origin_in_parent_mjcf``: where OUR frame of the parent body sits inside the
        parent's MJCF body frame (its last hinge anchor; zero for the base)."""
        if not link.joints:               # base
            idx = 0
            names.append(link.name)
            m, c, I = _link_inertial(link)
            T["mass"][0], T["com"][0], I3[0] = m, c, I
            geoms_of_body[0] = (link, np.zeros(3))
            my_origin = np.zeros(3)
        else:
            # a body with k hinges becomes k links; the first k-1 are massless
            # (one Bullet link per hinge, dummy links first).
            idx = None
            prev = parent_idx
            prev_origin = parent_frame_origin_in_parent_mjcf
            for k, j in enumerate(link.joints):
                idx = len(names)
                assert PARENT[idx] == prev, (link.name, j["name"], idx, prev)
                assert JOINT_NAMES[idx - 1] == j["name"], (JOINT_NAMES[idx - 1], j["name"])
                names.append(j["name"])
                if k == 0:
                    T["anchor"][idx] = link.pos + link.rot @ j["pos"] - prev_origin
                    T["R0"][idx] = link.rot.reshape(9)
                else:
                    T["anchor"][idx] = j["pos"] - link.joints[k - 1]["pos"]
                T["axis"][idx] = j["axis"]
                T["q_lo"][idx - 1], T["q_hi"][idx - 1] = j["lo"], j["hi"]
                T["depth"][idx] = T["depth"][prev] + 1
                prev = idx
            my_origin = link.joints[-1]["pos"]
            m, c, I = _link_inertial(link, mass_override=mass_override.get(link.name))
            T["mass"][idx], T["com"][idx], I3[idx] = m, c - my_origin, I
            geoms_of_body[idx] = (link, my_origin)
            if link.name in ("right_foot", "left_foot"):
                foot_ref[link.name] = (idx, c - my_origin)
        note_link(link, idx, -my_origin, np.eye(3), parent_name)
        for ch in link.children:
            if ch.joints:
                walk(ch, idx, my_origin, link.name)
            else:
                add_fixed(idx, ch, ch.pos - my_origin, ch.rot, link.name)
        return idx

    walk(base, -1, np.zeros(3))
    assert len(names) == NB, names
    for i in range(NB):
        I = I3[i]
        T["inertia"][i] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]

    # foot colliders: the two capsules of each foot (walker3d.xml:44-45, 60-61)
    col_body, col_foot, col_r, col_p0, col_p1 = [], [], [], [], []
    for f, fname in enumerate(("right_foot", "left_foot")):
        bidx, _ = foot_ref[fname]
        link, origin = geoms_of_body[bidx]
        for g in link.geoms:
            assert g.kind == "capsule"
            col_body.append(bidx); col_foot.append(f); col_r.append(g.radius)
            col_p0.append(g.p0 - origin); col_p1.append(g.p1 - origin)
    assert len(col_body) == NCOL
    T.update(col_body=np.array(col_body, np.int32), col_foot=np.array(col_foot, np.int32),
             col_radius=np.array(col_r), col_p0=np.array(col_p0), col_p1=np.array(col_p1))
    T["foot_body"] = np.array([foot_ref["right_foot"][0], foot_ref["left_foot"][0]], np.int32)
    T["foot_ref"] = np.array([foot_ref["right_foot"][1], foot_ref["left_foot"][1]])

    # manifold breaking threshold: 0.02 x angular-motion disc of the foot's compound shape
    link, origin = geoms_of_body[int(T["foot_body"][0])]
    shift = T["foot_ref"][0] + origin
    lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
    for g in link.geoms:
        a, b = g.aabb(shift)
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
    disc = 0.5 * np.linalg.norm(hi - lo) + np.linalg.norm(0.5 * (hi + lo))
    T["contact_breaking"] = 0.02 * disc


    T["tau_max"] = np.array(TAU_W3D if robot_kind == 0 else TAU_MIKE, dtype=np.float64)
    T["pd_p"] = np.array(PD_P, dtype=np.float64)
    T["pd_d"] = T["pd_p"] / 10.0
    q0 = np.zeros(NJ)                      # robots.py:535-545 "running_start"
    q0[[5, 6]] = -math.pi / 8
    q0[10] = math.pi / 10
    q0[[13, 17]] = math.pi / 3
    q0[14] = -math.pi / 6
    q0[18] = math.pi / 6
    q0[[16, 20]] = math.pi / 3
    T["init_q"] = q0
    T["init_pos"] = np.array([0.3, 0.0, 1.25 if robot_kind == 0 else 1.0])
    T["total_mass"] = float(T["mass"].sum())
    T["body_names"] = names

    # ---- collision geoms of every link (URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS, robots.py:496-500) ----
    # order: the four foot capsules first (same numbering as col_*), then the other capsules, then the spheres
    def order_key(g):
        foot = g["link"] in ("right_foot", "left_foot")
        return (0 if foot else (1 if g["kind"] == "capsule" else 2))
    geoms = sorted(all_geoms, key=order_key)          # stable: XML order inside each class
    feet = [g for g in geoms if g["link"] in ("right_foot", "left_foot")]
    assert [g["body"] for g in feet] == list(col_body)
    assert len(geoms) <= NGEOM
    STATIC_GROUP, STATIC_MASK = 2, -1 ^ 2             # btBroadphaseProxy::StaticFilter, AllFilter ^ StaticFilter

    def ancestors(name):
        out = []
        while mj_parent.get(name) is not None:
            name = mj_parent[name]; out.append(name)
        return out
    G = dict(geom_body=np.zeros(NGEOM, np.int32), geom_foot=np.full(NGEOM, -1, np.int32), geom_stone=np.zeros(NGEOM, np.int32),
             geom_sphere=np.zeros(NGEOM, np.int32), geom_radius=np.zeros(NGEOM), geom_p0=np.zeros((NGEOM, 3)),
             geom_p1=np.zeros((NGEOM, 3)), geom_breaking=np.zeros(NGEOM))
    for k, g in enumerate(geoms):
        grp, msk = link_filter[g["link"]]
        G["geom_body"][k] = g["body"]
        G["geom_foot"][k] = {"right_foot": 0, "left_foot": 1}.get(g["link"], -1)
        G["geom_stone"][k] = int(bool(grp & STATIC_MASK) and bool(STATIC_GROUP & msk))
        G["geom_sphere"][k] = int(g["kind"] == "sphere")
        G["geom_radius"][k] = g["radius"]; G["geom_p0"][k] = g["p0"]; G["geom_p1"][k] = g["p1"]
        G["geom_breaking"][k] = 0.02 * link_disc[g["link"]]
    pa, pb = [], []
    for i in range(len(geoms)):
        for j in range(i + 1, len(geoms)):
            a, b = geoms[i], geoms[j]
            if a["link"] == b["link"] or a["body"] == b["body"]:
                continue                               # same link, or rigidly attached to the same body (constant distance)
            if a["link"] in ancestors(b["link"]) or b["link"] in ancestors(a["link"]):
                continue                               # btMultiBodyLinkCollider: BT_MULTIBODYLINKFLAGS_DISABLE_ALL_PARENT_COLLISION
            (ga, ma), (gb, mb) = link_filter[a["link"]], link_filter[b["link"]]
            if not ((ga & mb) and (gb & ma)):
                continue                               # contype / conaffinity (walker3d.xml:5,21,28,31)
            pa.append(i); pb.append(j)
    assert len(pa) <= MAX_PAIRS, len(pa)
    # order the pairs by how close they can get (smallest gap over random poses first): the GPU evaluates them 32 at a
    # time and skips a whole batch when none of its pairs is within reach, so likely pairs go first
    def fk_points(qj):
        Rw = [np.eye(3)] * NB; ow = [np.zeros(3)] * NB
        for b in range(1, NB):
            pr = PARENT[b]; ax = T["axis"][b]; th = qj[b - 1]
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            Rj = np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)
            Rw[b] = Rw[pr] @ T["R0"][b].reshape(3, 3) @ Rj
            ow[b] = ow[pr] + Rw[pr] @ T["anchor"][b]
        return [(ow[g["body"]] + Rw[g["body"]] @ g["p0"], ow[g["body"]] + Rw[g["body"]] @ g["p1"]) for g in geoms]

    def seg_dist(p1, q1, p2, q2):
        best = np.inf
        for s in np.linspace(0, 1, 5):
            for t in np.linspace(0, 1, 5):
                best = min(best, np.linalg.norm(p1 + s * (q1 - p1) - p2 - t * (q2 - p2)))
        return best
    rng = np.random.RandomState(0)
    gap = np.full(len(pa), np.inf)
    for k in range(60):
        qj = T["init_q"] if k == 0 else rng.uniform(T["q_lo"], T["q_hi"])
        pts = fk_points(qj)
        for n, (i, j) in enumerate(zip(pa, pb)):
            gap[n] = min(gap[n], seg_dist(*pts[i], *pts[j]) - geoms[i]["radius"] - geoms[j]["radius"])
    order = np.argsort(gap, kind="stable")
    pa = [pa[k] for k in order]; pb = [pb[k] for k in order]
    G["pair_a"] = np.zeros(MAX_PAIRS, np.int32); G["pair_b"] = np.zeros(MAX_PAIRS, np.int32)
    G["pair_a"][:len(pa)] = pa; G["pair_b"][:len(pb)] = pb
    G["n_geom"] = len(geoms); G["n_pair"] = len(pa)
    G["geom_names"] = [g["name"] for g in geoms]
    T.update(G)
    return T


def default_params(robot_kind, terrain_mode=None):
    """Constants the reference sets in code, as an ``ss_params``-shaped dict."""
    k_stone, k_robot = 30000.0, 1e18       # bullet_objects.py:66; btCollisionObject default
    P = dict(
        gravity=9.81, dt=1.0 / 240.0, n_substeps=4, n_iters=20,      # bullet_utils.py:313-329, env_base.py:50-55
        erp_limit=0.2,
        contact_stiffness=1.0 / (1.0 / k_stone + 1.0 / k_robot),
        contact_damping=1000.0 + 0.1,                                # bullet_objects.py:67 + default
        friction=1.2 * 1.0,                                          # walker3d.xml:5 x bullet_objects.py:64
        linear_slop=1e-5, warmstart=0.85, lin_damping=0.04, ang_damping=0.04,
        max_coord_vel=100.0, limit_max_impulse=100.0,
        stone_half=[0.25, 0.375, 0.125], stone_box_z=-0.125, stone_com_z=0.1375,
        broadphase_margin=0.35,
        robot_kind=robot_kind,
        terrain_mode=(robot_kind if terrain_mode is None else terrain_mode),
        n_steps=100 if robot_kind == 0 else 40,                      # walker3d_envs.py:507 / mike.py:50
        rendered=100 if robot_kind == 0 else 30,                     # walker3d_envs.py:497 / mike.py:40
        stop_frames=30, max_episode_steps=1000,
        n_phantoms=20 if robot_kind == 0 else 0,                     # walker3d_envs.py:489,566 / mike.py:32
        use_sampler=0 if robot_kind == 0 else 1,                     # walker3d_envs.py:1083 / mike.py:593
        actuator_mode=0, collision_mode=2,                           # robots.py:496-500: self-collision on
        self_friction=1.2 * 1.2, self_erp=0.9,                       # walker3d.xml:5; bullet_utils.py:324
        stone_breaking=0.02 * (0.0125 + math.sqrt(0.25 ** 2 + 0.375 ** 2 + 0.125 ** 2)),
        self_cross_terms=0, reserved1=0,
        done_height=0.7, electricity_cost=4.5, stall_torque_cost=0.225, joints_at_limit_cost=0.1,
        step_radius=0.25, r_max=1.5, tilt_limit_deg=0.0,
        mirror_prob=0.0 if robot_kind == 0 else 0.5,                 # robots.py:583 / :846
        control_dt=1.0 / 60.0,
    )
    if robot_kind == 0:                                              # walker3d_envs.py:529-533
        P["yaw_samples"] = (np.linspace(-10, -10, GRID) * DEG).tolist()
        P["pitch_samples"] = (np.linspace(0, 0, GRID) * DEG).tolist()
    else:                                                            # mike.py:72-76
        P["yaw_samples"] = (np.linspace(-20, 20, GRID) * DEG).tolist()
        P["pitch_samples"] = (np.linspace(-50, 50, GRID) * DEG).tolist()
    P["r_samples"] = np.linspace(0.65, 0.65, GRID).tolist()
    P["xtilt_samples"] = (np.linspace(0, 0, GRID) * DEG).tolist()
    P["ytilt_samples"] = (np.linspace(0, 0, GRID) * DEG).tolist()
    return P


# --------------------------------------------------------------------------------------------
# (de)serialisation
# --------------------------------------------------------------------------------------------
ENV_IDS = {
    "mocca_envs:Walker3DStepperEnv-v0": 0, "Walker3DStepperEnv-v0": 0, "walker3d": 0,
    "mocca_envs:MikeStepperEnv-v0": 1, "MikeStepperEnv-v0": 1, "mike": 1,
}


def tables_to_json(T, path):
    out = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in T.items()}
    with open(path, "w") as f:
        json.dump(out, f, indent=0)


def load_tables(robot_kind):
    name = "walker3d.json" if robot_kind == 0 else "mike.json"
    with open(os.path.join(DATA_DIR, name)) as f:
        raw = json.load(f)
    T = {}
    for k, v in raw.items():
        if k in ("body_names", "geom_names"):
            T[k] = v
        elif k in ("n_geom", "n_pair"):
            T[k] = int(v)
        elif k in ("parent", "depth", "col_body", "col_foot", "foot_body", "geom_body", "geom_foot", "geom_stone",
                   "geom_sphere", "pair_a", "pair_b"):
            T[k] = np.array(v, dtype=np.int32)
        elif isinstance(v, list):
            T[k] = np.array(v, dtype=np.float64)
        else:
            T[k] = v
    assert list(T["parent"]) == PARENT, "model topology differs from the compiled-in humanoid tree"
    return T


def _fill(struct, name, value):
    field = getattr(struct, name)
    if isinstance(field, (int, float)):
        setattr(struct, name, value)
        return
    flat = np.ascontiguousarray(value).reshape(-1)
    nbytes = ctypes.sizeof(field)
    assert nbytes % flat.size == 0, (name, nbytes, flat.size)
    data = flat.astype(np.int32 if nbytes // flat.size == 4 else np.float64).tobytes()
    assert len(data) == nbytes, (name, len(data), nbytes)
    ctypes.memmove(ctypes.addressof(field), data, nbytes)


def make_model(robot_kind, tables=None):
    T = load_tables(robot_kind) if tables is None else tables
    m = SSModel()
    for name, _ in SSModel._fields_:
        _fill(m, name, T[name])
    return m


def make_params(robot_kind, overrides=None, terrain_mode=None):
    P = default_params(robot_kind, terrain_mode)
    if overrides:
        for k, v in overrides.items():
            if k not in P:
                raise KeyError("unknown ss_params field %r" % k)
            P[k] = v
    p = SSParams()
    for name, _ in SSParams._fields_:
        _fill(p, name, P[name])
    return p
