"""Kinematic stand-in for the PyBullet wheel -- GOLDEN-GENERATION TOOLING ONLY (tools/gen_golden.py).

PyBullet is not installed in the build container, so the reference's env classes cannot be imported as they are.
This module implements just enough of the pybullet API for mocca_envs' Walker3DStepperEnv / MikeStepperEnv to
construct and run their OWN Python code (reset, observation, reward, termination, terrain generator, curriculum
sampler) unmodified.  It contains no dynamics: `stepSimulation` calls a hook that installs a post-physics state
supplied by the caller; link poses and contact lists are answered by hooks too.  Nothing under
steppingstone_b200/ or tests/ imports it.
"""
import math
import os
import xml.etree.ElementTree as ET

GUI, DIRECT, SHARED_MEMORY = 1, 2, 3
COV_ENABLE_GUI, COV_ENABLE_RENDERING = 1, 2
POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 2, 0, 1
MJCF_COLORS_FROM_FILE, URDF_USE_SELF_COLLISION, URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS = 512, 8, 32
GEOM_SPHERE, GEOM_BOX, GEOM_MESH = 2, 3, 5


class error(Exception):
    pass


class _Body:
    def __init__(self):
        self.pos = [0.0, 0.0, 0.0]; self.orn = [0.0, 0.0, 0.0, 1.0]
        self.lin = [0.0, 0.0, 0.0]; self.ang = [0.0, 0.0, 0.0]
        self.joints = []          # dicts: name, link, lo, hi, q, qd, torque
        self.com_offset = [0.0, 0.0, 0.0]
        self.is_robot = False


WORLD = {"bodies": {}, "next_id": 0, "step_hook": None, "link_pose_hook": None, "contact_hook": None, "steps": 0}


def reset_world():
    WORLD.update(bodies={}, next_id=0, step_hook=None, link_pose_hook=None, contact_hook=None, steps=0)


def _new_body():
    b = _Body(); i = WORLD["next_id"]; WORLD["next_id"] += 1; WORLD["bodies"][i] = b
    return i, b


def connect(*a, **k): return 0
def disconnect(*a, **k): return None
def configureDebugVisualizer(*a, **k): return None
def setGravity(*a, **k): return None
def setDefaultContactERP(*a, **k): return None
def setPhysicsEngineParameter(*a, **k): return None
def saveState(*a, **k): return 0
def changeDynamics(*a, **k): return None
def setCollisionFilterGroupMask(*a, **k): return None
def loadTexture(*a, **k): return 0
def changeVisualShape(*a, **k): return None


def loadMJCF(path, flags=0):
    """Links in the order Bullet's MJCF importer creates them: depth-first over <body>; a body with k hinges
    becomes k links (k-1 dummy links first, the last one carries the body name); a body without a joint hangs on
    a fixed joint named jointfix_*."""
    root = ET.parse(path).getroot()
    deg = root.find("compiler").get("angle") == "degree"
    base_el = root.find("worldbody").find("body")
    i, b = _new_body(); b.is_robot = True
    b.pos = [float(x) for x in base_el.get("pos", "0 0 0").split()]

    def walk(el):
        for ch in el.findall("body"):
            joints = ch.findall("joint")
            if not joints:
                b.joints.append(dict(name="jointfix_%d" % len(b.joints), link=ch.get("name"), lo=0.0, hi=-1.0, q=0.0, qd=0.0, torque=0.0))
            for k, j in enumerate(joints):
                lo, hi = [float(x) for x in j.get("range").split()]
                if deg:
                    lo, hi = math.radians(lo), math.radians(hi)
                link = ch.get("name") if k == len(joints) - 1 else "link_%s_%d" % (ch.get("name"), k)
                b.joints.append(dict(name=j.get("name"), link=link, lo=lo, hi=hi, q=0.0, qd=0.0, torque=0.0))
            walk(ch)
    walk(base_el)
    return (i,)


def loadURDF(path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=False, globalScaling=1.0, **k):
    root = ET.parse(path).getroot()
    links = root.findall("link")
    i, b = _new_body()
    org = links[0].find("inertial").find("origin").get("xyz")
    b.com_offset = [float(x) * globalScaling for x in org.split()]
    b.pos = [basePosition[k2] + b.com_offset[k2] for k2 in range(3)]     # the base pose Bullet reports is the COM frame
    b.orn = list(baseOrientation)
    for l in links[1:]:
        b.joints.append(dict(name="fixed", link=l.get("name"), lo=0.0, hi=-1.0, q=0.0, qd=0.0, torque=0.0))
    return i


def getNumJoints(body): return len(WORLD["bodies"][body].joints)


def getJointInfo(body, j):
    J = WORLD["bodies"][body].joints[j]
    return (j, J["name"].encode(), 0, 0, 0, 0, 0.0, 0.0, J["lo"], J["hi"], 0.0, 0.0, J["link"].encode(), (0, 0, 1), (0, 0, 0), (0, 0, 0, 1), j - 1)


def setJointMotorControl2(bodyIndex=None, jointIndex=None, controlMode=None, force=None, *args, **k):
    if args and bodyIndex is None:
        pass
    b = WORLD["bodies"].get(bodyIndex)
    if b is not None and controlMode == TORQUE_CONTROL:
        b.joints[jointIndex]["torque"] = force


def getJointState(body, j):
    J = WORLD["bodies"][body].joints[j]
    return (J["q"], J["qd"], (0.0,) * 6, 0.0)


def resetJointState(body, j, targetValue=0.0, targetVelocity=0.0):
    J = WORLD["bodies"][body].joints[j]
    J["q"], J["qd"] = float(targetValue), float(targetVelocity)


def getBasePositionAndOrientation(body):
    b = WORLD["bodies"].get(body)
    if b is None:
        return (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0)
    return tuple(b.pos), tuple(b.orn)


def resetBasePositionAndOrientation(body, posObj=None, ornObj=None):
    b = WORLD["bodies"].get(body)
    if b is None:
        return
    b.pos = [float(x) for x in posObj]; b.orn = [float(x) for x in ornObj]
    b.lin = [0.0, 0.0, 0.0]; b.ang = [0.0, 0.0, 0.0]      # Bullet's init-pose command zeroes the base velocities


def resetBaseVelocity(body, linearVelocity=None, angularVelocity=None):
    b = WORLD["bodies"][body]
    if linearVelocity is not None:
        b.lin = [float(x) for x in linearVelocity]
    if angularVelocity is not None:
        b.ang = [float(x) for x in angularVelocity]


def getBaseVelocity(body):
    b = WORLD["bodies"][body]
    return tuple(b.lin), tuple(b.ang)


def getLinkState(body, link, computeLinkVelocity=0, **k):
    hook = WORLD["link_pose_hook"]
    xyz = hook(body, link) if hook else None
    if xyz is None:
        xyz = (0.0, 0.0, 0.0)
    out = (tuple(xyz), (0.0, 0.0, 0.0, 1.0), (0, 0, 0), (0, 0, 0, 1), tuple(xyz), (0.0, 0.0, 0.0, 1.0))
    if computeLinkVelocity:
        out = out + ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    return out


def getContactPoints(bodyA=None, linkIndexA=None, **k):
    hook = WORLD["contact_hook"]
    return hook(bodyA, linkIndexA) if hook else ()


def stepSimulation(*a, **k):
    WORLD["steps"] += 1
    if WORLD["step_hook"]:
        WORLD["step_hook"]()


def getQuaternionFromEuler(e):
    r, p, y = e
    cr, sr, cp, sp, cy, sy = math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2), math.cos(y / 2), math.sin(y / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def getEulerFromQuaternion(q):
    x, y, z, w = q
    sarg = -2 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * math.pi, 2 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * math.pi, 2 * math.atan2(-x, y))
    return (math.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), math.asin(sarg),
            math.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z))
