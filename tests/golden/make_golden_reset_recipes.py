"""Golden "scene recipes": what the reference's OWN `reset()` asks the physics engine to build, recorded by running
`FeedingJacoEnv.reset()` / `ScratchItchJacoEnv.reset()` of the reference
package (third-party modules stubbed) against a RECORDING `pybullet`: every call is logged, bodies get consecutive ids, joint
tables are answered from the compiled URDF models (and, for `createMultiBody` bodies, from the call's own arrays in PyBullet's
depth-first numbering), poses set with `resetBasePositionAndOrientation` are remembered, all link states are the origin (so tool /
food placements come out as the OFFSETS the reference composes), IK returns zeros and nothing ever collides.
The salient calls are written to tests/golden/reset_recipes.json; tests/test_reference_reset_recipes.py holds the batched scene
builders (`feeding_batch.py`, ...) against them.

usage: python tests/golden/make_golden_reset_recipes.py [/root/reference]"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden_env_logic import Z, install_stubs, jsonable  # noqa: E402
from make_golden_feeding_semantics import qmul, qrot  # noqa: E402

ASSETS = os.path.join(ROOT, 'assistive_gym_b200', 'assets')
URDF2MODEL = {'plane.urdf': 'plane', 'j2s7s300_gym.urdf': 'jaco', 'wheelchair_jaco.urdf': 'wheelchair_jaco', 'wheelchair.urdf': 'wheelchair', 'table_tall.urdf': 'table_tall',
              'bowl.urdf': 'bowl', 'sawyer.urdf': 'sawyer', 'bed.urdf': 'bed', 'wiper.urdf': 'wiper', 'pr2_no_torso_lift_tall.urdf': 'pr2', 'tool_scratch.urdf': 'tool_scratch'}
KEEP = ('loadURDF', 'createCollisionShape', 'createMultiBody', 'createConstraint', 'changeConstraint', 'setGravity', 'setJointMotorControlArray',
        'resetBasePositionAndOrientation', 'setPhysicsEngineParameter', 'changeDynamics', 'calculateInverseKinematics', 'loadCloth', 'clothParams', 'setTimeStep')


def dfs_order(parents):
    kids = {i: [] for i in range(len(parents) + 1)}
    for i, p in enumerate(parents):
        kids[int(p)].append(i + 1)
    order = []

    def walk(k):
        for c in kids[k]:
            order.append(c)
            walk(c)
    walk(0)
    return order


class Recorder:
    def __init__(self):
        self.calls, self.bodies, self.pose, self.shapes, self.ik_calls, self.ik_goal = [], [], {}, [], 0, {}

    def install(self):
        R = self

        def generic(name):
            def f(*a, **k):
                k.pop('physicsClientId', None)
                R.calls.append((name, a, k))
                return Z(0)
            return f

        class M(types.ModuleType):
            def __getattr__(s, k):
                if k.startswith('__'):
                    raise AttributeError(k)
                if k.isupper() or k.split('_')[0] in ('GEOM', 'JOINT', 'COV', 'URDF'):
                    return Z(0)
                return generic(k)
        m = M('pybullet')
        sys.modules['pybullet'] = m
        m.JOINT_REVOLUTE, m.JOINT_PRISMATIC, m.JOINT_FIXED = 0, 1, 4
        m.GEOM_SPHERE, m.GEOM_BOX, m.GEOM_CYLINDER, m.GEOM_MESH, m.GEOM_PLANE, m.GEOM_CAPSULE = 2, 3, 4, 5, 6, 7

        def new_body(kind, info):
            R.bodies.append(dict(kind=kind, **info))
            bid = len(R.bodies) - 1
            R.pose[bid] = (np.array(info.get('pos', [0, 0, 0]), float), np.array(info.get('orn', [0, 0, 0, 1]), float))
            return bid

        def loadURDF(fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=0, flags=0, globalScaling=1.0, physicsClientId=None, **k):
            base = os.path.basename(fileName)
            model = json.load(open(os.path.join(ASSETS, URDF2MODEL[base] + '.agmodel.json'))) if base in URDF2MODEL else None
            bid = new_body('urdf', dict(file=base, pos=list(map(float, basePosition)), orn=list(map(float, baseOrientation)), model=model))
            R.calls.append(('loadURDF', (base,), dict(body=bid, basePosition=list(map(float, basePosition)), baseOrientation=list(map(float, baseOrientation)), useFixedBase=int(useFixedBase))))
            return bid
        m.loadURDF = loadURDF

        def createCollisionShape(shapeType=None, *a, **k):
            k.pop('physicsClientId', None)
            R.shapes.append(dict(shapeType=int(shapeType) if shapeType is not None else (int(a[0]) if a else None),
                                 **{kk: (np.asarray(v).tolist() if not isinstance(v, str) else os.path.basename(v)) for kk, v in k.items()}))
            R.calls.append(('createCollisionShape', (), dict(R.shapes[-1], shape=len(R.shapes) - 1)))
            return len(R.shapes) - 1
        m.createCollisionShape = createCollisionShape
        m.createVisualShape = lambda *a, **k: -1

        def createMultiBody(*a, **k):
            k.pop('physicsClientId', None)
            info = {kk: (np.asarray(v).tolist() if not isinstance(v, (int, float)) else v) for kk, v in k.items() if 'Visual' not in kk}
            if 'batchPositions' in k:
                ids = [new_body('multibody', dict(pos=list(map(float, pp)), **{kk: vv for kk, vv in info.items() if kk != 'batchPositions'})) for pp in k['batchPositions']]
                R.calls.append(('createMultiBody', (), dict(info, bodies=ids)))
                return ids[-1]
            bid = new_body('multibody', dict(pos=list(map(float, k.get('basePosition', [0, 0, 0]))), **info))
            R.calls.append(('createMultiBody', (), dict(info, bodies=[bid])))
            return bid
        m.createMultiBody = createMultiBody

        def nj(body):
            b = R.bodies[body]
            if b['kind'] == 'urdf':
                return len(b['model']['links']) - 1 if b['model'] else 0
            return len(b.get('linkMasses', []))
        m.getNumJoints = lambda body, physicsClientId=None: nj(body)

        def getJointInfo(body, j, physicsClientId=None):
            b = R.bodies[body]
            if b['kind'] == 'urdf' and b['model']:
                lk = b['model']['links'][j + 1]
                jt = lk['joint']
                t = {'revolute': 0, 'continuous': 0, 'prismatic': 1, 'fixed': 4}[jt['type']]
                lo, hi = (0.0, -1.0) if jt['type'] in ('continuous', 'fixed') else (jt['lower'], jt['upper'])
                return (j, jt['name'].encode(), t, 0, 0, 0, 0.0, 0.0, lo, hi, jt.get('effort', 0.0), jt.get('velocity', 0.0), lk['name'].encode(), tuple(jt.get('axis', [0, 0, 1])),
                        (0, 0, 0), (0, 0, 0, 1), lk['parent'] - 1)
            i = dfs_order(b['linkParentIndices'])[j] - 1 if 'linkParentIndices' in b else j      # PyBullet numbers createMultiBody links depth-first
            lo = b['linkLowerLimits'][i] if 'linkLowerLimits' in b else 0.0
            hi = b['linkUpperLimits'][i] if 'linkUpperLimits' in b and i < len(b['linkUpperLimits']) else -1.0
            return (j, b'joint', int(b['linkJointTypes'][i]) if 'linkJointTypes' in b else 0, 0, 0, 0, 0.0, 0.0, lo, hi, 0.0, 0.0, b'link', (0, 0, 1), (0, 0, 0), (0, 0, 0, 1), -1)
        m.getJointInfo = getJointInfo
        m.getJointState = lambda body, j, physicsClientId=None: (0.0, 0.0, (0.0,) * 6, 0.0)
        m.getJointStates = lambda body, jointIndices=None, physicsClientId=None: [(0.0, 0.0, (0.0,) * 6, 0.0) for _ in jointIndices]
        origin = (np.zeros(3), np.array([0, 0, 0, 1.0]))
        def getLinkState(body, link, **k):
            # the link an IK call was last asked to move is reported AT the goal (the search loops of position_robot_toc then terminate)
            pos, orn = R.ik_goal.get((int(body), int(link)), origin)
            return (pos, orn, origin[0], origin[1], pos, orn, np.zeros(3), np.zeros(3))
        m.getLinkState = getLinkState
        m.getBasePositionAndOrientation = lambda body, physicsClientId=None: R.pose.get(int(body), origin)
        m.getBaseVelocity = lambda body, physicsClientId=None: (np.zeros(3), np.zeros(3))

        def resetBase(body, pos, orn, physicsClientId=None):
            R.pose[int(body)] = (np.array(pos, float), np.array(orn, float))
            R.calls.append(('resetBasePositionAndOrientation', (int(body),), dict(pos=np.array(pos, float).tolist(), orn=np.array(orn, float).tolist())))
        m.resetBasePositionAndOrientation = resetBase
        m.getClosestPoints = lambda **k: []
        m.getContactPoints = lambda **k: []
        m.getDynamicsInfo = lambda body, link, physicsClientId=None: (1.0, 0.5, (1, 1, 1), (0, 0, 0), (0, 0, 0, 1), 0, 0, 0, 0, 0)
        m.getAABB = lambda body, linkIndex=-1, physicsClientId=None: ((0, 0, 0), (0, 0, 0))

        def ik(body, ee, targetPosition=None, targetOrientation=None, **k):
            if R.ik_calls < 4:
                R.calls.append(('calculateInverseKinematics', (int(body), int(ee)), dict(targetPosition=np.asarray(targetPosition, float).tolist(),
                                targetOrientation=None if targetOrientation is None else np.asarray(targetOrientation, float).tolist())))
            R.ik_calls += 1
            R.ik_goal[(int(body), int(ee))] = (np.asarray(targetPosition, float), origin[1] if targetOrientation is None else np.asarray(targetOrientation, float))
            return np.zeros(max(1, sum(1 for j in range(nj(body)) if getJointInfo(body, j)[2] != 4)))
        m.calculateInverseKinematics = ik

        def quat(e, physicsClientId=None):
            r, pt, y = [float(v) for v in e]
            cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(pt / 2), np.sin(pt / 2), np.cos(y / 2), np.sin(y / 2)
            return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
        m.getQuaternionFromEuler = quat
        m.getEulerFromQuaternion = lambda q, physicsClientId=None: (0.0, 0.0, 0.0)

        def inv(position=None, orientation=None, physicsClientId=None):
            qi = np.array([-orientation[0], -orientation[1], -orientation[2], orientation[3]], float)
            return -qrot(qi, position), qi
        m.invertTransform = inv
        m.multiplyTransforms = lambda positionA=None, orientationA=None, positionB=None, orientationB=None, physicsClientId=None: (
            np.asarray(positionA, float) + qrot(orientationA, positionB), qmul(np.asarray(orientationA, float), np.asarray(orientationB, float)))
        def jac(body, link, localPosition=None, objPositions=None, objVelocities=None, objAccelerations=None, physicsClientId=None):
            n = len(objPositions)
            g = np.random.RandomState(0).normal(size=(6, n))
            return g[:3].tolist(), g[3:].tolist()
        m.calculateJacobian = jac
        m.connect = lambda *a, **k: 0
        return m


def record(cls_path, seed=1001):
    rec = Recorder()
    rec.install()
    mod, name = cls_path.rsplit('.', 1)
    for k in [k for k in sys.modules if k.startswith('assistive_gym')]:
        del sys.modules[k]                                     # re-import against this recorder
    import importlib
    env = getattr(importlib.import_module(mod), name)()
    env.seed(seed)
    env.reset()
    names = {i: (b.get('file') or 'multibody') for i, b in enumerate(rec.bodies)}
    calls = [dict(fn=c[0], args=jsonable(list(c[1])), kw=jsonable({k: (np.asarray(v).tolist() if isinstance(v, np.ndarray) else v) for k, v in c[2].items()}))
             for c in rec.calls if c[0] in KEEP]
    human_resets = [dict(joint=int(c[2].get('jointIndex', -1)), value=float(c[2].get('targetValue', 0.0))) for c in rec.calls
                    if c[0] == 'resetJointState' and int(c[1][0]) == getattr(env.human, 'body', -99) and abs(float(c[2].get('targetValue', 0.0))) > 1e-12]
    return dict(bodies=names, calls=calls, human_body=int(env.human.body), robot_body=int(env.robot.body), tool_body=int(getattr(env.tool, 'body', -1) or -1),
                human_gender=env.human.gender, human_joint_resets=human_resets, n_step_simulation=sum(1 for c in rec.calls if c[0] == 'stepSimulation'),
                motor_gains=dict(robot=float(env.robot.motor_gains), human=float(env.human.motor_gains)),
                human_zero_mass_joints=sorted({int(c[1][1]) for c in rec.calls if c[0] == 'changeDynamics' and int(c[1][0]) == int(env.human.body) and c[2].get('mass') == 0}))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    install_stubs(ref)
    out = {}
    for key, path in (('feeding', 'assistive_gym.envs.feeding_envs.FeedingJacoEnv'), ('scratch_itch', 'assistive_gym.envs.scratch_itch_envs.ScratchItchJacoEnv'),
                      ('bed_bathing', 'assistive_gym.envs.bed_bathing_envs.BedBathingSawyerEnv'), ('dressing', 'assistive_gym.envs.dressing_envs.DressingPR2Env'),
                      ('drinking', 'assistive_gym.envs.drinking_envs.DrinkingJacoEnv')):
        try:
            out[key] = record(path)
            print(key, 'ok:', len(out[key]['calls']), 'calls kept,', out[key]['n_step_simulation'], 'stepSimulation')
        except Exception as e:                                 # a reset that needs more of the engine than the recorder offers
            import traceback
            print(key, 'FAILED:', repr(e))
            traceback.print_exc(limit=3)
    json.dump(out, open(os.path.join(HERE, 'reset_recipes.json'), 'w'), separators=(',', ':'))


if __name__ == '__main__':
    main()
