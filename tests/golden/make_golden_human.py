"""Golden description of the capsule person, recorded from the REFERENCE's own code (this container only).

`assistive_gym/envs/human_creation.py::HumanCreation.create_human` is executed against a stub `pybullet` module that records the
arguments of `createCollisionShape`, `createMultiBody` and `setCollisionFilterPair` instead of building anything.  The result --
what the reference asks the physics engine to create -- is written to tests/golden/human_<gender>.json; tests/test_human_model.py
compares `assistive_gym_b200/human_model.py` (a restatement written from reading that file) against it.

usage: python tests/golden/make_golden_human.py [/root/reference]"""
import importlib.util
import json
import os
import sys
import types

import numpy as np


def make_stub(rec):
    p = types.ModuleType('pybullet')
    p.GEOM_SPHERE, p.GEOM_BOX, p.GEOM_CYLINDER, p.GEOM_MESH, p.GEOM_PLANE, p.GEOM_CAPSULE = 2, 3, 4, 5, 6, 7
    p.JOINT_REVOLUTE, p.JOINT_PRISMATIC, p.JOINT_FIXED = 0, 1, 4
    p.URDF_USE_SELF_COLLISION = 8

    def quat(e, physicsClientId=None):
        r, pt, y = [float(v) for v in e]
        cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(pt / 2), np.sin(pt / 2), np.cos(y / 2), np.sin(y / 2)
        return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
    p.getQuaternionFromEuler = quat

    def create_visual(*a, **k):
        return -2
    p.createVisualShape = create_visual

    def create_collision(shape=None, shapeType=None, radius=0.5, height=1.0, fileName=None, meshScale=None, collisionFramePosition=(0, 0, 0),
                         collisionFrameOrientation=(0, 0, 0, 1), physicsClientId=None, **k):
        shape = shapeType if shape is None else shape
        rec['shapes'].append(dict(kind={2: 'sphere', 7: 'capsule', 5: 'mesh'}[shape], radius=float(radius), height=float(height),
                                  file=os.path.basename(fileName) if fileName else None, mesh_scale=[float(v) for v in meshScale] if meshScale is not None else None,
                                  frame_pos=[float(v) for v in collisionFramePosition], frame_quat=[float(v) for v in collisionFrameOrientation]))
        return len(rec['shapes']) - 1
    p.createCollisionShape = create_collision

    def create_multibody(**k):
        rec['multibody'] = {n: (np.asarray(v, dtype=np.float64).tolist() if n not in ('linkCollisionShapeIndices', 'linkParentIndices', 'linkJointTypes', 'baseCollisionShapeIndex') else
                                (np.asarray(v).astype(int).tolist())) for n, v in k.items() if n not in ('linkVisualShapeIndices', 'baseVisualShapeIndex', 'physicsClientId')}
        rec['n_links'] = len(k['linkMasses'])
        return 0
    p.createMultiBody = create_multibody
    p.getNumJoints = lambda body, physicsClientId=None: rec['n_links']

    def filt(a, b, la, lb, on, physicsClientId=None):
        key = (min(la, lb), max(la, lb))
        if on:
            rec['pairs'].add(key)
        else:
            rec['pairs'].discard(key)
    p.setCollisionFilterPair = filt
    p.getJointStates = lambda body, jointIndices=None, physicsClientId=None: [(0.0, 0.0, (0,) * 6, 0.0) for _ in jointIndices]

    def joint_info(body, j, physicsClientId=None):
        mb = rec['multibody']
        return (j, b'joint%d' % j, mb['linkJointTypes'][j], 0, 0, 0, 0.0, 0.0, mb['linkLowerLimits'][j], mb['linkUpperLimits'][j])
    p.getJointInfo = joint_info
    p.resetJointState = lambda *a, **k: rec['clamped'].append(int(k.get('jointIndex', a[1] if len(a) > 1 else -1)))
    return p


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for gender in ('male', 'female'):
        rec = dict(shapes=[], pairs=set(), clamped=[])
        sys.modules['pybullet'] = make_stub(rec)
        spec = importlib.util.spec_from_file_location('ref_human_creation', os.path.join(ref, 'assistive_gym', 'envs', 'human_creation.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        hc = mod.HumanCreation(pid=0, np_random=np.random.RandomState(0))
        hc.create_human(static=True, limit_scale=1.0, skin_color=[0.8, 0.6, 0.4, 1], gender=gender)
        del sys.modules['pybullet']
        rec['pairs'] = sorted([list(k) for k in rec['pairs']])
        rec['radii'] = dict(hand=hc.hand_radius, elbow=hc.elbow_radius, shoulder=hc.shoulder_radius)
        json.dump(rec, open(os.path.join(out_dir, 'human_%s.json' % gender), 'w'), separators=(',', ':'))
        print(gender, 'links', rec['n_links'], 'shapes', len(rec['shapes']), 'self-collision pairs', len(rec['pairs']), 'clamped at creation', rec['clamped'])


if __name__ == '__main__':
    main()
