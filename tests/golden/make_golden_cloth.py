#!/usr/bin/env python3
"""Generates tests/golden/cloth_gown_1step.npz from the cloth ORACLE (not from Bullet: parity unpinned): the reference's gown
over the synthetic obstacle scene of tests/cloth_cases.py, state before and after one stepSimulation of 8 substeps with a moving
arm capsule.  Committed together with its output; the product (host-compiled kernel bodies, CUDA) and the oracle are both
checked against the stored arrays."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from assistive_gym_b200 import capi  # noqa: E402
from assistive_gym_b200.cloth import ClothModel  # noqa: E402
from oracle.oracle_py import OracleSim  # noqa: E402
from tests import cloth_cases as cc  # noqa: E402

ANCHORS = [2086, 2087, 2088, 2041]


def start_state(model, n=1):
    x0 = np.repeat(model.rest[None], n, axis=0)
    x0 = x0 - x0.mean(axis=1, keepdims=True) + np.array([0.2, 0.15, 0.36])
    return x0


def setup(sim, model, x0, v0, arm_joint, links, static):
    sim.cloth_init(model, links, static, ANCHORS, model.rest[ANCHORS] - model.rest[ANCHORS[0]], max_contacts=2048)
    sim.cloth_set_state(x0, v0)
    sim.cloth_set_anchor(x0[:, ANCHORS[0]].copy())
    sim.set_joint_state([arm_joint], q=np.full((len(x0), 1), -0.8), qd=np.full((len(x0), 1), 2.0))
    sim.forward_kinematics()


if __name__ == '__main__':
    model = ClothModel.load()
    scene, links, static, arm_joint = cc.obstacle_scene()
    orc = OracleSim(scene, capi.default_config(num_substeps=8), 1)
    x0 = start_state(model)
    setup(orc, model, x0, np.zeros_like(x0), arm_joint, links, static)
    orc.step(2)                                   # let contacts form
    xa, va = orc.cloth_get_state()
    rigid_a = orc.state_get()
    orc.step(1)
    xb, vb = orc.cloth_get_state()
    cnt, node, pos, force, link = orc.cloth_get_contacts(4096)
    out = dict(x_before=xa.astype(np.float32), v_before=va.astype(np.float32), rigid_before=rigid_a, x_after=xb.astype(np.float32), v_after=vb.astype(np.float32),
               contact_count=cnt, contact_node=node[:, :cnt.max()].astype(np.int16), contact_link=link[:, :cnt.max()].astype(np.int16))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cloth_gown_1step.npz'), **out)
    print('wrote fixture', {k: getattr(v, 'shape', v) for k, v in out.items()}, 'contacts', cnt)
