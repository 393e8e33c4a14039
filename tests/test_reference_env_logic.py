"""The repo's mirrors of the reference's env-level logic against vectors recorded from the reference's own methods
(tests/golden/env_logic.json, written by tests/golden/make_golden_env_logic.py, which imports the reference package with every
third-party module stubbed): the action -> motor-target rule of `take_step`, `human_preferences`, the reward weights of
config.ini, and the robot classes' constant tables."""
import json
import os

import numpy as np
import pytest

from assistive_gym_b200 import envs
from assistive_gym_b200.envs.agents.robot import PR2, Jaco, Sawyer
from tests.parity_cases import take_step_targets

_ALL = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'env_logic.json')))
G = _ALL['tasks']
OURS = {'FeedingJacoEnv': ('FeedingJaco-v1', Jaco, 'right'), 'BedBathingSawyerEnv': ('BedBathingSawyer-v1', Sawyer, 'left'),
        'DressingPR2Env': ('DressingPR2-v1', PR2, 'left'), 'ScratchItchJacoEnv': ('ScratchItchJaco-v1', Jaco, 'left')}


@pytest.mark.parametrize('name', sorted(G))
def test_take_step_targets_are_the_reference_s(name):
    """(a) the numpy restatement that drives oracle and product in the parity tests, (b) `AssistiveEnv.take_step` of the host mirror"""
    g = G[name]
    ts = g['take_step']
    lo, hi = np.array(ts['lower']), np.array(ts['upper'])
    am = g['robot'].get('action_multiplier', 1)
    env = envs.make(OURS[name][0], n_envs=1)
    robot = env.robot
    robot.controllable_joint_lower_limits, robot.controllable_joint_upper_limits = lo, hi
    env.agents = [robot]
    clamped = 0
    for s in ts['steps']:
        q, a, want = np.array(s['q']), np.array(s['action']), np.array(s['targets'])
        got = take_step_targets(q[None], a[None], lo, hi, mult=0.05 * am)[0]
        assert np.allclose(got, want, atol=1e-12)
        clamped += int(np.any((want == lo) | (want == hi)))
        rec = {}
        robot.get_joint_angles = lambda idx, q=q: np.array(q)
        robot.control = lambda idx, tgt, gains, forces, rec=rec: rec.update(idx=list(idx), tgt=np.array(tgt), gains=gains, forces=forces)
        env.take_step(a[None], step_sim=False)
        assert rec['idx'] == s['indices'] and np.allclose(np.ravel(rec['tgt']), want, atol=1e-12)
        assert rec['gains'] == s['gains'] and rec['forces'] == s['forces']
    assert clamped > 10                                                  # the limit clamp was exercised


@pytest.mark.parametrize('name', sorted(G))
def test_human_preferences_and_weights_are_the_reference_s(name):
    g = G[name]
    env = envs.make(OURS[name][0], n_envs=1)
    assert env.task == g['task'] and env.obs_robot_len == g['obs_robot_len'] and env.action_robot_len == g['action_robot_len']
    for k, v in g['weights'].items():
        assert abs(env.config(k) - v) < 1e-12, k
    for k, v in g['preference_weights'].items():
        assert abs(getattr(env, k) - v) < 1e-12, k
    for p in g['human_preferences']:
        kw = {k: v for k, v in p.items() if k != 'out'}
        assert abs(float(env.human_preferences(**kw)) - p['out']) < 1e-9


@pytest.mark.parametrize('name', sorted(G))
def test_robot_constant_tables_are_the_reference_s(name):
    g = G[name]['robot']
    _, cls, arm = OURS[name]
    r = cls(arm)
    for k, v in g.items():
        ours = getattr(r, k)
        if isinstance(v, dict):
            for task, val in v.items():
                if task in ours:
                    assert np.allclose(np.asarray(ours[task], dtype=np.float64).ravel(), np.asarray(val, dtype=np.float64).ravel(), atol=1e-12), (k, task)
            task = G[name]['task']
            assert task not in v or task in ours, (k, task)              # the entry of the task this env is built for must be there
        else:
            assert np.allclose(np.asarray(ours, dtype=np.float64), np.asarray(v, dtype=np.float64), atol=1e-12), k


def test_base_pose_ranking_score_is_the_reference_s():
    """joint-limit weighting and the JLWKI score of `position_robot_toc` (agents/robot.py:173-186, :223-235) as `toc.py` restates them"""
    from assistive_gym_b200.toc import jlwki, joint_limited_weighting
    floor = 0
    for c in _ALL['jlwki']:
        q, lo, hi, J = (np.array(c[k]) for k in ('q', 'lower', 'upper', 'J'))
        assert np.allclose(joint_limited_weighting(q, lo, hi), c['weights'], rtol=1e-12, atol=1e-15)
        assert abs(jlwki(J[None], q[None], lo, hi)[0] - c['jlwki']) < 1e-12 * (1 + c['jlwki'])
        floor += int(np.any(np.array(c['weights']) == 0.001))
    assert floor >= 1                                                     # the 0.001 floor (angles at / beyond a limit) is exercised
