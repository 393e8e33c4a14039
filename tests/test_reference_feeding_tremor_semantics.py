"""Feeding step with a `tremor` person (reference envs/env.py:130-131, :212-215, :226-229): the repo's restatement
(`tests/parity_cases.apply_tremor` + `take_step_targets` + `feeding_semantics_reference`, what the fused kernels -- `feeding_pre_body`'s
tremor targets included -- are checked against) replays the rollout of tests/golden/feeding_tremor_semantics.npz, produced by the
reference's OWN step code on the CPU oracle through a pybullet facade (tests/golden/make_golden_feeding_tremor_semantics.py).
Pins the sign convention (the parity of the ALREADY incremented step counter) and the per-substep limit clamp of the head joints."""
import os

import numpy as np

from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from oracle.oracle_py import OracleSim
from tests.parity_cases import apply_tremor, feeding_semantics_reference, head_q, take_step_targets

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'feeding_tremor_semantics.npz'))


def test_restated_tremor_step_reproduces_the_reference_s_rollout():
    fb = FeedingBatch()
    sim = OracleSim(fb.scene, capi.default_config(), 1)
    smp = fb.reset(sim, np.random.default_rng(int(G['seed'])), settle_steps=25, impairment='tremor')                 # the generator's call
    assert all(np.array_equal(np.asarray(smp[k]), G['sample_' + k]) for k in smp if 'sample_' + k in G.files) and smp['impairment'][0] == 3
    sim.state_set(G['start_state']); sim.forward_kinematics()               # exactly the generator's start state
    state = dict(male=smp['male'], foods=np.ones((1, 8), dtype=bool), active=np.ones((1, 8), dtype=bool), iteration=np.zeros(1, dtype=int), task_success=np.zeros(1, dtype=int))
    for t, a in enumerate(G['actions']):
        apply_tremor(fb, [sim], smp, t + 1)                                # the counter is incremented before the targets are set (env.py:185)
        q = sim.get_joint_states(fb.arm_links)[0]
        sim.set_motor_targets(fb.arm_links, take_step_targets(q, a[None], fb.arm_lower, fb.arm_upper))
        sim.step(5)
        obs, rew, done, total = feeding_semantics_reference(fb, sim, a[None], state)
        assert np.allclose(head_q(fb, sim, smp)[0], G['head_q'][t], rtol=0, atol=1e-9), (t, np.abs(head_q(fb, sim, smp)[0] - G['head_q'][t]).max())
        assert np.allclose(obs[0], G['obs'][t], rtol=0, atol=1e-9) and abs(rew[0] - G['reward'][t]) < 1e-9
    d = G['head_q'][:, 1] - fb.tremor_rest_of(smp)[0, 1]
    assert np.all(np.sign(d[1:]) == -np.sign(d[:-1])) and np.abs(d).max() > 1e-3      # the head really shakes: the offset flips every step
