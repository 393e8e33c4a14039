"""Golden rollout of the reference's OWN `FeedingEnv.step` (envs/feeding.py:12-112 with `AssistiveEnv.take_step`, `human_preferences`
and every `Agent` getter it calls, all unmodified) executed in this container on top of the CPU oracle: the reference package is
imported with inert stubs for its third-party imports, and `pybullet` is replaced by a small facade that answers the calls of this
path (getJointStates, getLinkState, getBasePositionAndOrientation, getBaseVelocity, getContactPoints, getClosestPoints,
setJointMotorControlArray, resetBasePositionAndOrientation, stepSimulation, the transform helpers) from an `OracleSim` holding the
FeedingJaco scene.  What is pinned is therefore the SEMANTICS of the step around the physics (action -> targets, observation,
food bookkeeping, reward, done, info), by the reference's code itself; the physics under both is the oracle.

Output: tests/golden/feeding_semantics.npz (the start sample, the actions, the forced events, and the reference's obs / reward /
done / info per step).  tests/test_reference_feeding_semantics.py replays the same rollout with the repo's restatement
(`tests/parity_cases.feeding_semantics_reference`, the function the fused CUDA kernels are checked against).

usage: python tests/golden/make_golden_feeding_semantics.py [/root/reference]"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_STEPS, EAT_STEP, SEED = 40, 12, 11
EAT_V0 = 0.5886


def qmul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def qrot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v, dtype=np.float64)


class Facade:
    """the pybullet calls of the Feeding step path, answered by an OracleSim with one env"""

    def __init__(self, sim, scene, f32_targets=False):
        # f32_targets: motor targets rounded to fp32 before they reach the oracle, as the repo's host mirror (`Agent.control`) hands them over
        self.sim, self.sc, self.f32 = sim, scene, f32_targets

    def gl(self, body, link):
        return int(self.sc['body_link0'][body]) + 1 + int(link)

    def install(self, p):
        sim, sc, gl = self.sim, self.sc, self.gl
        p.POSITION_CONTROL = 2

        def getJointStates(body, jointIndices=None, physicsClientId=None):
            q, qd, tau = sim.get_joint_states([gl(body, j) for j in jointIndices])
            return [(float(q[0, i]), float(qd[0, i]), (0.0,) * 6, float(tau[0, i])) for i in range(len(jointIndices))]
        p.getJointStates = getJointStates

        def getLinkState(body, link, computeForwardKinematics=False, computeLinkVelocity=False, physicsClientId=None):
            s = sim.get_link_states([gl(body, link)])
            return (s['com_pos'][0, 0], s['com_quat'][0, 0], None, None, s['pos'][0, 0], s['quat'][0, 0], s['lin_vel'][0, 0], s['ang_vel'][0, 0])
        p.getLinkState = getLinkState

        def getBasePositionAndOrientation(body, physicsClientId=None):
            s = sim.get_link_states([gl(body, -1)])
            return s['com_pos'][0, 0], s['com_quat'][0, 0]
        p.getBasePositionAndOrientation = getBasePositionAndOrientation

        def getBaseVelocity(body, physicsClientId=None):
            s = sim.get_link_states([gl(body, -1)])
            return s['lin_vel'][0, 0], s['ang_vel'][0, 0]
        p.getBaseVelocity = getBaseVelocity

        def records(c, k, body_a, body_b):
            out = []
            for i in range(min(int(k[0]), c.shape[1])):
                r = c[0, i]
                la, lb = int(r['link_a']) - int(sc['body_link0'][body_a]) - 1, int(r['link_b']) - int(sc['body_link0'][int(sc['link_body'][int(r['link_b'])])]) - 1
                out.append((0, body_a, body_b, la, lb, np.array(r['pos_a'], dtype=np.float64), np.array(r['pos_b'], dtype=np.float64), np.array(r['normal'], dtype=np.float64),
                            float(r['distance']), float(r['normal_force'])))
            return out

        def getContactPoints(bodyA=None, bodyB=None, linkIndexA=None, linkIndexB=None, physicsClientId=None):
            c, k = sim.get_contacts(bodyA, -2 if bodyB is None else bodyB, -2 if linkIndexA is None else gl(bodyA, linkIndexA),
                                    -2 if linkIndexB is None else gl(bodyB, linkIndexB), max_pts=256)
            assert int(k[0]) <= 256
            return records(c, k, bodyA, bodyB)
        p.getContactPoints = getContactPoints

        def getClosestPoints(bodyA=None, bodyB=None, distance=0.0, linkIndexA=None, linkIndexB=None, physicsClientId=None):
            c, k = sim.closest_points(bodyA, bodyB, distance, max_pts=2048)
            assert int(k[0]) <= 2048
            return records(c, k, bodyA, bodyB)
        p.getClosestPoints = getClosestPoints

        def setJointMotorControlArray(body, jointIndices=None, controlMode=None, targetPositions=None, positionGains=None, forces=None, physicsClientId=None, **k):
            links = [gl(body, j) for j in jointIndices]
            tgt = np.asarray(targetPositions, dtype=np.float64)
            if self.f32:
                tgt = tgt.astype(np.float32).astype(np.float64)
            sim.set_motor(links, 1, target=tgt[None], kp=list(np.asarray(positionGains, dtype=np.float64)),
                          kd=[1.0] * len(links), max_force=list(np.asarray(forces, dtype=np.float64)))
        p.setJointMotorControlArray = setJointMotorControlArray

        def resetBasePositionAndOrientation(body, pos, orient, physicsClientId=None):
            sim.set_base_pose(body, np.asarray(pos, dtype=np.float64)[None], np.asarray(orient, dtype=np.float64)[None])
        p.resetBasePositionAndOrientation = resetBasePositionAndOrientation
        p.stepSimulation = lambda physicsClientId=None: sim.step(1)

        def invertTransform(pos, orient, physicsClientId=None):
            qi = np.array([-orient[0], -orient[1], -orient[2], orient[3]], dtype=np.float64)
            return -qrot(qi, pos), qi
        p.invertTransform = invertTransform
        p.multiplyTransforms = lambda pa, qa, pb, qb, physicsClientId=None: (np.asarray(pa, dtype=np.float64) + qrot(qa, pb), qmul(np.asarray(qa, dtype=np.float64), np.asarray(qb, dtype=np.float64)))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    from make_golden_env_logic import install_stubs
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.agents.agent import Agent
    from assistive_gym.envs.feeding_envs import FeedingJacoEnv
    from assistive_gym_b200 import capi
    from assistive_gym_b200.feeding_batch import FeedingBatch
    from oracle.oracle_py import OracleSim
    fb = FeedingBatch()
    sim = OracleSim(fb.scene, capi.default_config(), 1)
    rng = np.random.default_rng(SEED)
    smp = fb.reset(sim, rng, settle_steps=25, impairment='none')
    male = bool(smp['male'][0])
    # ---- the reference env, wired to the oracle through the facade
    env = FeedingJacoEnv()
    Facade(sim, fb.scene).install(sys.modules['pybullet'])
    env.robot.body, env.tool.body = fb.robot, fb.tool
    env.human.body = fb.humans['male' if male else 'female']
    env.human.gender = 'male' if male else 'female'
    for a in (env.robot, env.tool, env.human):
        a.id = 0
    env.robot.controllable_joint_lower_limits = np.array(fb.arm_lower, dtype=np.float64)
    env.robot.controllable_joint_upper_limits = np.array(fb.arm_upper, dtype=np.float64)
    env.robot.motor_gains = env.human.motor_gains = 0.025               # feeding.py:122
    env.agents = [env.robot]
    env.foods = []
    for f in fb.foods:
        a = Agent(); a.body, a.id = f, 0
        env.foods.append(a)
    env.foods_active = list(env.foods)
    env.total_food_count = len(env.foods)
    env.mouth_pos = [0, -0.11, 0.03] if male else [0, -0.1, 0.03]       # feeding.py:186
    env.target = types.SimpleNamespace(set_base_pos_orient=lambda *a, **k: None)
    env.iteration, env.task_success, env.last_sim_time, env.gui = 0, 0, None, False
    env.action_space = types.SimpleNamespace(low=-np.ones(7), high=np.ones(7))
    env.np_random = np.random.RandomState(0)
    env.update_targets()
    start_state = sim.state_get()                                       # stored: the replay does not depend on the IK of the reset being bit-reproducible
    arng = np.random.default_rng(SEED + 1)
    actions = arng.uniform(-1, 1, size=(N_STEPS, 7)) * 0.3              # gentle, so that the food stays on the spoon for a while
    obs, rew, done, total, success, n_foods, n_active = [], [], [], [], [], [], []
    for t in range(N_STEPS):
        if t == EAT_STEP:                                               # forced event: a food particle is tossed up from the mouth target so that
            eat_food = fb.foods.index(env.foods[0].body)                # (one that is still on the spoon)
            sim.set_base_pose(fb.foods[eat_food], env.target_pos[None], np.array([[0, 0, 0, 1.0]]))      # it is back there after the 5 substeps (0.1 v0 = 15 g dt^2)
            sim.set_base_velocity(fb.foods[eat_food], np.array([[0, 0, EAT_V0]]), np.zeros((1, 3)))
        o, r, d, info = env.step(actions[t].copy())
        obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r)); done.append(bool(d)); total.append(float(info['total_force_on_human']))
        success.append(int(env.task_success)); n_foods.append(len(env.foods)); n_active.append(len(env.foods_active))
    out = {('sample_' + k): np.asarray(v) for k, v in smp.items()}
    out.update(actions=actions, obs=np.array(obs), reward=np.array(rew), done=np.array(done), total_force=np.array(total), task_success=np.array(success),
               n_foods=np.array(n_foods), n_foods_active=np.array(n_active), eat_step=np.array(EAT_STEP), eat_v0=np.array(EAT_V0), eat_food=np.array(eat_food), seed=np.array(SEED))
    out['start_state'] = start_state
    np.savez_compressed(os.path.join(HERE, 'feeding_semantics.npz'), **out)
    print('steps', N_STEPS, 'foods left', n_foods[-1], 'active', n_active[-1], 'task_success', success[-1], 'reward range', min(rew), max(rew), 'max force', max(total))


if __name__ == '__main__':
    main()
