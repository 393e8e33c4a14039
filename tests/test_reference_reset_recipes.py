"""The batched scene builders against what the reference's OWN `reset()` asks the physics engine to build
(tests/golden/reset_recipes.json, recorded by tests/golden/make_golden_reset_recipes.py: the reference package run against a
recording pybullet).  Link states are the origin in the recording, so tool and food placements appear there as the offsets the
reference composes; randomised quantities (bowl offset, IK target offset, head angles, plane friction) are checked as
"nominal + range"."""
import inspect
import json
import os

import numpy as np

from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import HUMAN_PRESET as FEED_PRESET
from assistive_gym_b200.feeding_batch import JACO as FEED_JACO
from assistive_gym_b200.feeding_batch import TREMOR_JOINTS, FeedingBatch
from assistive_gym_b200.kinematics import q_from_rpy
from assistive_gym_b200.scratch_itch_batch import HUMAN_PRESET as SCRATCH_PRESET
from assistive_gym_b200.scratch_itch_batch import JACO as SCRATCH_JACO
from assistive_gym_b200.scratch_itch_batch import RIGHT_ARM_JOINTS, ScratchItchBatch
from oracle.oracle_py import OracleSim

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reset_recipes.json')))


def _by(calls, fn):
    return [c for c in calls if c['fn'] == fn]


def _same_rotation(qa, qb):
    return min(np.abs(np.asarray(qa) - np.asarray(qb)).max(), np.abs(np.asarray(qa) + np.asarray(qb)).max()) < 1e-9


def _common(g, batch, jaco, preset, task_gripper, ik_nominal):
    sc = batch.scene
    loads = {c['args'][0]: c['kw'] for c in _by(g['calls'], 'loadURDF')}
    last_pose = {}
    for c in _by(g['calls'], 'resetBasePositionAndOrientation'):
        last_pose[c['args'][0]] = c['kw']
    gender = g['human_gender']
    # bodies: where the reference puts them
    assert np.allclose(loads['wheelchair_jaco.urdf']['basePosition'], sc['base_pos0'][batch.wheelchair])
    assert np.allclose(last_pose[g['robot_body']]['pos'], batch.robot_base_pos) and _same_rotation(last_pose[g['robot_body']]['orn'], batch.robot_base_quat)
    assert np.allclose(last_pose[g['human_body']]['pos'], batch.builder.bodies[batch.humans[gender]].base_pos)
    # the tool rides on a fixed constraint at link 8 with the task's offsets, 500 N
    con = _by(g['calls'], 'createConstraint')[0]
    assert con['args'][:5] == [g['robot_body'], jaco['tool_joint'], g['tool_body'], -1, 4]
    assert np.allclose(con['kw']['parentFramePosition'], batch.tool_pos_offset) and _same_rotation(con['kw']['parentFrameOrientation'], batch.tool_quat_offset)
    assert _by(g['calls'], 'changeConstraint')[0]['kw']['maxForce'] == 500 == float(sc['con_maxforce'][0])
    cl = np.asarray(sc['con_link']).ravel()
    assert int(cl[0]) == batch.gl(batch.robot, jaco['tool_joint']) and int(cl[1]) == int(sc['body_link0'][batch.tool])
    # gravity: off for robot, person and tool, on for everything else
    off = {c['kw']['body'] for c in _by(g['calls'], 'setGravity') if 'body' in c['kw'] and c['args'] == [0, 0, 0]}
    assert off == {g['robot_body'], g['human_body'], g['tool_body']}
    assert _by(g['calls'], 'setGravity')[0]['args'] == [0, 0, -9.81]
    gz = np.asarray(sc['body_gravity'])[:, 2]
    for b in (batch.robot, batch.tool, batch.humans['male'], batch.humans['female']):
        assert gz[b] == 0.0
    assert gz[batch.wheelchair] == -9.81 and gz[batch.plane] == -9.81
    # the person's joint presets; joints whose range excludes 0 start at the nearer limit (human_creation.py:301-314)
    resets, seen = {}, {}
    for r in g['human_joint_resets']:                                        # (the recorder's joint states are always 0, so later limit checks write the limit again)
        resets.setdefault(r['joint'], r['value'])
        seen.setdefault(r['joint'], []).append(r['value'])
    for j, deg in preset.items():
        assert any(abs(v - np.deg2rad(deg)) < 1e-12 for v in seen[j]), (j, seen[j])
    hb = batch.humans[gender]
    for j in (3, 13):
        lo, hi = float(sc['link_lower'][batch.gl(hb, j)]), float(sc['link_upper'][batch.gl(hb, j)])
        want = float(np.clip(0.0, lo, hi))                                       # the creation-time clamp of the zero pose
        assert abs(resets[j] - want) < 1e-9, (j, resets[j], want)
    # gripper opened to the task's position with gain 0.05 / 500 N (robot.py:76-79)
    grip = [c for c in _by(g['calls'], 'setJointMotorControlArray') if c['args'][0] == g['robot_body']][0]['kw']
    assert grip['jointIndices'] == jaco['gripper'] and np.allclose(grip['targetPositions'], task_gripper) and np.allclose(grip['positionGains'], 0.05) and np.allclose(grip['forces'], 500)
    assert np.allclose(jaco['gripper_pos'], task_gripper)
    # IK goal of the start pose: nominal + U(-0.05, 0.05)^3, the task's end-effector orientation
    ik = _by(g['calls'], 'calculateInverseKinematics')[0]
    assert ik['args'] == [g['robot_body'], jaco['ee']] and np.all(np.abs(np.array(ik['kw']['targetPosition']) - ik_nominal) <= 0.05 + 1e-12)
    assert _same_rotation(ik['kw']['targetOrientation'], q_from_rpy(jaco['ee_orient_rpy']))
    fr = [c['kw']['lateralFriction'] for c in _by(g['calls'], 'changeDynamics') if c['args'] == [0, -1] and 'lateralFriction' in c['kw']][0]
    assert 0.025 <= fr <= 0.5
    smp = batch.sample(256, np.random.default_rng(0))
    assert smp['plane_friction'].min() >= 0.025 and smp['plane_friction'].max() <= 0.5 and np.abs(smp['ee_offset']).max() <= 0.05
    return loads, last_pose, resets


def test_feeding_scene_recipe_is_the_reference_s():
    g = G['feeding']
    fb = FeedingBatch()
    sc = fb.scene
    loads, last_pose, resets = _common(g, fb, FEED_JACO, FEED_PRESET, [1.33] * 3, np.array([-0.15, -0.65, 1.15]))
    assert np.allclose(loads['table_tall.urdf']['basePosition'], sc['base_pos0'][fb.table])
    bowl = np.array(loads['bowl.urdf']['basePosition'])
    assert np.all(np.abs(bowl[:2] - [-0.15, -0.65]) <= 0.05) and bowl[2] == 0.75
    assert np.abs(fb.sample(256, np.random.default_rng(1))['bowl_offset'][:, :2]).max() <= 0.05
    for j in (21, 22, 23):                                                   # the head: U(-30, 30) degrees (feeding.py:125)
        assert abs(resets[j]) <= np.deg2rad(30)
    assert g['motor_gains'] == {'robot': 0.025, 'human': 0.025}              # feeding.py:122
    assert g['n_step_simulation'] == 25 == inspect.signature(fb.reset).parameters['settle_steps'].default
    # every joint of a person without tremor is made static (mass 0); the template keeps only the head chain's mass for the tremor envs
    assert g['human_zero_mass_joints'] == list(range(42))
    hb = fb.humans[g['human_gender']]
    assert all(float(sc['link_mass'][fb.gl(hb, j)]) == 0.0 for j in range(42) if j not in TREMOR_JOINTS)
    # the spoon: the reference's mesh at scale 0.08, 1 kg
    spoon_shape = [c['kw'] for c in _by(g['calls'], 'createCollisionShape') if c['kw'].get('fileName') == 'spoon_vhacd.obj'][0]
    assert spoon_shape['meshScale'] == [0.08] * 3
    mb = {c['kw']['bodies'][0]: c['kw'] for c in _by(g['calls'], 'createMultiBody')}
    assert mb[g['tool_body']]['baseMass'] == 1 == float(sc['link_mass'][int(sc['body_link0'][fb.tool])])
    # the food: 8 spheres of radius 5 mm and 1 g on a 2 x 2 x 2 grid above the spoon
    food = [c['kw'] for c in _by(g['calls'], 'createMultiBody') if 'batchPositions' in c['kw']][0]
    assert len(food['bodies']) == 8 == len(fb.foods) and food['baseMass'] == 0.001
    food_shape = _by(g['calls'], 'createCollisionShape')[food['baseCollisionShapeIndex']]['kw']
    assert food_shape['shapeType'] == 2 and food_shape['radius'] == 0.005
    offsets = np.array(food['batchPositions']) - np.array(last_pose[g['tool_body']]['pos'])
    sim = OracleSim(sc, capi.default_config(), 1)
    fb.reset(sim, np.random.default_rng(0), settle_steps=0, impairment='none')
    ls = sim.get_link_states([int(sc['body_link0'][fb.tool])] + [int(sc['body_link0'][f]) for f in fb.foods])
    ours = ls['pos'][0, 1:] - ls['com_pos'][0, 0]
    assert np.allclose(ours, offsets, atol=1e-9)
    for f in fb.foods:
        k = int(sc['body_link0'][f])
        c = int(np.where(np.asarray(sc['col_link']) == k)[0][0])
        assert abs(float(sc['link_mass'][k]) - 0.001) < 1e-15 and abs(float(sc['col_radius'][c]) - 0.005) < 1e-15


def test_scratch_itch_scene_recipe_is_the_reference_s():
    g = G['scratch_itch']
    sb = ScratchItchBatch()
    sc = sb.scene
    loads, last_pose, resets = _common(g, sb, SCRATCH_JACO, SCRATCH_PRESET, [1.0] * 3, np.array([-0.6, 0.0, 0.8]))
    assert g['motor_gains'] == {'robot': 0.05, 'human': 0.05} and g['n_step_simulation'] == 0
    # the scratcher is loaded at the gripper with the task's offsets (tool.py:20-25)
    assert np.allclose(loads['tool_scratch.urdf']['basePosition'], sb.tool_pos_offset) and _same_rotation(loads['tool_scratch.urdf']['baseOrientation'], sb.tool_quat_offset)
    # the person's right arm stays dynamic and is held by position motors of gain 0.01 and force 1 x strength (human.py:123-127)
    assert g['human_zero_mass_joints'] == [j for j in range(42) if j not in RIGHT_ARM_JOINTS]
    hold = [c for c in _by(g['calls'], 'setJointMotorControlArray') if c['args'][0] == g['human_body']][0]['kw']
    assert hold['jointIndices'] == RIGHT_ARM_JOINTS and np.allclose(hold['positionGains'], 0.01)
    assert 0.25 <= hold['forces'][0] <= 1.0 and len(set(hold['forces'])) == 1
    hb = sb.humans[g['human_gender']]
    assert all(float(sc['link_mass'][sb.gl(hb, j)]) == 0.0 for j in range(42) if j not in RIGHT_ARM_JOINTS)
    assert all(float(sc['link_mass'][sb.gl(hb, j)]) > 0.0 or j in (3, 4, 0, 1, 6, 8) for j in RIGHT_ARM_JOINTS)     # (massless helper links of the 3-axis joints)


def test_bed_bathing_scene_recipe_is_the_reference_s():
    from assistive_gym_b200.bed_bathing_batch import SAWYER, BedBathingBatch
    g = G['bed_bathing']
    bb = BedBathingBatch()
    sc = bb.scene
    loads = {c['args'][0]: c['kw'] for c in _by(g['calls'], 'loadURDF')}
    assert np.allclose(loads['bed.urdf']['basePosition'], sc['base_pos0'][bb.bed]) and loads['sawyer.urdf']['useFixedBase'] == 1
    assert [c['kw'] for c in _by(g['calls'], 'changeDynamics') if c['args'] == [g['bodies'] and 3, -1] and 'lateralFriction' in c['kw']][0]['lateralFriction'] == 5
    # the person is laid on the bed: base pose (-0.15, 0.2, 0.95) turned by -90 degrees about x, every joint perturbed by U(-0.1, 0.1), right shoulder 30 degrees,
    # dropped for 100 steps under (0, 0, -1), then made static (bed_bathing.py:118-137)
    lying = [c['kw'] for c in _by(g['calls'], 'resetBasePositionAndOrientation') if c['args'] == [g['human_body']]][-1]
    assert np.allclose(lying['pos'], [-0.15, 0.2, 0.95]) and _same_rotation(lying['orn'], q_from_rpy([-np.pi / 2.0, 0, 0]))
    seen = {}
    for r in g['human_joint_resets']:
        seen.setdefault(r['joint'], []).append(r['value'])
    assert any(abs(v - np.deg2rad(30)) < 1e-12 for v in seen[3])
    assert all(abs(v[0]) <= 0.1 for j, v in seen.items() if j not in (3, 13))
    smp = bb.sample(64, np.random.default_rng(0))
    assert np.abs(smp['joint_noise']).max() <= 0.1
    assert g['n_step_simulation'] == 100 and [0, 0, -1] in [c['args'] for c in _by(g['calls'], 'setGravity') if not c['kw']]
    assert set(range(42)) <= set(g['human_zero_mass_joints'])
    for hb in bb.humans.values():
        assert all(float(sc['link_mass'][bb.gl(hb, j)]) == 0.0 for j in range(42))
    # gravity after the drop: robot and wiper 0, the person (0, 0, -1) (bed_bathing.py:157-163)
    per_body = {c['kw']['body']: c['args'] for c in _by(g['calls'], 'setGravity') if 'body' in c['kw']}
    assert per_body == {g['robot_body']: [0, 0, 0], g['human_body']: [0, 0, -1], g['tool_body']: [0, 0, 0]}
    gv = np.asarray(sc['body_gravity'])
    assert np.allclose(gv[bb.robot], 0) and np.allclose(gv[bb.tool], 0) and all(np.allclose(gv[hb], [0, 0, -1]) for hb in bb.humans.values())
    # the wiper on a fixed constraint at link 18 with the task's offsets
    con = _by(g['calls'], 'createConstraint')[0]
    assert con['args'][:5] == [g['robot_body'], SAWYER['tool_joint'], g['tool_body'], -1, 4]
    assert np.allclose(con['kw']['parentFramePosition'], bb.tool_pos_offset) and _same_rotation(con['kw']['parentFrameOrientation'], bb.tool_quat_offset)
    assert _by(g['calls'], 'changeConstraint')[0]['kw']['maxForce'] == 500 == float(sc['con_maxforce'][0])
    grip = [c for c in _by(g['calls'], 'setJointMotorControlArray') if c['args'][0] == g['robot_body']][0]['kw']
    assert grip['jointIndices'] == SAWYER['gripper'] and np.allclose(grip['targetPositions'], SAWYER['gripper_pos']) and np.allclose(grip['forces'], 500)
    # base-pose search: (-0.85, -0.4, 0) + the task's offset + (U(-0.5, 0), U(-0.5, 0.5), 0), yaw U(-30, 30) degrees (robot.py:142-144)
    base0 = np.array([-0.85, -0.4, 0]) + np.array(SAWYER['toc_base_pos_offset'])
    poses = [c['kw'] for c in _by(g['calls'], 'resetBasePositionAndOrientation') if c['args'] == [g['robot_body']]]
    d = np.array([p['pos'] for p in poses]) - base0
    assert len(poses) >= 50 and d[:, 0].min() >= -0.5 and d[:, 0].max() <= 0 and np.abs(d[:, 1]).max() <= 0.5 and np.allclose(d[:, 2], 0)
    yaw = np.array([2 * np.arctan2(p['orn'][2], p['orn'][3]) for p in poses])
    assert np.abs(yaw).max() <= np.deg2rad(30) + 1e-9 and np.allclose([p['orn'][:2] for p in poses], 0)
    ik = _by(g['calls'], 'calculateInverseKinematics')[0]
    assert ik['args'] == [g['robot_body'], SAWYER['ee']] and np.all(np.abs(np.array(ik['kw']['targetPosition']) - [-0.6, 0.2, 1.0]) <= 0.05 + 1e-12)
    assert _same_rotation(ik['kw']['targetOrientation'], q_from_rpy(SAWYER['ee_orient_rpy']))
    assert g['motor_gains'] == {'robot': 0.05, 'human': 0.05}


def test_dressing_scene_recipe_is_the_reference_s():
    from assistive_gym_b200.cloth import DRESSING_PARAMS
    from assistive_gym_b200.dressing_batch import CLOTH_ANCHORS, CLOTH_ORIG_POS, CLOTH_POSITION, CLOTH_SCALE, LEFT_ARM_JOINTS, PR2, DressingBatch
    from assistive_gym_b200.dressing_batch import HUMAN_PRESET as DRESS_PRESET
    g = G['dressing']
    db = DressingBatch()
    sc = db.scene
    loads = {c['args'][0]: c['kw'] for c in _by(g['calls'], 'loadURDF')}
    assert np.allclose(loads['wheelchair.urdf']['basePosition'], sc['base_pos0'][db.wheelchair]) and loads['pr2_no_torso_lift_tall.urdf']['useFixedBase'] == 1
    seen = {}
    for r in g['human_joint_resets']:
        seen.setdefault(r['joint'], []).append(r['value'])
    for j, deg in DRESS_PRESET.items():
        assert any(abs(v - np.deg2rad(deg)) < 1e-12 for v in seen[j]), (j, seen[j])
    # the left arm is held by position motors of gain 0.01 and force 1 x strength; everything else of the person is static
    hold = [c for c in _by(g['calls'], 'setJointMotorControlArray') if c['args'][0] == g['human_body']][0]['kw']
    assert hold['jointIndices'] == LEFT_ARM_JOINTS and np.allclose(hold['positionGains'], 0.01) and 0.25 <= hold['forces'][0] <= 1.0
    assert g['human_zero_mass_joints'] == [j for j in range(42) if j not in LEFT_ARM_JOINTS]
    assert g['motor_gains'] == {'robot': 0.01, 'human': 0.01}                # dressing.py:121
    # base-pose search on the person's left (right_side=False): x offset U(0, 0.5), facing backwards (yaw pi +- 30 degrees)
    base0 = np.array([-0.85, -0.4, 0]) + np.array(PR2['toc_base_pos_offset'])
    poses = [c['kw'] for c in _by(g['calls'], 'resetBasePositionAndOrientation') if c['args'] == [g['robot_body']]]
    d = np.array([p['pos'] for p in poses]) - base0
    assert len(poses) >= 50 and d[:, 0].min() >= 0 and d[:, 0].max() <= 0.5 and np.abs(d[:, 1]).max() <= 0.5 and np.allclose(d[:, 2], 0)
    yaw = np.array([2 * np.arctan2(p['orn'][2], p['orn'][3]) for p in poses])
    assert np.abs(np.abs(yaw) - np.pi).max() <= np.deg2rad(30) + 1e-9
    # start goal and the three arm goals 10 cm above shoulder / elbow / wrist with the task's orientations (dressing.py:129-134)
    iks = _by(g['calls'], 'calculateInverseKinematics')
    assert iks[0]['args'] == [g['robot_body'], PR2['ee']] and np.all(np.abs(np.array(iks[0]['kw']['targetPosition']) - [0.45, -0.3, 1.0]) <= 0.05 + 1e-12)
    assert _same_rotation(iks[0]['kw']['targetOrientation'], q_from_rpy(PR2['ee_orient_rpy']))
    assert np.allclose(iks[1]['kw']['targetPosition'], [0, 0, 0.1]) and _same_rotation(iks[1]['kw']['targetOrientation'], q_from_rpy(PR2['ee_orient_shoulder_rpy']))
    assert _same_rotation(iks[2]['kw']['targetOrientation'], q_from_rpy(PR2['ee_orient_rpy'])) and _same_rotation(iks[3]['kw']['targetOrientation'], q_from_rpy(PR2['ee_orient_rpy']))
    grip = [c for c in _by(g['calls'], 'setJointMotorControlArray') if c['args'][0] == g['robot_body']][0]['kw']
    assert grip['jointIndices'] == PR2['gripper'] and np.allclose(grip['targetPositions'], PR2['gripper_pos']) and np.allclose(grip['forces'], 500)
    # the gown: mesh scale, mass, anchors, margin, solver coefficients; placed relative to the end effector (at the origin in the recording)
    cloth = _by(g['calls'], 'loadCloth')[0]['kw']
    assert cloth['scale'] == CLOTH_SCALE == 1.4 and cloth['mass'] == DRESSING_PARAMS['total_mass'] and cloth['anchors'] == CLOTH_ANCHORS
    assert cloth['collisionMargin'] == DRESSING_PARAMS['margin'] and _same_rotation(cloth['orientation'], db.cloth_quat)
    start_ee = np.array(iks[3]['kw']['targetPosition'])                     # where the recorder reports the end effector: at the last IK goal
    assert np.allclose(cloth['position'], CLOTH_POSITION + (start_ee - CLOTH_ORIG_POS) / CLOTH_SCALE, atol=1e-9)          # dressing.py:139-146
    cp = _by(g['calls'], 'clothParams')[0]['kw']
    for k in ('kLST', 'kDP', 'kDG', 'kDF', 'kCHR', 'kKHR', 'kAHR', 'piterations'):
        assert cp[k] == DRESSING_PARAMS[k], k
    # gravity: halved while the gown settles for 50 steps, robot 0, person (0, 0, -1); 8 substeps per stepSimulation
    glob = [c['args'] for c in _by(g['calls'], 'setGravity') if not c['kw']]
    assert glob == [[0, 0, -9.81], [0, 0, -9.81 / 2], [0, 0, -9.81]] and g['n_step_simulation'] == 50
    assert inspect.signature(db.reset).parameters['settle_steps'].default == 50
    per_body = {c['kw']['body']: c['args'] for c in _by(g['calls'], 'setGravity') if 'body' in c['kw']}
    assert per_body[g['robot_body']] == [0, 0, 0] and per_body[g['human_body']] == [0, 0, -1]
    gv = np.asarray(sc['body_gravity'])
    assert np.allclose(gv[db.robot], 0) and all(np.allclose(gv[hb], [0, 0, -1]) for hb in db.humans.values())
    assert _by(g['calls'], 'setPhysicsEngineParameter')[0]['kw'] == {'numSubSteps': 8} and DressingBatch.config().num_substeps == 8


def test_drinking_scene_recipe_is_the_reference_s():
    from assistive_gym_b200.drinking_batch import JACO as DRINK_JACO
    from assistive_gym_b200.drinking_batch import N_WATER, WATER_MASS, WATER_RADIUS, DrinkingBatch
    g = G['drinking']
    db = DrinkingBatch()
    sc = db.scene
    loads, last_pose, resets = _common(g, db, DRINK_JACO, FEED_PRESET, [0.63] * 3, np.array([-0.2, -0.5, 1.1]))
    assert g['motor_gains'] == {'robot': 0.005, 'human': 0.005}              # drinking.py:126
    assert g['n_step_simulation'] == 50 == inspect.signature(db.reset).parameters['settle_steps'].default
    assert _by(g['calls'], 'setPhysicsEngineParameter')[0]['kw'] == {'numSubSteps': 4, 'numSolverIterations': 10}
    cfg = DrinkingBatch.config()
    assert cfg.num_substeps == 4 and cfg.num_solver_iters == 10
    for j in (21, 22, 23):
        assert abs(resets[j]) <= np.deg2rad(30)
    # the cup: the reference's mesh at scale 0.045, 1 kg
    cup_shape = [c['kw'] for c in _by(g['calls'], 'createCollisionShape') if c['kw'].get('fileName') == 'plastic_coffee_cup_vhacd.obj'][0]
    assert cup_shape['meshScale'] == [0.045] * 3
    mb = {c['kw']['bodies'][0]: c['kw'] for c in _by(g['calls'], 'createMultiBody')}
    assert mb[g['tool_body']]['baseMass'] == 1 == float(sc['link_mass'][int(sc['body_link0'][db.tool])])
    # the water: 64 spheres of radius 5 mm and 1 g on a 4 x 4 x 4 grid above the cup
    water = [c['kw'] for c in _by(g['calls'], 'createMultiBody') if 'batchPositions' in c['kw']][0]
    assert len(water['bodies']) == 64 == N_WATER == len(db.waters) and water['baseMass'] == WATER_MASS == 0.001
    ws = _by(g['calls'], 'createCollisionShape')[water['baseCollisionShapeIndex']]['kw']
    assert ws['shapeType'] == 2 and ws['radius'] == WATER_RADIUS == 0.005
    offsets = np.array(water['batchPositions']) - np.array(last_pose[g['tool_body']]['pos'])
    sim = OracleSim(sc, cfg, 1)
    db.reset(sim, np.random.default_rng(0), settle_steps=0, impairment='none')
    ls = sim.get_link_states([int(sc['body_link0'][db.tool])] + [int(sc['body_link0'][w]) for w in db.waters])
    assert np.allclose(ls['pos'][0, 1:] - ls['com_pos'][0, 0], offsets, atol=1e-9)
