"""Parity scenarios shared by the CPU kernel-logic tests (harness build) and the GPU tests (CUDA build).

Every case runs the product implementation (`make_sim`) next to the CPU oracle on identical scenes,
seeds and actions and returns the error metrics the north star names: joint angles (rad), link
poses / contact positions (m), tool-on-body contact force (relative).
Tolerances (BASELINE.json north_star): 1e-4 rad, 1e-3 m, 5 % force over 200 substeps.
"""
import numpy as np

from assistive_gym_b200 import capi
from oracle.oracle_py import OracleSim

TOL_RAD, TOL_M, TOL_FORCE = 1e-4, 1e-3, 0.05


def synced_pair(fb, make_sim, n, seed, cfg, settle=25, threads=4, impairment='random'):
    """Oracle and product sims in the same post-reset state (oracle does the reset, state is copied)."""
    cpu = OracleSim(fb.scene, cfg, n, threads=threads)
    dev = make_sim(fb.scene, cfg, n)
    s = fb.reset(cpu, np.random.default_rng(seed), settle_steps=settle, impairment=impairment)
    fb.reset(dev, np.random.default_rng(seed), settle_steps=0, sample=s)
    dev.state_set(cpu.state_get())
    q = cpu.get_joint_states(fb.arm_links)[0]
    cpu.set_motor_targets(fb.arm_links, q)
    dev.set_motor_targets(fb.arm_links, q)
    return cpu, dev, s


def head_links(fb, s):
    """[n, 4] global link ids of each env's own head joints."""
    from assistive_gym_b200.feeding_batch import TREMOR_JOINTS
    m = np.array([fb.gl(fb.humans['male'], j) for j in TREMOR_JOINTS])
    f = np.array([fb.gl(fb.humans['female'], j) for j in TREMOR_JOINTS])
    return np.where(s['male'].astype(bool)[:, None], m, f)


def head_q(fb, sim, s):
    from assistive_gym_b200.feeding_batch import TREMOR_JOINTS
    qm = sim.get_joint_states([fb.gl(fb.humans['male'], j) for j in TREMOR_JOINTS])[0]
    qf = sim.get_joint_states([fb.gl(fb.humans['female'], j) for j in TREMOR_JOINTS])[0]
    return np.where(s['male'].astype(bool)[:, None], qm, qf)


def apply_tremor(fb, sims, s, iteration):
    """env.py:212-215: head joints of tremor envs are driven to target_joint_angles +- tremors, the
    sign flipping with the parity of the (already incremented) env-step counter."""
    from assistive_gym_b200.feeding_batch import TREMOR_JOINTS
    if not np.any(s['impairment'] == 3):
        return
    tgt = fb.tremor_rest_of(s) + (1.0 if iteration % 2 == 0 else -1.0) * s['tremors']
    for hb in fb.humans.values():
        hl = [fb.gl(hb, j) for j in TREMOR_JOINTS]
        for sim in sims:
            sim.set_motor_targets(hl, tgt)


def feeding_links(fb):
    sc = fb.scene
    foods = [int(sc['body_link0'][f]) for f in fb.foods]
    return dict(foods=foods, tool=int(sc['body_link0'][fb.tool]), bowl=int(sc['body_link0'][fb.bowl]), ee=fb.ee_link)


def take_step_targets(q, action, lower, upper, mult=0.05, frame_skip=5):
    """env.py:187-217 restated in numpy (used to drive oracle and product with identical targets)."""
    a = np.clip(action, -1, 1) * mult
    q = q.copy()
    a = a.copy()
    for _ in range(frame_skip):
        below, above = q + a < lower, q + a > upper
        a[below | above] = 0
        q = np.where(below, lower, np.where(above, upper, q))
        q = q + a
    return q


def rollout_errors(fb, make_sim, n=4, seed=0, env_steps=40, residual_threshold=0.0, foods=True, impairment='random'):
    """200 substeps (40 env steps x 5) of random actions; max errors over the rollout.

    foods=False switches the eight 1 g food spheres off: their bouncing is chaotic (a 1e-6 m
    difference in where a sphere lands decides whether it later hits the hand), so with them on, an
    fp32-vs-fp64 comparison over 200 substeps occasionally picks up a 1e-4..1e-3 rad kick on the
    weakly actuated arm (max motor torque 1 N m).  The strict north-star tolerance is asserted on the
    deterministic sub-system (arm + tool + bowl); the foods-on rollout is asserted at a looser bound
    and its measured error is reported."""
    cfg = capi.default_config(residual_threshold=residual_threshold)
    cpu, dev, s = synced_pair(fb, make_sim, n, seed, cfg, impairment=impairment)
    if not foods:
        for f in fb.foods:
            cpu.set_body_active(f, 0)
            dev.set_body_active(f, 0)
    L = feeding_links(fb)
    links = L['foods'] + [L['tool'], L['bowl'], L['ee']]
    rng = np.random.default_rng(seed + 100)
    err = dict(q=0.0, tool=0.0, ee=0.0, bowl=0.0, food=0.0, head=0.0, head_travel=0.0)
    h0 = head_q(fb, cpu, s)
    q_env = np.zeros(n)
    for it in range(env_steps):
        act = rng.uniform(-1, 1, size=(n, 7))
        apply_tremor(fb, (cpu, dev), s, it + 1)
        tgt = take_step_targets(cpu.get_joint_states(fb.arm_links)[0], act, fb.arm_lower, fb.arm_upper)
        cpu.set_motor_targets(fb.arm_links, tgt)
        dev.set_motor_targets(fb.arm_links, tgt)
        cpu.step(5)
        dev.step(5)
        a, c = cpu.get_link_states(links), dev.get_link_states(links)
        dq = np.abs(cpu.get_joint_states(fb.arm_links)[0] - dev.get_joint_states(fb.arm_links)[0])
        q_env = np.maximum(q_env, dq.max(axis=1))
        err['q'] = max(err['q'], dq.max())
        err['tool'] = max(err['tool'], np.abs(a['pos'][:, 8] - c['pos'][:, 8]).max())
        err['bowl'] = max(err['bowl'], np.abs(a['pos'][:, 9] - c['pos'][:, 9]).max())
        err['ee'] = max(err['ee'], np.abs(a['pos'][:, 10] - c['pos'][:, 10]).max())
        err['food'] = max(err['food'], np.abs(a['pos'][:, :8] - c['pos'][:, :8]).max())
        ha, hc = head_q(fb, cpu, s), head_q(fb, dev, s)
        err['head'] = max(err['head'], np.abs(ha - hc).max())
        err['head_travel'] = max(err['head_travel'], np.abs(ha - h0).max())
    err['q_env'] = q_env
    return err


def onestep_errors(fb, make_sim, n=4, seed=1, steps=30, residual_threshold=0.0):
    """State is re-synchronised before every substep: isolates the step function from chaotic drift."""
    cfg = capi.default_config(residual_threshold=residual_threshold)
    cpu, dev, _ = synced_pair(fb, make_sim, n, seed, cfg, settle=0)
    nb = fb.scene.n_bodies
    rng = np.random.default_rng(seed + 7)
    out = dict(q=0.0, qd=0.0, pos=0.0, tool_pos=0.0)
    for i in range(steps):
        if i % 5 == 0:
            tgt = cpu.get_joint_states(fb.arm_links)[0] + rng.uniform(-0.25, 0.25, size=(n, 7))
            cpu.set_motor_targets(fb.arm_links, tgt)
            dev.set_motor_targets(fb.arm_links, tgt)
        dev.state_set(cpu.state_get())
        cpu.step(1)
        dev.step(1)
        a, c = cpu.state_get(), dev.state_get().astype(np.float64)
        d = np.abs(a - c)
        base = d[:, :nb * 13].reshape(n, nb, 13)
        jq = d[:, nb * 13:].reshape(n, -1, 2)
        out['q'] = max(out['q'], jq[:, :, 0].max())
        out['qd'] = max(out['qd'], jq[:, :, 1].max())
        out['pos'] = max(out['pos'], base[:, :, :3].max())
        out['tool_pos'] = max(out['tool_pos'], base[:, fb.tool, :3].max())
    return out


def tool_contact_case(fb, make_sim, n=4, seed=2):
    """Spoon pressed against the person's head: contact positions and tool-on-body force."""
    cfg = capi.default_config(residual_threshold=0.0)
    cpu, dev, s = synced_pair(fb, make_sim, n, seed, cfg, settle=5)
    sc = fb.scene
    # teleport the spoon so that its tip overlaps the head by ~2 mm, detach it from the gripper by
    # switching the robot off, and let it be pushed by a constant gravity-like motor-free contact:
    # the spoon (mass 1, gravity 0) is given a velocity towards the head instead.
    head = {1: fb.gl(fb.humans['male'], 23), 0: fb.gl(fb.humans['female'], 23)}
    hl = np.array([head[int(m)] for m in s['male']])
    hp = np.stack([cpu.get_link_states([int(h)])['pos'][e, 0] for e, h in enumerate(hl)])
    for sim in (cpu, dev):
        sim.set_body_active(fb.robot, 0)
        for f in fb.foods:
            sim.set_body_active(f, 0)
        sim.set_base_pose(fb.tool, hp + np.array([0.0, -0.25, 0.0]), np.array([0.7071068, 0, 0, 0.7071068]))
        sim.set_base_velocity(fb.tool, np.tile([0.0, 0.6, 0.0], (n, 1)), np.zeros((n, 3)))
        sim.forward_kinematics()
    res = dict(force_rel=0.0, pos=0.0, contacts=0, force=0.0)
    for i in range(40):
        cpu.step(1)
        dev.step(1)
        for gender in ('male', 'female'):
            hb = fb.humans[gender]
            fa, fc = cpu.contact_force_sum(fb.tool, hb), dev.contact_force_sum(fb.tool, hb)
            big = fa > 0.5
            if big.any():
                res['force_rel'] = max(res['force_rel'], np.abs(fa[big] - fc[big]).max() / fa[big].max())
                res['force'] = max(res['force'], fa.max())
                ca, na = cpu.get_contacts(fb.tool, hb, max_pts=16)
                cc, nc = dev.get_contacts(fb.tool, hb, max_pts=16)
                for e in np.nonzero(big)[0]:
                    if na[e] == nc[e] and na[e] > 0:
                        res['contacts'] += int(na[e])
                        res['pos'] = max(res['pos'], np.abs(ca[e, :na[e]]['pos_a'] - cc[e, :nc[e]]['pos_a']).max())
    tl = [int(sc['body_link0'][fb.tool])]
    res['tool_pos'] = np.abs(cpu.get_link_states(tl)['pos'] - dev.get_link_states(tl)['pos']).max()
    return res


def feeding_semantics_reference(fb, sim, action, state):
    """FeedingEnv.step read-back restated in numpy on top of a sim's getters (feeding.py:12-112,
    env.py:237-274).  `state` carries foods / foods_active / iteration / task_success per env."""
    n = sim.n
    sc = fb.scene
    from assistive_gym_b200.kinematics import q_conj, q_mul, q_rot
    tool_l, robot_l = int(sc['body_link0'][fb.tool]), int(sc['body_link0'][fb.robot])
    head_m, head_f = fb.gl(fb.humans['male'], 23), fb.gl(fb.humans['female'], 23)
    ls = sim.get_link_states([tool_l, robot_l, head_m, head_f, fb.ee_link] + [int(sc['body_link0'][f]) for f in fb.foods])
    male = state['male'].astype(bool)
    hp = np.where(male[:, None], ls['pos'][:, 2], ls['pos'][:, 3])
    hq = np.where(male[:, None], ls['quat'][:, 2], ls['quat'][:, 3])
    mouth = np.where(male[:, None], fb.mouth['male'], fb.mouth['female'])
    target = hp + q_rot(hq, mouth)
    rp, rq = ls['pos'][:, 1], ls['quat'][:, 1]
    rqi = q_conj(rq)
    sp, sq = ls['com_pos'][:, 0], ls['com_quat'][:, 0]
    sp_r, sq_r = q_rot(rqi, sp - rp), q_mul(rqi, sq)
    hp_r, hq_r = q_rot(rqi, hp - rp), q_mul(rqi, hq)
    tg_r = q_rot(rqi, target - rp)
    q = sim.get_joint_states(fb.arm_links)[0]
    qw = (q + np.pi) % (2 * np.pi) - np.pi
    robot_f = np.zeros(n)
    spoon_f = np.zeros(n)
    food_hit = np.zeros((n, 8), dtype=bool)
    for hb in fb.humans.values():
        robot_f += sim.contact_force_sum(fb.robot, hb)
        spoon_f += sim.contact_force_sum(fb.tool, hb)
        for i, f in enumerate(fb.foods):
            food_hit[:, i] |= sim.get_contacts(f, hb, max_pts=1)[1] > 0
    obs = np.concatenate([sp_r, sq_r, sp_r - tg_r, qw, hp_r, hq_r, spoon_f[:, None]], axis=1)
    reward_food = np.zeros(n)
    vel_sum = np.zeros(n)
    hit_r = np.zeros(n)
    active_entry = state['active'].copy()
    for i, f in enumerate(fb.foods):
        fp = ls['pos'][:, 5 + i]
        dist = np.linalg.norm(target - fp, axis=1)
        near = sim.closest_points(f, fb.tool, 0.1, max_pts=1)[1] > 0
        infood = state['foods'][:, i]
        eaten = infood & (dist < 0.03)
        spilled = infood & ~eaten & ~near
        reward_food += 20.0 * eaten - 5.0 * spilled
        state['task_success'] += eaten
        vel_sum += eaten * np.linalg.norm(ls['lin_vel'][:, 5 + i], axis=1)
        state['foods'][:, i] &= ~(eaten | spilled)
        state['active'][:, i] &= ~eaten
        state['eaten_now'] = state.get('eaten_now', np.zeros((n, 8), dtype=bool))
        state['eaten_now'][:, i] = eaten
    for i in range(8):
        hit = active_entry[:, i] & food_hit[:, i]
        hit_r -= hit
        state['active'][:, i] &= ~hit
    ee_vel = np.linalg.norm(ls['lin_vel'][:, 4], axis=1)
    total = robot_f + spoon_f
    pref = 0.25 * (-ee_vel) + 0.01 * (-total) + 0.05 * np.where(spoon_f < 10, 0.0, -spoon_f) + 1.0 * hit_r + 1.0 * (-vel_sum)
    reward = -np.linalg.norm(target - sp, axis=1) - 0.01 * np.linalg.norm(action, axis=1) + reward_food + pref
    state['iteration'] += 1
    done = state['iteration'] >= 200
    return obs, reward, done, total


def readback_errors(fb, make_sim, n=4, seed=6):
    """every read-back call of agents/agent.py (getJointStates, getLinkState incl. COM frame and
    velocities, getContactPoints, getClosestPoints) on the product next to the oracle, from the same state."""
    cfg = capi.default_config(residual_threshold=0.0)
    cpu, dev, s = synced_pair(fb, make_sim, n, seed, cfg, settle=25)
    rng = np.random.default_rng(seed)
    tgt = cpu.get_joint_states(fb.arm_links)[0] + rng.uniform(-0.2, 0.2, size=(n, 7))
    for sim in (cpu, dev):
        sim.set_motor_targets(fb.arm_links, tgt)
    dev.state_set(cpu.state_get())
    cpu.step(1)
    dev.step(1)
    out = {}
    sc = fb.scene
    l0, nl = int(sc['body_link0'][fb.robot]), int(sc['body_nlinks'][fb.robot])
    links = list(range(l0, l0 + nl)) + [int(sc['body_link0'][fb.tool]), int(sc['body_link0'][fb.bowl])]
    qa, qda, ta = cpu.get_joint_states(fb.arm_links + fb.gripper_links)
    qb, qdb, tb = dev.get_joint_states(fb.arm_links + fb.gripper_links)
    out.update(q=np.abs(qa - qb).max(), qd=np.abs(qda - qdb).max(), tau=np.abs(ta - tb).max(), tau_max=np.abs(ta).max())
    a, b = cpu.get_link_states(links), dev.get_link_states(links)
    for k in ('pos', 'com_pos', 'lin_vel', 'ang_vel'):
        out[k] = np.abs(a[k] - b[k]).max()
    for k in ('quat', 'com_quat'):
        out[k] = np.minimum(np.abs(a[k] - b[k]).max(axis=-1), np.abs(a[k] + b[k]).max(axis=-1)).max()
    # contacts of the bowl (rests on the table) and of the tool (holds the food): same count, positions, forces
    cerr, ferr, ncontacts = 0.0, 0.0, 0
    for body in (fb.bowl, fb.tool):
        ca, na = cpu.get_contacts(body, max_pts=32)
        cb, nb_ = dev.get_contacts(body, max_pts=32)
        out['count_equal'] = out.get('count_equal', True) and bool(np.array_equal(na, nb_))
        for e in range(n):
            if na[e] != nb_[e] or na[e] == 0:
                continue
            # order by (link_b, position) so the comparison does not depend on the listing order
            ka = np.lexsort(np.round(ca[e, :na[e]]['pos_a'], 4).T[::-1]); kb = np.lexsort(np.round(cb[e, :nb_[e]]['pos_a'], 4).T[::-1])
            cerr = max(cerr, np.abs(ca[e, :na[e]]['pos_a'][ka] - cb[e, :nb_[e]]['pos_a'][kb]).max())
            ferr = max(ferr, np.abs(ca[e, :na[e]]['normal_force'][ka] - cb[e, :nb_[e]]['normal_force'][kb]).max())
            ncontacts += int(na[e])
    out.update(contact_pos=cerr, contact_force=ferr, n_contacts=ncontacts)
    # closest points: robot vs table / wheelchair within 0.3 m, tool vs bowl within 1 m
    derr = 0.0
    for ba, bb, dist in ((fb.robot, fb.table, 0.3), (fb.robot, fb.wheelchair, 0.3), (fb.tool, fb.bowl, 1.0)):
        pa, na = cpu.closest_points(ba, bb, dist, max_pts=64)
        pb, nb_ = dev.closest_points(ba, bb, dist, max_pts=64)
        out['count_equal'] = out['count_equal'] and bool(np.array_equal(na, nb_))
        for e in range(n):
            if na[e] and na[e] == nb_[e]:
                derr = max(derr, abs(np.sort(pa[e, :na[e]]['distance'])[0] - np.sort(pb[e, :nb_[e]]['distance'])[0]))
    out['closest_dist'] = derr
    return out


def population_errors(fb, make_sim, n=1024, seed=21, env_steps=40, threads=8, f32_control=True):
    """The BENCHMARKED configuration (bench.py: foods on, default early exit at 1e-7, random actions), 200 substeps, as a
    population: per-env maximum error of the product against the fp64 oracle, and -- the control -- of the fp32 build
    of the SAME oracle against the fp64 one, which measures how much of the spread is fp32-vs-fp64 sensitivity
    (1 g food spheres bouncing, active-set flips at the early exit) rather than the CUDA implementation.
    Returns {name: {'q': [n], 'tool': [n], 'ee': [n]}} for name in ('product', 'oracle_f32')."""
    cfg = capi.default_config()
    cpu, dev, s = synced_pair(fb, make_sim, n, seed, cfg, threads=threads)
    sims = {'product': dev}
    if f32_control:
        c32 = OracleSim(fb.scene, cfg, n, f32=True, threads=threads)
        fb.reset(c32, np.random.default_rng(seed), settle_steps=0, sample=s)
        c32.state_set(cpu.state_get())
        c32.set_motor_targets(fb.arm_links, cpu.get_joint_states(fb.arm_links)[0])
        sims['oracle_f32'] = c32
    L = feeding_links(fb)
    links = [L['tool'], L['ee']]
    rng = np.random.default_rng(seed + 100)
    out = {k: dict(q=np.zeros(n), tool=np.zeros(n), ee=np.zeros(n)) for k in sims}
    for it in range(env_steps):
        act = rng.uniform(-1, 1, size=(n, 7))
        apply_tremor(fb, (cpu,) + tuple(sims.values()), s, it + 1)
        tgt = take_step_targets(cpu.get_joint_states(fb.arm_links)[0], act, fb.arm_lower, fb.arm_upper)
        for sim in (cpu,) + tuple(sims.values()):
            sim.set_motor_targets(fb.arm_links, tgt)
            sim.step(5)
        qa, pa = cpu.get_joint_states(fb.arm_links)[0], cpu.get_link_states(links)['pos']
        for k, sim in sims.items():
            dq = np.abs(qa - sim.get_joint_states(fb.arm_links)[0]).max(axis=1)
            dp = np.abs(pa - sim.get_link_states(links)['pos']).max(axis=2)
            out[k]['q'] = np.maximum(out[k]['q'], dq)
            out[k]['tool'] = np.maximum(out[k]['tool'], dp[:, 0])
            out[k]['ee'] = np.maximum(out[k]['ee'], dp[:, 1])
    return out


def population_summary(e):
    """median / p90 / p99 / max and the fraction of envs inside the north-star tolerances"""
    r = {}
    for k, tol in (('q', TOL_RAD), ('tool', TOL_M), ('ee', TOL_M)):
        x = e[k]
        r[k] = dict(median=float(np.median(x)), p90=float(np.quantile(x, 0.9)), p99=float(np.quantile(x, 0.99)), max=float(x.max()),
                    within=float((x < tol).mean()))
    return r
