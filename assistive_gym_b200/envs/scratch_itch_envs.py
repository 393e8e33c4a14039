from .agents.human import Human, right_arm_joints
from .agents.robot import Jaco
from .scratch_itch import ScratchItchEnv

robot_arm = 'left'
human_controllable_joint_indices = right_arm_joints         # scratch_itch_envs.py:15


class ScratchItchJacoEnv(ScratchItchEnv):
    """`assistive_gym:ScratchItchJaco-v1` (reference envs/scratch_itch_envs.py)."""

    def __init__(self, n_envs=1, device=0, seed=1001, config=None):
        super().__init__(robot=Jaco(robot_arm), human=Human(human_controllable_joint_indices, controllable=False),
                         n_envs=n_envs, device=device, seed=seed, config=config)


class ScratchItchJacoHumanEnv(ScratchItchEnv):
    """`assistive_gym:ScratchItchJacoHuman-v1` (reference envs/scratch_itch_envs.py): robot and person are both agents; `step` takes
    {'robot': a7, 'human': a10} and returns dict observations / rewards / dones (RLlib MultiAgentEnv shape, learn.py:41-59).  The
    person's right arm is driven by its action and kept inside the realistic joint limits; per-call API path."""

    def __init__(self, n_envs=1, device=0, seed=1001, config=None):
        super().__init__(robot=Jaco(robot_arm), human=Human(human_controllable_joint_indices, controllable=True),
                         n_envs=n_envs, device=device, seed=seed, config=config)
