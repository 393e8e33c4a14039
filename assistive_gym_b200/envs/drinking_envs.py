from .agents.human import Human, head_joints
from .agents.robot import Jaco
from .drinking import DrinkingEnv

robot_arm = 'right'
human_controllable_joint_indices = head_joints


class DrinkingJacoEnv(DrinkingEnv):
    """`assistive_gym:DrinkingJaco-v1` (reference envs/drinking_envs.py:27-29), per-call API path."""

    def __init__(self, n_envs=1, device=0, seed=1001, config=None):
        super().__init__(robot=Jaco(robot_arm), human=Human(human_controllable_joint_indices, controllable=False),
                         n_envs=n_envs, device=device, seed=seed, config=config)
