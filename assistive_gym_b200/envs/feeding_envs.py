from .agents.human import Human, head_joints
from .agents.robot import Jaco
from .feeding import FeedingEnv

robot_arm = 'right'
human_controllable_joint_indices = head_joints


class FeedingJacoEnv(FeedingEnv):
    """`assistive_gym:FeedingJaco-v1` (reference envs/feeding_envs.py:29-31)."""

    def __init__(self, n_envs=1, device=0, seed=1001, config=None):
        super().__init__(robot=Jaco(robot_arm), human=Human(human_controllable_joint_indices, controllable=False),
                         n_envs=n_envs, device=device, seed=seed, config=config)
